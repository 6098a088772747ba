"""Drop-in import tree: the reference's module paths, class names and signatures, backed by the
CUDA engine.  Put this directory FIRST on sys.path (``nirrt_star_b200.dropin.install()``) and the
reference's drivers (eval_planning_3d.py, demo_planning_3d.py) import these modules instead of
their numpy originals -- see INTEGRATION.md."""
import os
import sys
import types

DROPIN_DIR = os.path.dirname(os.path.abspath(__file__))


def install():
    """Puts the drop-in tree first on sys.path.  The reference's top-level ``datasets`` /
    ``datasets_3d`` are namespace directories (no __init__.py) that a site-packages ``datasets``
    distribution would shadow, so they are registered explicitly: drop-in modules first, then any
    directory of that name further down sys.path (the reference checkout, when present) so that
    the parts this repo does not replace (problem builders) keep resolving."""
    if DROPIN_DIR in sys.path:
        sys.path.remove(DROPIN_DIR)
    sys.path.insert(0, DROPIN_DIR)
    for pkg in ("datasets", "datasets_3d"):
        paths = [os.path.join(DROPIN_DIR, pkg)]
        for d in sys.path[1:]:
            cand = os.path.join(d or ".", pkg)
            if os.path.isdir(cand) and not os.path.exists(os.path.join(cand, "__init__.py")) and cand not in paths:
                paths.append(cand)
        m = sys.modules.get(pkg)
        if m is None or getattr(m, "__file__", None):      # absent, or a regular (foreign) package
            m = types.ModuleType(pkg)
            sys.modules[pkg] = m
        m.__path__ = paths
    return DROPIN_DIR
