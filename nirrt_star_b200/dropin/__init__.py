"""Drop-in import tree: the reference's module paths, class names and signatures, backed by the
CUDA engine.  Put this directory FIRST on sys.path (``nirrt_star_b200.dropin.install()``) and the
reference's drivers (eval_planning_3d.py, demo_planning_3d.py) import these modules instead of
their numpy originals -- see INTEGRATION.md."""
import os
import sys

DROPIN_DIR = os.path.dirname(os.path.abspath(__file__))


def install():
    if DROPIN_DIR in sys.path:
        sys.path.remove(DROPIN_DIR)
    sys.path.insert(0, DROPIN_DIR)
    return DROPIN_DIR
