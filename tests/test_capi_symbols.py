"""CPU-only: libnirrt_b200.so builds, loads, and exports every function include/nirrt_b200.h
declares; compute entry points refuse to run without an sm_100 device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for hdr in sorted(os.listdir(os.path.join(ROOT, "include"))):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(nirrt\w*|pn2\w*)\s*\(", text)
    return sorted(set(n for n in names if not n.endswith("_desc")))


def test_library_exports_every_declared_symbol():
    from nirrt_star_b200 import _lib
    L = _lib.lib()
    decl = _declared()
    assert len(decl) >= 20
    missing = [n for n in decl if not hasattr(L, n)]
    assert not missing, missing
    assert L.nirrt_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nirrt_star_b200 import _lib, batch
    from nirrt_star_b200.synthetic import make_problem_3d
    assert _lib.lib().nirrt_device_count() == 0
    with pytest.raises(_lib.NirrtError):
        batch.BatchPlanner3D([make_problem_3d(0)], 10)
