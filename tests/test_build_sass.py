"""SASS checks on the built library (no GPU needed): instruction selections the numerics rely on."""
import re
import shutil
import subprocess

import pytest

from nirrt_star_b200 import build


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if build.needs_build():
        build.build_locked()
    try:
        txt = subprocess.run([exe, "-sass", build.LIB], check=True, capture_output=True, text=True).stdout
    except (OSError, subprocess.CalledProcessError) as e:
        pytest.skip(f"cuobjdump unavailable: {e}")
    funcs = {}
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name, body = part.split("\n", 1)
        funcs[name.strip()] = body
    return funcs


def test_fps_squares_are_rounded_before_they_are_summed(sass):
    """farthest_point_sample sums (dx*dx, dy*dy, dz*dz) with every product rounded (pointnet2_utils.py:81).  ptxas
    contracts packed multiplies feeding packed adds into FFMA2 even for .rn operands, so k_fps adds the unpacked
    halves with scalar FADD: packed subtract / multiply yes, FFMA2 (and scalar FFMA) never."""
    ks = {n: b for n, b in sass.items() if re.match(r"_Z5k_fpsILi\d+ELi\d+EE", n)}
    assert len(ks) >= 5
    for n, b in ks.items():
        assert "FFMA2" not in b and not re.search(r"\bFFMA\b", b), n
        assert "FMUL2" in b and "FADD2" in b, n


def test_pair_distance_kernels_use_the_packed_fp32_pipe(sass):
    """the exhaustive query_ball_point scan and the 3-NN search evaluate -2 * (a . b) + |a|^2 + |b|^2 two points per
    instruction."""
    bq = [b for n, b in sass.items() if n.startswith("_Z15k_ball_query_bfPKf")]
    knn = [b for n, b in sass.items() if n.startswith("_Z8k_interpILi1EE")]
    assert len(bq) == 1 and len(knn) == 1
    for b in bq + knn:
        assert "FFMA2" in b and "FMUL2" in b
