"""Builds libnirrt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnirrt_b200.so")

# -fmad=false: planner arithmetic must round every operation (explicit intrinsics already do;
# this keeps the compiler from contracting anything that slips through).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]
UNITS = [("planner3d.cu", ["-fmad=false"]), ("pointnet2.cu", [])]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libnirrt_b200.so cannot be built")
    return exe


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "nirrt_b200.h"))
    out.append(os.path.join(os.path.dirname(HERE), "include", "nirrt_pointnet2.h"))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build_locked():
    """build() under an exclusive file lock: concurrent ranks wait for the first one instead of all writing
    csrc/*.o and the library at once; whoever gets the lock second finds nothing left to do."""
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return build(force=False)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def build_variant(name, extra_flags):
    """Development builds beside the product library (e.g. -DNIRRT_PHASE_TIMING): libnirrt_b200_<name>.so, selected at
    load time with NIRRT_LIB_VARIANT=<name> (profiles/tools/phase_timing.py).  Never used by tests or the bench."""
    lib = os.path.join(HERE, f"libnirrt_b200_{name}.so")
    objs = []
    for unit, extra in UNITS:
        obj = os.path.join(CSRC, unit.replace(".cu", f".{name}.o"))
        subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                               "-Xcompiler", "-fPIC", "-c", os.path.join(CSRC, unit), "-o", obj] + extra + list(extra_flags))
        objs.append(obj)
    subprocess.check_call([_nvcc(), "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
    return lib


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    tmp_lib = LIB + f".tmp{os.getpid()}"
    for name, extra in UNITS:
        obj = os.path.join(CSRC, name.replace(".cu", ".o"))
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "-c", os.path.join(CSRC, name), "-o", obj] + extra
        if verbose:
            cmd += ["-Xptxas", "-v"]
        subprocess.check_call(cmd)
        objs.append(obj)
    subprocess.check_call([_nvcc(), "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp_lib] + objs)
    os.replace(tmp_lib, LIB)      # atomic: a concurrent loader sees the old or the new library, never a partial one
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
