"""ctypes binding of libnirrt_b200.so (include/nirrt_b200.h).  No fallback: if the library is
missing or no sm_100 device is visible, every compute entry point raises."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_i64p = C.POINTER(C.c_int64)
c_u32p = C.POINTER(C.c_uint32)
c_u8p = C.POINTER(C.c_uint8)
c_fp = C.POINTER(C.c_float)

MAX_OBSTACLES = 32


class NirrtError(RuntimeError):
    pass


class Pn2Layer(C.Structure):
    _fields_ = [("weight", c_fp), ("bias", c_fp), ("bn_weight", c_fp), ("bn_bias", c_fp), ("bn_mean", c_fp),
                ("bn_var", c_fp), ("c_in", C.c_int), ("c_out", C.c_int)]


class BatchDesc(C.Structure):
    _fields_ = [("dim", C.c_int), ("n_envs", C.c_int), ("capacity", C.c_int), ("record_capacity", C.c_int),
                ("near_capacity", C.c_int), ("device", C.c_int)]


def lib():
    """Loads (building if sources are newer and nvcc exists) the C-ABI library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    variant = os.environ.get("NIRRT_LIB_VARIANT")
    if variant:      # development build (profiles/tools/phase_timing.py); must have been built explicitly
        path = os.path.join(os.path.dirname(_build.LIB), f"libnirrt_b200_{variant}.so")
        if not os.path.exists(path):
            raise NirrtError(f"{path} not found: build it with nirrt_star_b200.build.build_variant")
    elif not os.path.exists(path) or _build.needs_build():
        # sources newer than the binary (or no binary): rebuild under a file lock (several ranks may get here at
        # once); a stale binary is never loaded silently -- its ABI may no longer match the argtypes below
        try:
            _build.build_locked()
        except Exception as exc:  # pragma: no cover - build environment problem
            raise NirrtError(f"libnirrt_b200.so is missing or older than its sources and could not be rebuilt: {exc}") from exc
    L = C.CDLL(path)
    V = C.c_void_p
    L.nirrt_last_error.restype = C.c_char_p
    L.nirrt_batch_create.argtypes = [C.POINTER(BatchDesc), C.POINTER(V)]
    L.nirrt_batch_destroy.argtypes = [V]
    L.nirrt_batch_set_problems.argtypes = [V, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip, c_dp, c_dp, c_ip, c_dp, c_dp, c_dp, V]
    L.nirrt_batch_set_problems_2d.argtypes = [V, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ip, c_dp, c_ip, c_dp, c_dp, c_dp, V]
    L.nirrt_batch_set_py_rng.argtypes = [V, c_u32p, c_ip, V]
    L.nirrt_batch_get_py_rng_sync.argtypes = [V, c_u32p, c_ip, V]
    L.nirrt_batch_set_rng.argtypes = [V, c_u32p, c_ip, V]
    L.nirrt_batch_get_rng_sync.argtypes = [V, c_u32p, c_ip, V]
    L.nirrt_batch_set_guidance.argtypes = [V, C.c_double, C.c_double]
    L.nirrt_batch_set_cloud.argtypes = [V, C.c_int, c_dp, C.c_int, V]
    L.nirrt_batch_load_trees.argtypes = [V, C.c_int, C.c_int, c_ip, c_dp, c_i64p, V]
    L.nirrt_batch_read_trees_sync.argtypes = [V, C.c_int, C.c_int, c_ip, c_dp, c_i64p, V]
    L.nirrt_batch_begin.argtypes = [V, C.c_int, C.c_int, C.c_int, C.c_int, V]
    L.nirrt_batch_run.argtypes = [V, C.c_int, V]
    L.nirrt_batch_status_sync.argtypes = [V, c_ip, c_ip, V]
    L.nirrt_batch_env_state_sync.argtypes = [V, c_ip, c_ip, c_ip, V]
    L.nirrt_batch_read_records_sync.argtypes = [V, C.c_int, C.c_int, c_dp, c_ip, V]
    L.nirrt_batch_read_solutions_sync.argtypes = [V, C.c_int, c_i64p, C.c_int, V]
    L.nirrt_batch_goal_parent_sync.argtypes = [V, c_i64p, c_dp, V]
    L.nirrt_batch_read_trace_sync.argtypes = [V, c_ip, c_ip, c_ip, c_ip, C.c_int, c_dp, V]
    L.nirrt_collide_edges_sync.argtypes = [V, C.c_int, c_dp, C.c_int64, c_u8p, V]
    L.nirrt_points_check_sync.argtypes = [V, C.c_int, C.c_int, c_dp, C.c_int64, c_u8p, V]
    L.nirrt_nearest_sync.argtypes = [V, C.c_int, c_dp, C.c_int64, c_i64p, V]
    L.nirrt_within_sync.restype = C.c_int64
    L.nirrt_within_sync.argtypes = [V, C.c_int, c_dp, C.c_double, c_i64p, C.c_int64, V]
    L.nirrt_costs_sync.argtypes = [V, C.c_int, c_i64p, C.c_int64, c_dp, V]
    L.nirrt_batch_read_cbest_sync.argtypes = [V, c_dp, c_dp, V]
    L.nirrt_fps_f64_sync.argtypes = [c_dp, C.c_int64, C.c_int, C.c_int, c_i64p, V]
    L.nirrt_batch_set_free_masks.argtypes = [V, c_u8p, C.c_int, C.c_int, V]
    L.nirrt_batch_sample_clouds_sync.argtypes = [V, c_ip, C.c_int, c_ip, c_dp, C.c_int, C.c_int, C.c_double, V, V, V, c_ip, V]
    L.nirrt_batch_read_sampled_clouds_sync.argtypes = [V, C.c_int, C.c_int, c_dp, V]
    L.nirrt_batch_commit_clouds.argtypes = [V, V, c_ip, C.c_int, V]
    L.nirrt_sincos_sync.argtypes = [c_dp, C.c_int64, c_dp, c_dp, V]
    L.nirrt_atan2_sync.argtypes = [c_dp, c_dp, C.c_int64, c_dp, V]
    L.nirrt_batch_set_stop_threshold.argtypes = [V, C.c_double]
    L.nirrt_batch_set_vertex_limit.argtypes = [V, C.c_int]
    L.nirrt_batch_run_profiled_sync.argtypes = [V, C.c_int, c_fp, V]
    L.nirrt_batch_counters.argtypes = [V, c_i64p, c_i64p]
    L.nirrt_batch_work_stats_sync.argtypes = [V, c_i64p, V]
    L.nirrt_batch_graph_stats.argtypes = [V, c_i64p, c_i64p, c_i64p]
    L.nirrt_batch_time_scan_sync.argtypes = [V, C.c_int, C.c_int, c_fp, c_i64p, V]
    # PointNet++ (include/nirrt_pointnet2.h)
    c_i32p = C.POINTER(C.c_int32)
    c_u16p = C.POINTER(C.c_uint16)
    L.nirrt_pn2_create.argtypes = [C.POINTER(Pn2Layer), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(V)]
    L.nirrt_pn2_destroy.argtypes = [V]
    L.nirrt_pn2_set_n_points.argtypes = [V, C.c_int]
    L.nirrt_pn2_classify_sync.argtypes = [V, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_i32p, c_i64p, c_fp, c_fp, V]
    L.nirrt_pn2_classify_device.argtypes = [V, C.c_int, C.c_int, V, V, V, V, V, V, V, V]
    L.nirrt_pn2_read_buffer_sync.restype = C.c_int64
    L.nirrt_pn2_read_buffer_sync.argtypes = [V, C.c_char_p, V, C.c_int64, V]
    L.nirrt_pn2_set_profiling.argtypes = [V, C.c_int]
    L.nirrt_pn2_last_stage_ms.argtypes = [V, c_fp]
    L.nirrt_pn2_launch_count.restype = C.c_int64
    L.nirrt_pn2_launch_count.argtypes = [V]
    L.nirrt_connect_analyse_batch_sync.argtypes = [c_fp, c_ip, C.c_int, C.c_int, C.c_int, c_u8p, c_fp, c_fp, C.c_float, c_ip, c_u8p, c_u8p, V]
    L.nirrt_connect_masks_device.argtypes = [V, C.c_int, C.c_int, C.c_int, V, C.c_float, V, V, V]
    L.nirrt_connect_trial_device.argtypes = [V, C.c_int, C.c_int, C.c_int, V, V, V, V, V, C.c_float, V, V, c_ip, c_ip, c_u8p, V]
    L.nirrt_connect_analyse_sync.argtypes = [c_fp, C.c_int, C.c_int, c_u8p, c_fp, c_fp, C.c_float, c_ip, c_u8p, c_u8p, V]
    L.nirrt_gemm_f16_sync.argtypes = [c_u16p, c_u16p, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_u16p, V]
    _LIB = L
    return L


def check(rc):
    if rc < 0:
        raise NirrtError(f"libnirrt_b200 error {rc}: {lib().nirrt_last_error().decode()}")
    return rc


def require_device():
    n = lib().nirrt_device_count()
    if n <= 0:
        raise NirrtError("no sm_100 (B200) CUDA device visible: nirrt_star_b200 has no CPU fallback")
    return n


def fp(a):
    return a.ctypes.data_as(c_fp)


def dp(a):
    return a.ctypes.data_as(c_dp)


def ip(a):
    return a.ctypes.data_as(c_ip)


def i64p(a):
    return a.ctypes.data_as(c_i64p)


def u8p(a):
    return a.ctypes.data_as(c_u8p)


def f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)
