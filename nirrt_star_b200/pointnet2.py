"""Host side of the PointNet++ guidance-state inference path (C ABI: include/nirrt_pointnet2.h).

``PointNet2Engine`` owns one ``nirrt_pn2`` handle: the network's weights (BatchNorm folded, fp16
K-major rows in HBM) and the activation buffers for up to ``max_batch`` clouds of ``n_points``
points.  ``classify`` is the batched equivalent of the reference's
``PNGWrapper.classify_path_points`` (wrapper{,_3d}/pointnet_pointnet2/pointnet2_wrapper.py:28-64);
the single-cloud drop-in wrappers live in nirrt_star_b200/dropin/wrapper{,_3d}.

There is no CPU fallback: constructing an engine without an sm_100 device raises.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import NirrtError, check, fp, i64p

SA_MLPS = (((16, 16, 32), (32, 32, 64)), ((64, 64, 128), (64, 96, 128)),
           ((128, 196, 256), (128, 196, 256)), ((256, 256, 512), (256, 384, 512)))   # pointnet2.py:11-14
FP_LAYERS = (("fp4", 2), ("fp3", 2), ("fp2", 2), ("fp1", 3))                          # pointnet2.py:15-18
NPOINTS = (1024, 256, 64)   # sizes the 2nd..4th farthest_point_sample calls draw their start index from


def layer_names():
    """(conv prefix, bn prefix or None) of the 35 convolutions in nirrt_pn2_create's order."""
    out = []
    for l in range(1, 5):
        for s in range(2):
            for j in range(3):
                out.append((f"sa{l}.conv_blocks.{s}.{j}", f"sa{l}.bn_blocks.{s}.{j}"))
    for name, n in FP_LAYERS:
        for j in range(n):
            out.append((f"{name}.mlp_convs.{j}", f"{name}.mlp_bns.{j}"))
    out.append(("conv1", "bn1"))
    out.append(("conv2", None))
    return out


def _np(t):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(t), dtype=np.float32)


def draw_fps_starts(batch, n_points):
    """The four start indices per cloud, drawn exactly as the reference's forward draws them:
    torch.randint(0, N, (B,)) on torch's default CPU generator, once per SA level
    (pointnet2_utils.py:77), so the global torch RNG stream advances identically."""
    import torch
    cols = [torch.randint(0, n, (batch,), dtype=torch.long).numpy() for n in (n_points,) + NPOINTS]
    return np.ascontiguousarray(np.stack(cols, axis=1), dtype=np.int32)


class PointNet2Engine:
    def __init__(self, state_dict, n_points=2048, max_batch=1, device=0, stream=None):
        _lib.require_device()
        self.L = _lib.lib()
        self.n_points, self.max_batch = int(n_points), int(max_batch)
        self.capacity = int(n_points)      # largest cloud the buffers hold; set_n_points selects any size up to it
        self.stream = C.c_void_p(stream) if stream else None
        names = layer_names()
        layers = (_lib.Pn2Layer * len(names))()
        self._keep = []
        for k, (conv, bn) in enumerate(names):
            w = _np(state_dict[conv + ".weight"])
            w = w.reshape(w.shape[0], -1)
            arrs = [np.ascontiguousarray(w), _np(state_dict[conv + ".bias"])]
            if bn is not None:
                arrs += [_np(state_dict[bn + ".weight"]), _np(state_dict[bn + ".bias"]),
                         _np(state_dict[bn + ".running_mean"]), _np(state_dict[bn + ".running_var"])]
            self._keep.append(arrs)
            layers[k].weight, layers[k].bias = fp(arrs[0]), fp(arrs[1])
            if bn is not None:
                layers[k].bn_weight, layers[k].bn_bias, layers[k].bn_mean, layers[k].bn_var = (fp(a) for a in arrs[2:])
            layers[k].c_in, layers[k].c_out = w.shape[1], w.shape[0]
        h = C.c_void_p()
        check(self.L.nirrt_pn2_create(layers, len(names), self.n_points, self.max_batch, int(device), C.byref(h)))
        self.h = h

    def set_n_points(self, n_points):
        """Cloud size of the following classify calls (16 .. capacity)."""
        if int(n_points) != self.n_points:
            check(self.L.nirrt_pn2_set_n_points(self.h, int(n_points)))
            self.n_points = int(n_points)

    def close(self):
        if getattr(self, "h", None):
            self.L.nirrt_pn2_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def classify(self, pc, start_mask, goal_mask, fps_start=None, return_logp=False, out=None):
        """pc [B][N][2|3] f32, masks [B][N] f32 -> (path_pred int64 [B][N], path_score f32 [B][N]
        [, log-probabilities f32 [B][N][2]]).  Host arrays in, host arrays out.  ``out=(pred, score)``
        lets the caller supply the result arrays (e.g. pinned memory, which the device-to-host copies
        reach at full PCIe rate); otherwise fresh arrays are returned like the reference's wrapper does."""
        pc = np.ascontiguousarray(pc, dtype=np.float32)
        if pc.ndim == 2:
            pc = pc[None]
        B, N, dim = pc.shape
        if N != self.n_points:
            raise ValueError(f"engine was built for {self.n_points}-point clouds, got {N}")
        sm = np.ascontiguousarray(start_mask, dtype=np.float32).reshape(B, N)
        gm = np.ascontiguousarray(goal_mask, dtype=np.float32).reshape(B, N)
        if fps_start is None:
            fps_start = draw_fps_starts(B, N)
        fs = np.ascontiguousarray(fps_start, dtype=np.int32).reshape(B, 4)
        if out is not None:
            pred, score = out
            if pred.shape != (B, N) or pred.dtype != np.int64 or score.shape != (B, N) or score.dtype != np.float32 \
                    or not pred.flags.c_contiguous or not score.flags.c_contiguous:
                raise ValueError("out=(pred int64 [B][N], score float32 [B][N]) C-contiguous arrays expected")
        else:
            pred = np.empty((B, N), dtype=np.int64)
            score = np.empty((B, N), dtype=np.float32)
        logp = np.zeros((B, N, 2), dtype=np.float32) if return_logp else None
        check(self.L.nirrt_pn2_classify_sync(self.h, B, dim, fp(pc), fp(sm), fp(gm),
                                             fs.ctypes.data_as(C.POINTER(C.c_int32)), i64p(pred), fp(score),
                                             fp(logp) if return_logp else None, self.stream))
        return (pred, score, logp) if return_logp else (pred, score)

    def classify_device(self, batch, dim, pc, sm, gm, fps_start, pred, score, logp=None):
        """Device-pointer variant (ints from tensor.data_ptr()); asynchronous on the engine's stream."""
        check(self.L.nirrt_pn2_classify_device(self.h, int(batch), int(dim), pc, sm, gm, fps_start, pred, score, logp,
                                               self.stream))

    def read_buffer(self, name, dtype, shape):
        out = np.zeros(shape, dtype=dtype)
        size = self.L.nirrt_pn2_read_buffer_sync(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes, self.stream)
        check(int(size))
        if size != out.nbytes:
            raise NirrtError(f"buffer {name}: {size} bytes on device, {out.nbytes} requested")
        return out

    def set_profiling(self, enabled):
        check(self.L.nirrt_pn2_set_profiling(self.h, int(bool(enabled))))

    def stage_ms(self):
        ms = (C.c_float * 8)()
        check(self.L.nirrt_pn2_last_stage_ms(self.h, ms))
        return dict(zip(("prep", "fps", "ball_query", "gather", "sa_mlp", "interp", "fp_mlp", "head"), [float(x) for x in ms]))

    def launches(self):
        return int(self.L.nirrt_pn2_launch_count(self.h))


def gemm_f16(A, W, bias, mode=0, group=16, stream=None):
    """relu(A @ W.T + bias) on the tensor cores (mode 0), or its max over groups of rows (mode 1).
    A [m][k], W [n][k] float16; returns float16."""
    _lib.require_device()
    L = _lib.lib()
    A = np.ascontiguousarray(A, dtype=np.float16); W = np.ascontiguousarray(W, dtype=np.float16)
    bias = np.ascontiguousarray(bias, dtype=np.float32)
    m, k = A.shape; n = W.shape[0]
    out = np.zeros((m if mode == 0 else m // group, n), dtype=np.float16)
    u16 = C.POINTER(C.c_uint16)
    check(L.nirrt_gemm_f16_sync(A.ctypes.data_as(u16), W.ctypes.data_as(u16), fp(bias), m, n, k, int(mode), int(group),
                                out.ctypes.data_as(u16), C.c_void_p(stream) if stream else None))
    return out


def connect_analyse_batch(pcs, path_masks, srcs, dsts, radius, stream=None):
    """connect_analyse for a list of (cloud, mask, src, dst) tuples in ONE launch (one CTA each).
    Returns (has_path bool[B], [visited_mask f32[n_b]], [boundary_mask f32[n_b]])."""
    _lib.require_device()
    L = _lib.lib()
    B = len(pcs)
    dim = np.asarray(pcs[0]).shape[1]
    n_pts = np.array([len(p) for p in pcs], dtype=np.int32)
    n_max = int(n_pts.max())
    pc = np.zeros((B, n_max, dim), dtype=np.float32); pm = np.zeros((B, n_max), dtype=np.uint8)
    s3 = np.zeros((B, 3), dtype=np.float32); d3 = np.zeros((B, 3), dtype=np.float32)
    for b in range(B):
        pc[b, :n_pts[b]] = pcs[b]; pm[b, :n_pts[b]] = np.asarray(path_masks[b]) != 0
        s3[b, :dim] = srcs[b]; d3[b, :dim] = dsts[b]
    hp = np.zeros(B, dtype=np.int32); vis = np.zeros((B, n_max), dtype=np.uint8); bnd = np.zeros((B, n_max), dtype=np.uint8)
    check(L.nirrt_connect_analyse_batch_sync(fp(pc), n_pts.ctypes.data_as(_lib.c_ip), n_max, dim, B, _lib.u8p(pm), fp(s3), fp(d3),
                                             C.c_float(float(radius)), hp.ctypes.data_as(_lib.c_ip), _lib.u8p(vis), _lib.u8p(bnd),
                                             C.c_void_p(stream) if stream else None))
    return hp.astype(bool), [vis[b, :n_pts[b]].astype(np.float32) for b in range(B)], [bnd[b, :n_pts[b]].astype(np.float32) for b in range(B)]


def connect_masks_device(d_pc, n_points, dim, batch, d_src, radius, d_start_mask, d_goal_mask, stream=None):
    """float32 start / goal neighbourhood masks of the first Neural Connect trial (nirrt_connect_masks_device)."""
    _lib.require_device()
    check(_lib.lib().nirrt_connect_masks_device(C.c_void_p(d_pc), n_points, dim, batch, C.c_void_p(d_src), C.c_float(float(radius)),
                                                C.c_void_p(d_start_mask), C.c_void_p(d_goal_mask), C.c_void_p(stream) if stream else None))


def connect_trial_device(d_pc, n_points, dim, batch, d_active, d_path_mask, d_pred, d_src, d_dst, radius, d_start_mask, d_goal_mask,
                         stream=None):
    """One Neural Connect trial on device buffers (nirrt_connect_trial_device; all d_* are device addresses).
    Returns (has_path int32[2*batch], ties int32[2*batch], tie_boundary u8[2*batch][n_points])."""
    _lib.require_device()
    L = _lib.lib()
    hp = np.zeros(2 * batch, dtype=np.int32); ties = np.zeros(2 * batch, dtype=np.int32)
    tb = np.zeros((2 * batch, n_points), dtype=np.uint8)
    check(L.nirrt_connect_trial_device(C.c_void_p(d_pc), n_points, dim, batch, C.c_void_p(d_active), C.c_void_p(d_path_mask),
                                       C.c_void_p(d_pred), C.c_void_p(d_src), C.c_void_p(d_dst), C.c_float(float(radius)),
                                       C.c_void_p(d_start_mask), C.c_void_p(d_goal_mask), hp.ctypes.data_as(_lib.c_ip),
                                       ties.ctypes.data_as(_lib.c_ip), _lib.u8p(tb), C.c_void_p(stream) if stream else None))
    return hp, ties, tb


def connect_analyse(pc, path_mask, src, dst, radius, stream=None):
    """(has_path, visited_mask f32 [n], boundary_mask f32 [n]) of the r-disc graph over
    [src, dst, pc[path_mask]] -- bfs_point_cloud + get_boundary_mask of the reference's Neural Connect
    (wrapper/utils/bfs_connect_heuristic.py) in one CUDA launch."""
    _lib.require_device()
    L = _lib.lib()
    pc = np.ascontiguousarray(pc, dtype=np.float32)
    n, dim = pc.shape
    pm = np.ascontiguousarray(np.asarray(path_mask) != 0, dtype=np.uint8)
    s = np.ascontiguousarray(src, dtype=np.float32).reshape(dim); d = np.ascontiguousarray(dst, dtype=np.float32).reshape(dim)
    hp = C.c_int(0)
    vis = np.zeros(n, dtype=np.uint8); bnd = np.zeros(n, dtype=np.uint8)
    check(L.nirrt_connect_analyse_sync(fp(pc), n, dim, _lib.u8p(pm), fp(s), fp(d), C.c_float(float(radius)), C.byref(hp),
                                       _lib.u8p(vis), _lib.u8p(bnd), C.c_void_p(stream) if stream else None))
    return bool(hp.value), vis.astype(np.float32), bnd.astype(np.float32)
