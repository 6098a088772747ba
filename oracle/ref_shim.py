"""Import shim for the upstream reference (TEST INFRASTRUCTURE ONLY).

Makes the pure-Python reference under /root/reference importable in THIS container so
that golden vectors can be generated from the reference's own code
(tests/golden/make_golden_*.py).  Never imported by the product path, and never used on
the GPU box (/root/reference does not exist there).

Shims:
  * matplotlib / mpl_toolkits are absent -> empty stub modules (the planners import the
    visualisers at module top, rrt_visualizer_3d.py:5-6, but never draw unless asked).
  * open3d is absent -> stub exposing geometry.PointCloud.farthest_point_down_sample and
    utility.Vector3dVector (datasets_3d/point_cloud_mask_utils_3d.py:2,49-52).
  * top-level `datasets` is a namespace dir shadowed by the HuggingFace package ->
    pre-register module objects pointing at the reference directories.
"""
import sys
import types

import numpy as np

REF_ROOT = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _fps_first_index_zero(points, num_samples):
    """Open3D 0.17 PointCloud::FarthestPointDownSample as recalled (start at index 0,
    running min of squared distance in f64, first argmax).  Parity unpinned."""
    pts = np.asarray(points, dtype=np.float64)
    n = pts.shape[0]
    sel = np.zeros(num_samples, dtype=np.int64)
    dist = np.full(n, np.inf)
    far = 0
    for i in range(num_samples):
        sel[i] = far
        d = ((pts - pts[far]) ** 2).sum(axis=1)
        dist = np.minimum(dist, d)
        far = int(np.argmax(dist))
    return sel


class _O3DPointCloud:
    def __init__(self):
        self.points = None

    def farthest_point_down_sample(self, num_samples):
        out = _O3DPointCloud()
        pts = np.asarray(self.points)
        out.points = pts[_fps_first_index_zero(pts, num_samples)]
        return out


def install():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot", __getattr__=lambda name: _Anything())
        mpl.patches = _stub("matplotlib.patches", __getattr__=lambda name: _Anything())
        tk = _stub("mpl_toolkits")
        tk.mplot3d = _stub("mpl_toolkits.mplot3d", __getattr__=lambda name: _Anything())
        tk.mplot3d.art3d = _stub("mpl_toolkits.mplot3d.art3d", Poly3DCollection=_Anything)
    if "open3d" not in sys.modules:
        o3d = _stub("open3d")
        o3d.geometry = _stub("open3d.geometry", PointCloud=_O3DPointCloud)
        o3d.utility = _stub("open3d.utility", Vector3dVector=lambda a: np.asarray(a))
    for pkg in ("datasets", "datasets_3d"):
        m = types.ModuleType(pkg)
        m.__path__ = [f"{REF_ROOT}/{pkg}"]
        sys.modules[pkg] = m
