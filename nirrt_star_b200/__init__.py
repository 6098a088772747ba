"""nirrt_star_b200 -- B200-native (sm_100a) implementation of the NIRRT* per-iteration hot path.

Scope: RRT*/IRRT*/NIRRT* Nearest/Near scans, segment-vs-obstacle collision
checks, cost walks / ChooseParent / Rewire, and PointNet++ guidance inference, behind the
reference's Python planner/wrapper API.  All compute runs in hand-written CUDA kernels reached
through the C-ABI library ``libnirrt_b200.so`` (include/nirrt_b200.h); there is no CPU fallback.
"""

__version__ = "0.1.0"
