"""recommend-sys_b200 — B200-native (sm_100a) KNN hot path of Oneaccount1/recommend-sys.

Only what the path needs lives here:
  csrc/   hand-written CUDA kernels + the C ABI (include/rs_knn.h) -> librs_knn_b200.so
  host/   host code of the drop-in that stays on the CPU in the reference too -> librs_host.so
  core.py Python mirror of the reference's Go `core` API for this path, over ctypes
  go/     the cgo bridge a maintainer of the reference would add (cannot be compiled here)
The directory name contains a hyphen; `import recommend_sys_b200` (a shim package at the
repository root) loads it.
"""
from .core import *  # noqa: F401,F403
from . import core  # noqa: F401
from . import shard  # noqa: F401
