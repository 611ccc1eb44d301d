"""side_b200 -- B200 (sm_100a) implementation of SIDE's stereo hot path behind the reference's module API.

Layout:
  csrc/                 hand-written CUDA kernels + the C ABI (include/side_b200.h) -> libside_b200.so
  _lib.py, ops.py       ctypes binding, tensor-level operator wrappers (autograd Functions)
  dcn_v2.py             DCN / DCNv2 / dcn_v2_conv drop-ins (+ `_ext` shim)
  decode.py             bbox_decode / ddd_decode drop-ins
  networks/             DLA-34 + DCN neck, canonical stereo_network, cost_volume, get_pose_net
  engine.py             CUDA-graph inference engine + pair-sharded multi-GPU helpers
"""
from . import _lib  # noqa: F401

__all__ = ["ops", "dcn_v2", "decode", "networks", "engine"]
__version__ = "0.1.0"
