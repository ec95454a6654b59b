"""rmnet_b200 -- B200 (sm_100a) implementation of hzxie/RMNet's per-frame regional memory-read hot path.

Layout: csrc/ (CUDA kernels + the C ABI of include/rmnet_b200.h), _lib.py (ctypes binding), ops.py (tensor
plumbing), modules.py (mirror of the reference's operator interface), dropin/ (top-level modules named like the
reference's compiled extensions).  There is no CPU / PyTorch fallback: a missing library raises.
"""
from ._lib import (CH_ABSENT, CH_KEEP, CH_NEW, ELEM_BF16, ELEM_FP16, RMNET_IMPL_AUTO, RMNET_IMPL_SIMT, RMNET_IMPL_UMMA, RMNET_PREC_MIXED, RMNET_PREC_SINGLE,  # noqa: F401
                   RMNET_PREC_SPLIT3, build, lib)
from .modules import (MemoryReader, RegionalAttentionMapGenerator, RegionalAttentionMapGeneratorFunction,  # noqa: F401
                      RegionalMemory, fused_forward, get_att_map, install, uninstall, warp)
from .frame_loop import RegionalFrameLoop  # noqa: F401
from .ops import MemoryBank, mask_epilogue, update_optical_flow  # noqa: F401
