"""ctypes binding of librmnet_b200.so (the C ABI declared in include/rmnet_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librmnet_b200.so")
CSRC = os.path.join(_HERE, "csrc")

RMNET_PREC_SPLIT3, RMNET_PREC_SINGLE, RMNET_PREC_MIXED = 0, 1, 2
RMNET_IMPL_AUTO, RMNET_IMPL_SIMT, RMNET_IMPL_UMMA = 0, 1, 2
ELEM_BF16, ELEM_FP16 = 0, 1
SAMPLER_CUDNN, SAMPLER_ATEN = 0, 1
CH_KEEP, CH_ABSENT, CH_NEW = 0, 1, 2

_lock = threading.Lock()
_lib = None

c_int, c_float, c_size_t, c_void_p, c_ll = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/rmnet_b200.h one to one
PROTOTYPES = {
    "rmnet_abi_version": (c_int, []),
    "rmnet_last_error": (ctypes.c_char_p, []),
    "rmnet_launch_count": (c_ll, []),
    "rmnet_launch_count_reset": (None, []),
    "rmnet_has_umma": (c_int, []),
    "rmnet_reg_att_map_workspace_bytes": (c_size_t, [c_int, c_int]),
    "rmnet_reg_att_map_forward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "rmnet_warp_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rmnet_warp_att_map_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rmnet_regional_boxes_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                                             c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rmnet_frame_regions_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                                            c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                            c_void_p]),
    "rmnet_cell_rects_from_bboxes": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "rmnet_bank_bytes": (c_size_t, [c_int, c_int]),
    "rmnet_bank_reset": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "rmnet_bank_memorize": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll,
                                    c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "rmnet_bank_stats_host": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rmnet_memory_read_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "rmnet_bank_memory_read": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p, c_ll, c_void_p, c_int,
                                       c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rmnet_memory_read_plan_host": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "rmnet_frame_step": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "rmnet_mask_epilogue_forward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p]),
    "rmnet_memory_reader_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "rmnet_memory_reader_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                            c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "rmnet_update_optical_flow": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "rmnet_update_optical_flow_cpu": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "rmnet_update_optical_flow_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                               c_size_t, c_void_p]),
}


def build(verbose=False):
    """Compile every CUDA source for sm_100a into rmnet_b200/librmnet_b200.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        raise RuntimeError("building librmnet_b200.so failed:\n" + out.stdout)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(rmnet_b200 has no CPU or PyTorch fallback)")
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in PROTOTYPES.items():
                    fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
                    fn.restype = res
                    fn.argtypes = args
                if L.rmnet_abi_version() != 1:
                    raise RuntimeError("librmnet_b200.so ABI version mismatch")
                _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().rmnet_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"librmnet_b200 {what} failed (code {rc}): {msg}")
