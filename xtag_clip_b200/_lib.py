"""ctypes binding of libxtag_b200.so (include/xtag_b200.h).  There is NO fallback: if the
library is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_size_t, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libxtag_b200.so")

XTAG_F32, XTAG_BF16 = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2
BWD_REUSE_DS = 1
ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = -1, -2, -3, -4

_lib = None

# name -> (restype, argtypes); mirrors include/xtag_b200.h one to one
SIGNATURES = {
    "xtag_version": (c_int, []),
    "xtag_last_error": (c_char_p, []),
    "xtag_device_check": (c_int, []),
    "xtag_launch_count": (c_uint64, []),
    "xtag_set_tune": (c_int, [c_int]),
    "xtag_get_tune": (c_int, []),
    "xtag_set_spin_timeout_ms": (c_int, [ctypes.c_longlong]),
    "xtag_prof_enable": (c_int, [c_int]),
    "xtag_prof_read": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "xtag_l2norm_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "xtag_l2norm_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                c_void_p]),
    "xtag_clip_fwd_ws_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "xtag_clip_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "xtag_clip_fwd_block_parts": (c_int, [c_int, c_int, c_void_p, c_void_p]),
    "xtag_clip_fwd_block": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                    c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "xtag_clip_fwd_stream": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                     c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "xtag_lse_reduce_log2": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "xtag_lse_reduce2_log2": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "xtag_lse_combine_loss_scratch_bytes": (c_size_t, []),
    "xtag_lse_combine_ptrs_loss": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "xtag_lse_combine": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "xtag_lse_combine_ptrs": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "xtag_sum_ptrs_bf16": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p]),
    "xtag_clip_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "xtag_clip_bwd_ws_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "xtag_clip_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                              c_void_p, c_void_p, c_float, c_float, c_float, c_void_p,
                              c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "xtag_tc_linear_bf16": (c_int, [c_void_p, ctypes.c_long, c_void_p, c_void_p, c_void_p, ctypes.c_long, c_int, c_int, c_int,
                                    c_void_p]),
    "xtag_xattn_bwd_ld": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                  c_void_p, c_void_p, c_void_p,
                                  c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_float, c_float, c_uint64, c_uint64, c_void_p, c_size_t, c_void_p]),
    "xtag_tc_gemm_nt": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "xtag_tc_gemm_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "xtag_tc_gemm_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p,
                                c_size_t, c_void_p]),
    "xtag_tc_gemm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p]),
    "xtag_xattn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                               c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_float, c_float, c_uint64, c_uint64, c_void_p]),
    "xtag_xattn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                               c_void_p, c_void_p, c_void_p,
                               c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_float, c_float, c_uint64, c_uint64, c_void_p, c_size_t, c_void_p]),
    "xtag_debug_tile_coords": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "xtag_debug_pick_cluster": (c_int, [c_int, c_int, c_int]),
    "xtag_debug_work_item": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "xtag_siglip_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_float,
                                c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "xtag_symm_ce_ws_bytes": (c_size_t, [c_int]),
    "xtag_symm_ce_fwd": (c_int, [c_void_p, c_int, c_int, ctypes.c_long, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "xtag_symm_ce_bwd": (c_int, [c_void_p, c_int, c_int, ctypes.c_long, c_void_p, c_void_p, c_void_p, c_void_p,
                                 ctypes.c_long, c_void_p]),
    "xtag_ln_res_bwd_ws_bytes": (c_size_t, [c_int, c_int]),
    "xtag_ln_res_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_int, c_float, c_float, c_uint64, c_uint64, c_void_p]),
    "xtag_ln_res_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_int, c_float, c_uint64, c_uint64, c_void_p, c_size_t, c_void_p]),
    "xtag_asl_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_float, c_float, c_float,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
}


def load() -> ctypes.CDLL:
    """Load (once) and type the shared library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `python -m xtag_clip_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().xtag_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


_replayed = 0


def note_replayed(n: int) -> None:
    """kernels of this library executed by a CUDA-graph replay (they were counted once, at capture)"""
    global _replayed
    _replayed += int(n)


def launch_count() -> int:
    """kernels of libxtag_b200 launched by this process: direct launches (counted inside the library) plus the
    library's kernel nodes of every CUDA-graph replay"""
    return int(load().xtag_launch_count()) + _replayed
