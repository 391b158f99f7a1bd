"""ctypes binding of libhudiff_b200.so (the C-ABI declared in include/hudiff.h).

The library is built in-tree by `hudiff_b200/csrc/build.sh` (or `__graft_entry__.build()`).
Loading fails loudly if it is missing: there is no Python/CPU fallback for this path.
"""
from __future__ import annotations

import ctypes as C
import os

_LIB = None
LIB_PATH = os.environ.get("HUDIFF_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhudiff_b200.so")

HUDIFF_AB, HUDIFF_NB = 0, 1
HUDIFF_FP16, HUDIFF_FP16X3 = 0, 1
ACT = {"relu": 1, "gelu": 2}

EXPORTS = [
    "hudiff_create", "hudiff_load_weights", "hudiff_workspace_bytes", "hudiff_dropout_words",
    "hudiff_forward", "hudiff_sample", "hudiff_sample_host", "hudiff_check_error", "hudiff_last_launch_count",
    "hudiff_set_gemm_impl", "hudiff_profile_step", "hudiff_debug_buffer", "hudiff_debug_attn_clocks", "hudiff_debug_attention", "hudiff_debug_gemm_clocks", "hudiff_destroy", "hudiff_last_error", "hudiff_version",
]


class HudiffCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "kind", "n_tokens", "d_model", "n_encoder_layers", "aa_kernel_size", "r", "n_side",
        "s_embedding", "n_region", "r_embedding", "max_len", "sum_d_model", "dual_layers",
        "att_model", "dim_feedforward", "nhead", "cs_layers", "activation", "precision")] + [
        ("dropout", C.c_double)]


class HudiffTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4), ("on_device", C.c_int32)]


class HudiffRng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("dropout", C.c_int32), ("seed", C.c_uint64),
                ("dropout_bits", C.c_void_p), ("exp_draws", C.c_void_p), ("seq_offset", C.c_uint64)]


class HudiffError(RuntimeError):
    pass


def load():
    """dlopen the library and declare the prototypes; raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise HudiffError(
            f"{LIB_PATH} not found: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()' or hudiff_b200/csrc/build.sh). "
            "hudiff_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32p = C.c_void_p, C.c_void_p
    lib.hudiff_create.argtypes = [C.POINTER(vp), C.POINTER(HudiffCfg), C.c_int]
    lib.hudiff_create.restype = C.c_int
    lib.hudiff_load_weights.argtypes = [vp, C.POINTER(HudiffTensor), C.c_int]
    lib.hudiff_load_weights.restype = C.c_int
    lib.hudiff_workspace_bytes.argtypes = [vp, C.c_int]
    lib.hudiff_workspace_bytes.restype = C.c_size_t
    lib.hudiff_dropout_words.argtypes = [vp, C.c_int]
    lib.hudiff_dropout_words.restype = C.c_size_t
    lib.hudiff_forward.argtypes = [vp, i32p, i32p, i32p, C.c_int, C.POINTER(HudiffRng), vp, vp]
    lib.hudiff_forward.restype = C.c_int
    lib.hudiff_sample.argtypes = [vp, i32p, i32p, i32p, i32p, C.c_int, C.c_int, C.POINTER(HudiffRng), vp, vp]
    lib.hudiff_sample.restype = C.c_int
    lib.hudiff_sample_host.argtypes = [vp, i32p, i32p, i32p, i32p, C.c_int, C.c_int, C.POINTER(HudiffRng), vp]
    lib.hudiff_sample_host.restype = C.c_int
    lib.hudiff_check_error.argtypes = [vp]
    lib.hudiff_check_error.restype = C.c_int
    lib.hudiff_last_launch_count.argtypes = [vp]
    lib.hudiff_last_launch_count.restype = C.c_longlong
    lib.hudiff_set_gemm_impl.argtypes = [vp, C.c_int]
    lib.hudiff_set_gemm_impl.restype = C.c_int
    lib.hudiff_profile_step.argtypes = [vp, i32p, i32p, i32p, C.c_int, C.POINTER(HudiffRng), C.POINTER(C.c_double),
                                        C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
    lib.hudiff_profile_step.restype = C.c_int
    lib.hudiff_debug_buffer.argtypes = [vp, C.c_int, C.c_char_p, C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.hudiff_debug_buffer.restype = C.c_int
    lib.hudiff_debug_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(C.c_float)]
    lib.hudiff_debug_attention.restype = C.c_int
    lib.hudiff_debug_attn_clocks.argtypes = [C.POINTER(C.c_longlong)]
    lib.hudiff_debug_attn_clocks.restype = C.c_int
    lib.hudiff_debug_gemm_clocks.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    lib.hudiff_debug_gemm_clocks.restype = C.c_int
    lib.hudiff_destroy.argtypes = [vp]
    lib.hudiff_destroy.restype = C.c_int
    lib.hudiff_last_error.restype = C.c_char_p
    lib.hudiff_version.restype = C.c_char_p
    _LIB = lib
    return lib


HUDIFF_ERR_INDEX = -5


def check(status: int, what: str):
    if status != 0:
        msg = load().hudiff_last_error().decode("utf-8", "replace")
        if status == HUDIFF_ERR_INDEX:          # the reference raises IndexError from its embedding lookups / indexing
            raise IndexError(f"{what}: {msg}")
        raise HudiffError(f"{what} failed (status {status}): {msg}")
