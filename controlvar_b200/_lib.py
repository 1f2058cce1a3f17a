"""ctypes binding of libcvar_sm100.so (the C ABI declared in include/cvar.h).

There is deliberately no fallback: if the shared library is missing or does not export a symbol the import of the
compute path fails loudly (build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C controlvar_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcvar_sm100.so")

c_f = C.c_void_p      # device pointers travel as void*
c_ll = C.c_longlong


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", c_ll), ("strideA", c_ll),
        ("A_lo", C.c_void_p),
        ("W", C.c_void_p), ("ldw", c_ll), ("strideW", c_ll), ("w_is_kn", C.c_int),
        ("W_hi", C.c_void_p), ("W_lo", C.c_void_p),
        ("bias", C.c_void_p),
        ("out", C.c_void_p), ("ldo", c_ll), ("strideO", c_ll),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int),
        ("epilogue", C.c_int), ("alpha", C.c_float),
        ("gamma", C.c_void_p), ("gamma_row_stride", c_ll), ("rows_per_sample", C.c_int),
        ("resid", C.c_void_p), ("ldr", c_ll), ("strideR", c_ll),
        ("out_lo", C.c_void_p),
        ("A16_hi", C.c_void_p), ("A16_lo", C.c_void_p),
        ("W16_hi", C.c_void_p), ("W16_lo", C.c_void_p),
        ("out16_hi", C.c_void_p), ("out16_lo", C.c_void_p),
    ]


class ConvArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("bias", C.c_void_p),
        ("out", C.c_void_p),
        ("in_a", C.c_void_p), ("in_b", C.c_void_p), ("in_silu", C.c_int),
        ("resid", C.c_void_p),
        ("B", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int), ("ks", C.c_int),
        ("upsample2x", C.c_int),
        ("out_mode", C.c_int), ("out_rows_total", C.c_int), ("row_offset", C.c_int),
        ("engine", C.c_int),
        ("x16_hi", C.c_void_p), ("x16_lo", C.c_void_p), ("w16_hi", C.c_void_p), ("w16_lo", C.c_void_p),
        ("downsample2x", C.c_int), ("ksplit", C.c_int), ("out_samples", C.c_int),
        ("gn_part", C.c_void_p), ("gn_groups", C.c_int),
    ]


# name -> (restype, argtypes); must list every symbol of include/cvar.h (tests/test_abi.py cross-checks the header)
PROTOTYPES = {
    "cvar_abi_version": (C.c_int, []),
    "cvar_last_error": (C.c_char_p, []),
    "cvar_launch_count": (c_ll, []),
    "cvar_set_gemm_engine": (C.c_int, [C.c_int]),
    "cvar_get_gemm_engine": (C.c_int, []),
    "cvar_set_epilogue_overlap": (C.c_int, [C.c_int]),
    "cvar_add_launch_count": (C.c_int, [c_ll]),
    "cvar_set_fast_mode": (C.c_int, [C.c_int]),
    "cvar_get_fast_mode": (C.c_int, []),
    "cvar_set_tc_kblock": (C.c_int, [C.c_int]),
    "cvar_debug_set_trace": (C.c_int, [C.c_void_p]),
    "cvar_debug_set_attn_trace": (C.c_int, [C.c_void_p]),
    "cvar_lvl_pos": (C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_void_p]),
    "cvar_prologue": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, c_f, c_f, c_f, C.c_void_p]),
    "cvar_prologue_rows": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, c_f, c_f, c_f, C.c_void_p]),
    "cvar_ln_modulate": (C.c_int, [c_f, c_f, c_f, c_ll, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "cvar_gemm": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "cvar_qkv_project": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f,
                                   C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, c_f, C.c_void_p]),
    "cvar_split_tf32": (C.c_int, [c_f, c_f, c_f, c_ll, C.c_void_p]),
    "cvar_split_f16": (C.c_int, [c_f, c_f, c_f, c_ll, C.c_void_p]),
    "cvar_attn_kvcache": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_float, C.c_int, C.c_void_p]),
    "cvar_qkv_project16": (C.c_int, [c_f] * 13 + [C.c_int] * 6 + [c_f, C.c_void_p]),
    "cvar_attn_kvcache16": (C.c_int, [c_f] * 9 + [C.c_int] * 5 + [C.c_float, C.c_int, C.c_void_p]),
    "cvar_attn_blockcausal16": (C.c_int, [c_f] * 9 + [C.c_int] * 4 + [C.c_float, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "cvar_cfg_sample": (C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double,
                                  C.c_void_p]),
    "cvar_cfg_sample_multi": (C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int,
                                        C.c_int, C.c_double, c_f, c_f, C.c_int, C.c_void_p]),
    "cvar_cfg_sample_masked": (C.c_int, [c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int,
                                         C.c_int, C.c_double, c_f, c_f, C.c_int, C.c_void_p]),
    "cvar_gumbel_embed": (C.c_int, [c_f, c_f, c_f, c_f, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_double, C.c_double,
                                    C.c_void_p]),
    "cvar_vq_step_ex": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvar_area_pool_nc": (C.c_int, [c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvar_nchw_to_nhwc_pad": (C.c_int, [c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvar_repack_conv_weight_pad": (C.c_int, [c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvar_vq_step": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_int, C.c_void_p]),
    "cvar_vq_nearest": (C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvar_nchw_to_nhwc": (C.c_int, [c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, c_ll, C.c_void_p]),
    "cvar_gn_chunks": (C.c_int, [C.c_int]),
    "cvar_gn_stats": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                C.c_void_p]),
    "cvar_conv2d": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "cvar_repack_conv_weight": (C.c_int, [c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvar_affine_nc": (C.c_int, [c_f, c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvar_conv2d_f16_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cvar_conv2d_gn_fusable": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cvar_gn_finalize_parts": (C.c_int, [c_f, c_f, c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "cvar_upsample2x_split_f16": (C.c_int, [c_f, c_f, c_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cvar_softmax_rows": (C.c_int, [c_f, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


class CvarError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the library once; raise (never fall back) when it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CvarError(
            f"{LIB_PATH} is missing: the sm_100a extension has not been built. "
            "Run `make -C controlvar_b200/csrc` (or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise CvarError(f"{LIB_PATH} does not export {name}; rebuild the extension") from e
        fn.restype = res
        fn.argtypes = args
    if lib.cvar_abi_version() != 1:
        raise CvarError(f"ABI version mismatch: library reports {lib.cvar_abi_version()}, binding expects 1")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cvar_last_error()
        raise CvarError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
