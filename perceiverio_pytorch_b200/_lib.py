"""ctypes binding of libpio_b200.so (the C ABI declared in include/pio_b200.h).

The product path never falls back: if the library is missing it is built with nvcc (build.py); if that fails,
or a call returns a non-zero status, a RuntimeError carrying pio_last_error() is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libpio_b200.so")

EXPORTED_SYMBOLS = (
    "pio_abi_version", "pio_last_error", "pio_check_device", "pio_launch_count",
    "pio_profile_enable", "pio_profile_read",
    "pio_layernorm_bf16", "pio_gemm_bf16", "pio_softmax_bf16",
    "pio_attention_fwd", "pio_attention_supported", "pio_attention_key_tile", "pio_attention_combine",
    "pio_linear_f32", "pio_layernorm_concat_bf16", "pio_hash_words",
    "pio_decoder_attention_fwd", "pio_decoder_attention_supported", "pio_gemm_stats_parts", "pio_gemm_pair_kernel",
)

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class LayerNormArgs(C.Structure):
    _fields_ = [("x", vp), ("ldx", i64), ("y", vp), ("ldy", i64), ("gamma", vp), ("beta", vp),
                ("rows", i64), ("C", i32), ("normalize", i32), ("eps", f32), ("split", i32), ("fp16", i32)]


class LayerNormConcatArgs(C.Structure):
    _fields_ = [("feat", vp), ("feat_stride_b", i64), ("feat_stride_n", i64), ("feat_stride_c", i64),
                ("pos", vp), ("y", vp), ("ldy", i64), ("gamma", vp), ("beta", vp),
                ("B", i32), ("N", i32), ("Cf", i32), ("Cp", i32), ("eps", f32), ("fp16", i32)]


class GemmArgs(C.Structure):
    _fields_ = [("A", vp), ("lda", i64), ("strideA", i64),
                ("B", vp), ("ldb", i64), ("strideB", i64),
                ("b_mn_major", i32),
                ("M", i32), ("N", i32), ("K", i32), ("batch", i32),
                ("bias", vp), ("bias_mode", i32),
                ("act", i32),
                ("alpha", f32),
                ("residual", vp), ("ldr", i64), ("strideR", i64),
                ("out_f32", vp), ("ldo32", i64), ("strideO32", i64),
                ("out_bf16", vp), ("ldo16", i64), ("strideO16", i64),
                ("tile_n", i32), ("max_ctas", i32), ("cluster_m", i32), ("kernel", i32),
                ("row_stats_out", vp), ("row_stats_in", vp), ("ln_colsum", vp), ("ln_channels", i32), ("ln_eps", f32),
                ("reverse_tiles", i32), ("row_stats_parts", i32), ("fp16", i32),
                ("out_lo16", vp), ("residual_hi16", vp), ("residual_lo16", vp), ("ldr16", i64)]


class SoftmaxArgs(C.Structure):
    _fields_ = [("S", vp), ("lds", i64), ("strideS", i64),
                ("P", vp), ("ldp", i64), ("strideP", i64),
                ("key_mask", vp), ("stride_km", i64),
                ("row_keep", vp), ("stride_rk", i64),
                ("batch", i32), ("rows", i32), ("cols", i32),
                ("scale", f32), ("split", i32),
                ("dense_mask", vp), ("dm_stride_b", i64), ("dm_stride_r", i64),
                ("bias", vp), ("bias_stride_b", i64), ("bias_stride_r", i64), ("bias_stride_c", i64),
                ("P_f32", vp), ("ldpf", i64), ("stridePf", i64), ("fp16", i32)]


class AttentionArgs(C.Structure):
    _fields_ = [("Q", vp), ("ldq", i64), ("strideQ", i64),
                ("K", vp), ("ldk", i64), ("strideK", i64),
                ("V", vp), ("ldv", i64), ("strideV", i64),
                ("B", i32), ("H", i32), ("Nq", i32), ("Nk", i32), ("dqk", i32), ("dv", i32),
                ("scale", f32),
                ("key_mask", vp), ("stride_km", i64),
                ("row_keep", vp), ("stride_rk", i64),
                ("O", vp), ("ldo", i64), ("strideO", i64),
                ("num_splits", i32), ("partial", i32),
                ("O_part", vp), ("m_part", vp), ("l_part", vp), ("fp16", i32)]


class DecoderAttentionArgs(C.Structure):
    _fields_ = [("Q", vp), ("ldq", i64), ("strideQ", i64),
                ("K", vp), ("ldk", i64), ("strideK", i64),
                ("V", vp), ("ldv", i64), ("strideV", i64),
                ("B", i32), ("Nq", i32), ("Nk", i32), ("dqk", i32), ("dv", i32),
                ("scale", f32),
                ("key_mask", vp), ("stride_km", i64),
                ("row_keep", vp), ("stride_rk", i64),
                ("bias", vp),
                ("residual", vp), ("ldr", i64), ("strideR", i64),
                ("out", vp), ("ldo", i64), ("strideO", i64),
                ("fp16", i32),
                ("out_ln", vp), ("ld_ln", i64), ("stride_ln", i64), ("ln_gamma", vp), ("ln_beta", vp), ("ln_eps", f32)]


class CombineArgs(C.Structure):
    _fields_ = [("O_part", vp), ("m_part", vp), ("l_part", vp),
                ("part_stride_O", i64), ("part_stride_ml", i64),
                ("parts", i32), ("B", i32), ("H", i32), ("Nq", i32), ("dv", i32),
                ("row_keep", vp), ("stride_rk", i64),
                ("O", vp), ("ldo", i64), ("strideO", i64),
                ("O_out_part", vp), ("m_out", vp), ("l_out", vp), ("part_ptrs", vp), ("fp16", i32),
                ("row_alive", vp), ("stride_ra", i64)]


class LinearF32Args(C.Structure):
    _fields_ = [("x", vp), ("ldx", i64), ("w", vp), ("ldw", i64), ("bias", vp), ("y", vp), ("ldy", i64),
                ("M", i64), ("N", i32), ("K", i32),
                ("x2", vp), ("ldx2", i64), ("w2", vp), ("ldw2", i64), ("K2", i32), ("x2_fp16", i32)]


_lib = None
_lock = threading.Lock()


def load(build_if_missing: bool = True):
    """Load (building first if needed) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise RuntimeError(f"{LIB_PATH} is missing; run `python -m perceiverio_pytorch_b200.build`")
            from . import build as _build
            _build.build()
        lib = C.CDLL(LIB_PATH)
        lib.pio_abi_version.restype = C.c_int
        lib.pio_last_error.restype = C.c_char_p
        lib.pio_check_device.restype = C.c_int
        lib.pio_launch_count.restype = C.c_int64
        for name, argt in (("pio_layernorm_bf16", LayerNormArgs), ("pio_gemm_bf16", GemmArgs),
                           ("pio_softmax_bf16", SoftmaxArgs), ("pio_attention_fwd", AttentionArgs),
                           ("pio_attention_combine", CombineArgs), ("pio_linear_f32", LinearF32Args),
                           ("pio_layernorm_concat_bf16", LayerNormConcatArgs),
                           ("pio_decoder_attention_fwd", DecoderAttentionArgs)):
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(argt), C.c_void_p]
        lib.pio_hash_words.restype = C.c_int
        lib.pio_hash_words.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p]
        lib.pio_gemm_stats_parts.restype = C.c_int
        lib.pio_gemm_stats_parts.argtypes = [C.c_int32, C.c_int32]
        lib.pio_gemm_pair_kernel.restype = C.c_int
        lib.pio_gemm_pair_kernel.argtypes = [C.c_int32, C.c_int32]
        lib.pio_decoder_attention_supported.restype = C.c_int
        lib.pio_decoder_attention_supported.argtypes = [C.c_int32, C.c_int32]
        lib.pio_attention_supported.restype = C.c_int
        lib.pio_attention_supported.argtypes = [C.c_int32, C.c_int32]
        lib.pio_attention_key_tile.restype = C.c_int
        lib.pio_attention_key_tile.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        lib.pio_profile_enable.restype = None
        lib.pio_profile_enable.argtypes = [C.c_int]
        lib.pio_profile_read.restype = C.c_int
        lib.pio_profile_read.argtypes = [C.POINTER(C.c_double), C.c_int]
        if lib.pio_abi_version() != 15:
            raise RuntimeError("libpio_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().pio_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(load().pio_launch_count())


KERNEL_FAMILIES = ("layernorm", "gemm", "softmax", "attention", "combine", "linear_f32")


def profile_enable(on: bool) -> None:
    load().pio_profile_enable(1 if on else 0)


def profile_read() -> dict:
    """Drain the library-side launch records: {family: dict(ms, flops, bytes, launches)}."""
    n = len(KERNEL_FAMILIES)
    buf = (C.c_double * (4 * n))()
    check(load().pio_profile_read(buf, n), "pio_profile_read")
    return {f: dict(ms=buf[4 * i], flops=buf[4 * i + 1], bytes=buf[4 * i + 2], launches=int(buf[4 * i + 3]))
            for i, f in enumerate(KERNEL_FAMILIES)}
