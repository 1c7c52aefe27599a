"""Tensor-level wrappers over the C ABI (include/pio_b200.h).  PyTorch is used for device memory and streams only;
all arithmetic on the hot path happens inside libpio_b200.so.  Every function raises if the tensors are not on a
CUDA device — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib

BF16 = torch.bfloat16
# 16-bit operand / activation format of the launches issued from here on: bf16 (default) or fp16 (pio_*_args.fp16 in
# include/pio_b200.h).  Set through engine.set_precision() / engine.precision_scope(); every 16-bit buffer a block
# allocates and every launch it issues use the format current at that moment.
FP16 = False


def dtype16() -> torch.dtype:
    return torch.float16 if FP16 else BF16


def _f16() -> int:
    return 1 if FP16 else 0

GEMM_CLUSTER_M = 0   # 0 = library default; tests / bench can force 1, 2 or 4
GEMM_KERNEL = 0      # 0 = auto, 1 = single-CTA kernel, 2 = CTA-pair (cta_group::2) kernel


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def empty_f32_rows(m: int, n: int, device) -> torch.Tensor:
    """fp32 [m, n] whose row pitch is a multiple of 4 floats (16 bytes): rows of 261 / 322 / 1026 channels then satisfy
    the TMA stride rule, so GEMMs writing / reading them as output or residual can use the CTA-pair kernel's TMA
    epilogue.  Returns a (possibly non-contiguous) [m, n] view; the pad columns are never read."""
    n4 = (n + 3) // 4 * 4
    t = torch.empty((m, n4), dtype=torch.float32, device=device)
    return t if n4 == n else t[:, :n]


def stats_parts(m: int, n: int) -> int:
    """Partial (sum, sum of squares) slots per row that the fused-LayerNorm producer GEMM of shape m x n writes with the
    automatic kernel / tile choice: two per column tile (pio_gemm_stats_parts)."""
    return int(_lib.load().pio_gemm_stats_parts(m, n))


def gemm_uses_pair_kernel(m: int, n: int) -> bool:
    """Whether pio_gemm_bf16 runs an M x N problem (batch 1, default tuning fields) on the CTA-pair kernel."""
    return _lib.load().pio_gemm_pair_kernel(m, n) == 1


def empty_row_stats(m: int, n: int, device, parts: int = 0) -> torch.Tensor:
    """[m, parts, 2] fp32 buffer for pio_gemm_args.row_stats_out (every slot is written: no zeroing); parts defaults to
    stats_parts(m, n) — pass 2 * ceil(n / T) when the GEMM is forced to a tile width T."""
    return torch.empty((m, parts or stats_parts(m, n), 2), dtype=torch.float32, device=device)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("perceiverio_pytorch_b200: tensors must live on a CUDA (sm_100) device; "
                               "there is no CPU path")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"perceiverio_pytorch_b200: tensors of one call live on different devices "
                               f"({dev} and {t.device})")


def _device_guard(fn):
    """Run the wrapped entry point with the tensors' device as the current CUDA device: the library launches on the
    current device (stream, SM count, per-device kernel attributes), so a module moved with .to("cuda:1") must not
    depend on what torch.cuda.current_device() happens to be."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


@_device_guard
def layernorm_bf16(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], *,
                   normalize: bool = True, eps: float = 1e-5, out: Optional[torch.Tensor] = None,
                   split: int = 0, ld: int = 0) -> torch.Tensor:
    """x fp32 [..., C] (last dim contiguous, uniform row stride) -> bf16 [rows, pad8(C)] (or [rows, ld], ld >= pad8(C) a
    multiple of 8: columns C .. ld-1 are zeros).
    split = 1 / 2 (validation precision): bf16 [rows, 3 * pad8(C)] laid out [hi | lo | hi] / [hi | hi | lo]."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    c = x.shape[-1]
    x2 = x.reshape(-1, c) if x.dim() != 2 else x
    rows = x2.shape[0]
    assert not (ld and split) and (ld == 0 or (ld % 8 == 0 and ld >= pad8(c)))
    ldy = ld if ld else pad8(c) * (3 if split else 1)
    if out is None:
        out = torch.empty((rows, ldy), dtype=BF16 if split else dtype16(), device=x.device)
    assert out.dtype == (BF16 if split else dtype16()) and out.shape[-1] == ldy and out.is_contiguous()
    a = _lib.LayerNormArgs(_ptr(x2), x2.stride(0), _ptr(out), ldy, _ptr(gamma), _ptr(beta), rows, c,
                           1 if normalize else 0, eps, split, 0 if split else _f16())
    _lib.check(_lib.load().pio_layernorm_bf16(C.byref(a), _stream()), "pio_layernorm_bf16")
    return out


def layernorm_concat_supported(B: int, N: int, Cf: int, Cp: int) -> bool:
    """Shapes pio_layernorm_concat_bf16 covers (include/pio_b200.h)."""
    if N % 4 != 0 or Cf <= 0 or Cp <= 0 or Cf + Cp > 1208:
        return False
    r = 16
    while r > 4 and (N % r != 0 or r * Cp * 4 > 17 * 1024 or B * r * Cf * 4 > 48 * 1024):
        r -= 4
    return N % r == 0 and B * r * Cf * 4 <= 48 * 1024 and r * Cp * 4 <= 40 * 1024


@_device_guard
def layernorm_concat_bf16(feat: torch.Tensor, pos: torch.Tensor, gamma: Optional[torch.Tensor],
                          beta: Optional[torch.Tensor], *, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm(cat([feat [B, N, Cf] (any strides), pos [N, Cp] broadcast over the batch], -1)) -> bf16
    [B * N, pad8(Cf + Cp)] without materialising the concatenated array."""
    _need_cuda(feat, pos, gamma, beta)
    assert feat.dtype == torch.float32 and pos.dtype == torch.float32 and feat.dim() == 3 and pos.dim() == 2
    B, N, Cf = feat.shape
    assert pos.shape[0] == N and pos.is_contiguous()
    Cp = pos.shape[1]
    ldy = pad8(Cf + Cp)
    out = torch.empty((B * N, ldy), dtype=dtype16(), device=feat.device)
    a = _lib.LayerNormConcatArgs(_ptr(feat), feat.stride(0), feat.stride(1), feat.stride(2), _ptr(pos), _ptr(out), ldy,
                                 _ptr(gamma), _ptr(beta), B, N, Cf, Cp, eps, _f16())
    _lib.check(_lib.load().pio_layernorm_concat_bf16(C.byref(a), _stream()), "pio_layernorm_concat_bf16")
    return out


@_device_guard
def gemm(A: torch.Tensor, B: torch.Tensor, *, M: int, N: int, K: int, batch: int = 1, b_mn_major: bool = False,
         strideA: int = 0, strideB: int = 0, lda: Optional[int] = None, ldb: Optional[int] = None,
         bias: Optional[torch.Tensor] = None, bias_mode: int = 1, act: int = 0, alpha: float = 1.0,
         residual: Optional[torch.Tensor] = None, ldr: int = 0, strideR: int = 0,
         out_f32: Optional[torch.Tensor] = None, ldo32: int = 0, strideO32: int = 0,
         out_bf16: Optional[torch.Tensor] = None, ldo16: int = 0, strideO16: int = 0,
         tile_n: int = 0, max_ctas: int = 0, cluster_m: Optional[int] = None, kernel: Optional[int] = None,
         row_stats_out: Optional[torch.Tensor] = None, row_stats_in: Optional[torch.Tensor] = None,
         ln_colsum: Optional[torch.Tensor] = None, ln_channels: int = 0, ln_eps: float = 1e-5,
         reverse_tiles: bool = False, row_stats_parts: int = 0,
         out_lo16: Optional[torch.Tensor] = None, residual_hi16: Optional[torch.Tensor] = None,
         residual_lo16: Optional[torch.Tensor] = None, ldr16: int = 0) -> None:
    """Raw batched GEMM + epilogue; see pio_gemm_args in include/pio_b200.h."""
    _need_cuda(A, B, bias, residual, out_f32, out_bf16, out_lo16, residual_hi16, residual_lo16)
    assert A.dtype == B.dtype and A.dtype in (BF16, torch.float16)
    lda = A.stride(-2) if lda is None else lda
    ldb = B.stride(-2) if ldb is None else ldb
    st = row_stats_out if row_stats_out is not None else row_stats_in
    if st is not None and not row_stats_parts:
        assert st.dim() == 3 and st.shape[2] == 2 and st.is_contiguous(), "row statistics are [M, parts, 2] fp32"
        row_stats_parts = st.shape[1]
    a = _lib.GemmArgs(_ptr(A), lda, strideA, _ptr(B), ldb, strideB, 1 if b_mn_major else 0,
                      M, N, K, batch, _ptr(bias), bias_mode if bias is not None else 0, act, alpha,
                      _ptr(residual), ldr, strideR, _ptr(out_f32), ldo32, strideO32,
                      _ptr(out_bf16), ldo16, strideO16, tile_n, max_ctas,
                      GEMM_CLUSTER_M if cluster_m is None else cluster_m,
                      GEMM_KERNEL if kernel is None else kernel,
                      _ptr(row_stats_out), _ptr(row_stats_in), _ptr(ln_colsum), ln_channels, ln_eps,
                      1 if reverse_tiles else 0, row_stats_parts, 1 if A.dtype == torch.float16 else 0,
                      _ptr(out_lo16), _ptr(residual_hi16), _ptr(residual_lo16), ldr16)
    _lib.check(_lib.load().pio_gemm_bf16(C.byref(a), _stream()), "pio_gemm_bf16")


def linear(x: torch.Tensor, K: int, w: torch.Tensor, N: int, bias: Optional[torch.Tensor] = None, *,
           act: int = 0, alpha: float = 1.0, residual: Optional[torch.Tensor] = None,
           want_f32: bool = False, want_bf16: bool = True, bias_mode: int = 1):
    """y = act(alpha * x @ w^T + bias) (+ residual).  x bf16 [M, ldx], w bf16 [N, ldw] (nn.Linear layout).
    Returns (y_f32 [M, N] or None, y_bf16 [M, pad8(N)] or None)."""
    m = x.shape[0]
    y32 = empty_f32_rows(m, N, x.device) if want_f32 else None
    y16 = torch.empty((m, pad8(N)), dtype=dtype16(), device=x.device) if want_bf16 else None
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(-1) == 1
    gemm(x, w, M=m, N=N, K=K, bias=bias, bias_mode=bias_mode, act=act, alpha=alpha,
         residual=residual, ldr=residual.stride(0) if residual is not None else 0,
         out_f32=y32, ldo32=y32.stride(0) if y32 is not None else 0, out_bf16=y16, ldo16=pad8(N))
    return y32, y16


@_device_guard
def linear_f32(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], *, x2: Optional[torch.Tensor] = None,
               w2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x @ w^T + bias (+ x2 @ w2^T) in fp32 on CUDA cores, for N <= 16 output channels (x fp32 [M, K], w fp32 [N, K];
    x2 16-bit [M, >= K2], w2 fp32 [N, K2])."""
    _need_cuda(x, w, bias, x2, w2)
    assert x.dtype == torch.float32 and w.dtype == torch.float32 and x.stride(-1) == 1 and w.stride(-1) == 1
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty((m, n), dtype=torch.float32, device=x.device)
    if x2 is not None:
        assert x2.dtype in (BF16, torch.float16) and w2.dtype == torch.float32 and x2.shape[0] == m and w2.shape[0] == n
        extra = (_ptr(x2), x2.stride(0), _ptr(w2), w2.stride(0), w2.shape[1], 1 if x2.dtype == torch.float16 else 0)
    else:
        extra = (None, 0, None, 0, 0, 0)
    a = _lib.LinearF32Args(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(y), n, m, n, k, *extra)
    _lib.check(_lib.load().pio_linear_f32(C.byref(a), _stream()), "pio_linear_f32")
    return y


@_device_guard
def softmax_bf16(S: torch.Tensor, cols: int, scale: float, key_mask: Optional[torch.Tensor] = None,
                 row_keep: Optional[torch.Tensor] = None, split: bool = False, *,
                 dense_mask: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                 probs_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """S fp32 [B, rows, lds] -> P bf16 [B, rows, pad8(cols)] (split: [B, rows, 3 * pad8(cols)] as [hi | lo | hi]).

    General attention arguments: dense_mask u8 [B, rows, cols]; bias fp32 [B, rows, cols] view (any strides, 0 =
    broadcast), added to S before the scale; probs_out fp32 [B, rows, cols] view receiving the probabilities."""
    _need_cuda(S, key_mask, row_keep, dense_mask, bias, probs_out)
    b, rows, lds = S.shape
    ldp = pad8(cols) * (3 if split else 1)
    P = torch.empty((b, rows, ldp), dtype=BF16 if split else dtype16(), device=S.device)
    if dense_mask is not None:
        assert dense_mask.dtype == torch.uint8 and dense_mask.shape == (b, rows, cols) and dense_mask.stride(2) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.shape == (b, rows, cols)
    if probs_out is not None:
        assert probs_out.dtype == torch.float32 and probs_out.shape == (b, rows, cols) and probs_out.stride(2) == 1
    a = _lib.SoftmaxArgs(_ptr(S), lds, S.stride(0), _ptr(P), ldp, P.stride(0),
                         _ptr(key_mask), key_mask.stride(0) if key_mask is not None else 0,
                         _ptr(row_keep), row_keep.stride(0) if row_keep is not None else 0,
                         b, rows, cols, scale, 1 if split else 0,
                         _ptr(dense_mask), dense_mask.stride(0) if dense_mask is not None else 0,
                         dense_mask.stride(1) if dense_mask is not None else 0,
                         _ptr(bias), *(bias.stride() if bias is not None else (0, 0, 0)),
                         _ptr(probs_out), probs_out.stride(1) if probs_out is not None else 0,
                         probs_out.stride(0) if probs_out is not None else 0, 0 if split else _f16())
    _lib.check(_lib.load().pio_softmax_bf16(C.byref(a), _stream()), "pio_softmax_bf16")
    return P


def attention_supported(dqk: int, dv: int) -> bool:
    return _lib.load().pio_attention_supported(dqk, dv) == 0


@_device_guard
def attention_fwd(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, *, B: int, H: int, Nq: int, Nk: int,
                  dqk: int, dv: int, strideQ: int, strideK: int, strideV: int,
                  ldq: int, ldk: int, ldv: int, scale: Optional[float] = None,
                  key_mask: Optional[torch.Tensor] = None, row_keep: Optional[torch.Tensor] = None,
                  num_splits: int = 1, partial: bool = False, out: Optional[torch.Tensor] = None):
    """Streaming attention.  Returns O bf16 [B, Nq, pad8(H*dv)] or, if partial, (O_part, m_part, l_part)."""
    _need_cuda(Q, K, V, key_mask, row_keep)
    scale = 1.0 / math.sqrt(dqk) if scale is None else scale
    dev = Q.device
    ldo = pad8(H * dv)
    emit_partial = partial or num_splits > 1
    O = out
    if O is None and not partial:
        O = torch.empty((B, Nq, ldo), dtype=dtype16(), device=dev)
    Op = mp = lp = None
    if emit_partial:
        Op = torch.empty((num_splits, B, H, Nq, dv), dtype=torch.float32, device=dev)
        mp = torch.empty((num_splits, B, H, Nq), dtype=torch.float32, device=dev)
        lp = torch.empty((num_splits, B, H, Nq), dtype=torch.float32, device=dev)
    a = _lib.AttentionArgs(_ptr(Q), ldq, strideQ, _ptr(K), ldk, strideK, _ptr(V), ldv, strideV,
                           B, H, Nq, Nk, dqk, dv, scale,
                           _ptr(key_mask), key_mask.stride(0) if key_mask is not None else 0,
                           _ptr(row_keep), row_keep.stride(0) if row_keep is not None else 0,
                           _ptr(O), ldo, (O.stride(0) if O is not None else 0),
                           num_splits, 1 if partial else 0, _ptr(Op), _ptr(mp), _ptr(lp),
                           1 if Q.dtype == torch.float16 else 0)
    _lib.check(_lib.load().pio_attention_fwd(C.byref(a), _stream()), "pio_attention_fwd")
    if partial:
        return Op, mp, lp
    if num_splits > 1:
        attention_combine(Op, mp, lp, row_keep=row_keep, out=O)
    return O


def decoder_attention_supported(dqk: int, dv: int) -> bool:
    return _lib.load().pio_decoder_attention_supported(dqk, dv) == 0


@_device_guard
def decoder_attention(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, *, B: int, Nq: int, Nk: int, dqk: int, dv: int,
                      ldq: int, ldk: int, ldv: int, strideQ: int, strideK: int, strideV: int, scale: float,
                      key_mask: Optional[torch.Tensor] = None, row_keep: Optional[torch.Tensor] = None,
                      bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
                      ldr: int = 0, strideR: int = 0, ln=None):
    """Query-tiled single-head decoder attention (pio_decoder_attention_fwd): returns fp32 [B * Nq, dv] (row pitch a
    multiple of 4 floats) = softmax(scale Q K^T) V + bias (+ residual).  ln = (gamma, beta, eps): additionally returns
    LayerNorm(out) as 16-bit rows [B * Nq, pad8(dv)] (written by the kernel's epilogue) -> (out, out_ln)."""
    _need_cuda(Q, K, V, key_mask, row_keep, bias, residual)
    assert Q.dtype == K.dtype == V.dtype and Q.dtype in (BF16, torch.float16)
    out = empty_f32_rows(B * Nq, dv, Q.device)
    ldo = out.stride(0)
    out_ln = None
    ln_args = (None, 0, 0, None, None, 0.0)
    if ln is not None:
        _need_cuda(ln[0], ln[1])
        out_ln = torch.empty((B * Nq, pad8(dv)), dtype=Q.dtype, device=Q.device)
        ln_args = (_ptr(out_ln), out_ln.stride(0), Nq * out_ln.stride(0), _ptr(ln[0]), _ptr(ln[1]), float(ln[2]))
    a = _lib.DecoderAttentionArgs(_ptr(Q), ldq, strideQ, _ptr(K), ldk, strideK, _ptr(V), ldv, strideV,
                                  B, Nq, Nk, dqk, dv, scale,
                                  _ptr(key_mask), key_mask.stride(0) if key_mask is not None else 0,
                                  _ptr(row_keep), row_keep.stride(0) if row_keep is not None else 0,
                                  _ptr(bias), _ptr(residual), ldr, strideR, _ptr(out), ldo, Nq * ldo,
                                  1 if Q.dtype == torch.float16 else 0, *ln_args)
    _lib.check(_lib.load().pio_decoder_attention_fwd(C.byref(a), _stream()), "pio_decoder_attention_fwd")
    return out if ln is None else (out, out_ln)


@_device_guard
def attention_combine(Op: Optional[torch.Tensor], mp: Optional[torch.Tensor], lp: Optional[torch.Tensor], *,
                      row_keep: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                      part_stride_O: int = 0, part_stride_ml: int = 0, shape=None,
                      merged_out=None, normalised: bool = True, part_ptrs_dev: int = 0, device=None,
                      row_alive: Optional[torch.Tensor] = None):
    """Merge partials -> O bf16 [B, Nq, pad8(H*dv)] and/or a merged un-normalised partial.

    Op/mp/lp are [parts, B, H, Nq, dv] / [parts, B, H, Nq] tensors, or flat buffers with explicit part strides and
    `shape=(parts, B, H, Nq, dv)`.  `merged_out=(O, m, l)` (fp32 [B,H,Nq,dv] / [B,H,Nq]) receives the merged partial.
    `part_ptrs_dev` (an integer device address of an array of `parts` peer pointers, e.g. a symmetric-memory handle's
    buffer_ptrs_dev) selects the fused NVLink exchange: every part is loaded from its rank's packed partial."""
    _need_cuda(Op, mp, lp, row_keep)
    parts, b, h, nq, dv = Op.shape if shape is None else shape
    ldo = pad8(h * dv)
    if out is None and normalised:
        out = torch.empty((b, nq, ldo), dtype=dtype16(), device=Op.device if Op is not None else device)
    mo = merged_out if merged_out is not None else (None, None, None)
    a = _lib.CombineArgs(_ptr(Op), _ptr(mp), _ptr(lp), part_stride_O, part_stride_ml, parts, b, h, nq, dv,
                         _ptr(row_keep), row_keep.stride(0) if row_keep is not None else 0,
                         _ptr(out), ldo, out.stride(0) if out is not None else 0,
                         _ptr(mo[0]), _ptr(mo[1]), _ptr(mo[2]), C.c_void_p(part_ptrs_dev) if part_ptrs_dev else None,
                         1 if (out is not None and out.dtype == torch.float16) else 0,
                         _ptr(row_alive), row_alive.stride(0) if row_alive is not None else 0)
    _lib.check(_lib.load().pio_attention_combine(C.byref(a), _stream()), "pio_attention_combine")
    return out


def content_hash(*tensors: Optional[torch.Tensor]) -> tuple:
    """128-bit content hash of the given CUDA tensors (pio_hash_words; None entries and shapes are part of the key).
    One pass over the data and one 16-byte read-back (a host synchronisation)."""
    dev = next(t.device for t in tensors if t is not None)
    acc = torch.zeros(2, dtype=torch.int64, device=dev)
    meta = []
    with torch.cuda.device(dev):
        for i, t in enumerate(tensors):
            if t is None:
                meta.append(None)
                continue
            _need_cuda(t)
            meta.append((tuple(t.shape), str(t.dtype)))
            c = t.contiguous()
            nbytes = c.numel() * c.element_size()
            if nbytes % 4 != 0 or c.data_ptr() % 16 != 0:      # odd byte counts (bool masks): widen to whole words
                flat = torch.zeros((nbytes + 3) // 4 * 4, dtype=torch.uint8, device=dev)
                flat[:nbytes] = c.reshape(-1).view(torch.uint8)
                c, nbytes = flat, flat.numel()
            if nbytes:
                _lib.check(_lib.load().pio_hash_words(_ptr(c), nbytes // 4, (i + 1) << 40, _ptr(acc), _stream()),
                           "pio_hash_words")
    h = acc.tolist()
    return (h[0], h[1], tuple(meta))
