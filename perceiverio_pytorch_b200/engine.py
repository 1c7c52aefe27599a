"""Host-side lowering of the Perceiver IO blocks onto the C-ABI kernels.

This is orchestration only: LayerNorm+cast, GEMM(+epilogue), streaming attention, softmax and combine launches on
the caller's CUDA stream.  Activations between kernels are bf16 (MMA operands) while the residual stream, LayerNorm
statistics, softmax statistics and all accumulators stay fp32.

Weight preparation (bf16 copies, fused QKV matrices and the single-head "folded" products described in DESIGN.md)
happens once per module and is cached as non-persistent derived state keyed on the parameters' versions.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from . import ops
from .inputs import PositionedInput
from .ops import pad8

# Arithmetic mode:
#   "bf16"   (default) bf16 MMA operands, fp32 residual stream / statistics / accumulators;
#   "fp16"   the same kernels at the same speed with IEEE-half operands: 11 instead of 8 mantissa bits, i.e. 8x less
#            operand rounding, fp16's range (conversions saturate).  The optical-flow recipe needs it — with bf16 operands
#            the reference ALGORITHM itself misses the 1e-2 bound there (1.86e-2 in a CPU emulation;
#            2.7e-2 measured on tests/golden/full/flow.npz) — and it is what the reference's own `mixed_precision=True`
#            means (fp16 autocast, flow_perceiver.py:14,129);
#   "bf16x3" validation precision: bf16 x 2 split operands, three MMAs per product (validate.py).
PRECISION = "bf16"
MODES = ("bf16", "fp16", "bf16x3")


def set_precision(mode: str) -> None:
    global PRECISION
    if mode not in MODES:
        raise ValueError(f"unknown precision {mode!r}: use one of {MODES}")
    PRECISION = mode
    ops.FP16 = (mode == "fp16")


class precision_scope:
    """`with precision_scope("fp16"): ...` — run a block in another arithmetic mode (None: keep the current one).
    PerceiverEncoder / PerceiverDecoder use it for their per-module `precision` attribute."""

    def __init__(self, mode: Optional[str]):
        if mode is not None and mode not in MODES:
            raise ValueError(f"unknown precision {mode!r}: use one of {MODES}")
        self.mode = mode

    def __enter__(self):
        self.prev = PRECISION
        if self.mode is not None:
            set_precision(self.mode)

    def __exit__(self, *exc):
        if self.mode is not None:
            set_precision(self.prev)
        return False


def fast() -> bool:
    """True for the 16-bit-operand modes (everything except the validation precision)."""
    return PRECISION != "bf16x3"


# Tower LayerNorms folded into the projections around them when the latent array has at least
# FUSE_LN_MIN_ROWS rows, or — with fp16 operands — at least FUSE_LN_MIN_CHANNELS channels.  Both GEMM kernels carry the
# fused epilogues and the row statistics are deterministic.  For the batch-1 towers the fusion removes two of the seven
# launches of a layer: with 1280 channels (language, 256 rows) a LayerNorm launch costs 4.9 us of a 53 us layer and the
# forward goes 1.95 -> 1.66 ms; with 512 channels (flow 2048 rows, multimodal 784 rows) it costs 2.5 - 2.8 us, the GEMM
# epilogues grow by as much and nothing is gained (3.57 vs 3.55, 1.15 vs 1.17 ms per forward).  The fused form rounds x to
# 16 bits before the mean is subtracted; at full size the language output moves from 7.9e-3 to 8.8e-3 of the 1e-2 bound
# with bf16 operands (1.1e-3 -> 1.2e-3 with fp16), so the small towers take it only in the fp16 mode, where the margin is
# an order of magnitude; the large ones (>= 4096 rows: 7.0e-3 at B = 8, classification) take it in both.
FUSE_LN = os.environ.get("PIO_FUSE_LN", "1") != "0"
FUSE_LN_MIN_ROWS = int(os.environ.get("PIO_FUSE_LN_MIN_ROWS", "4096"))
FUSE_LN_MIN_CHANNELS = int(os.environ.get("PIO_FUSE_LN_MIN_CHANNELS", "1024"))
FUSE_LN_MAX_OFFSET = 1.0     # max |mean| / std of a residual-stream row the fused form accepts (perceiver.PerceiverEncoder)

# Residual stream of the fused tower as a pair of 16-bit arrays between its producer GEMMs (pio_gemm_args.out_lo16) when
# every one of them runs on the CTA-pair kernel (ops.gemm_uses_pair_kernel): 4 instead of 6 bytes written per element.
SPLIT_STREAM = os.environ.get("PIO_SPLIT_STREAM", "1") != "0"
STREAM_INPLACE = os.environ.get("PIO_STREAM_INPLACE", "1") != "0"   # ... updated in place from the second producer on
REVERSE_FC2 = os.environ.get("PIO_REVERSE_FC2", "1") != "0"
REVERSE_FC1 = os.environ.get("PIO_REVERSE_FC1", "1") != "0"   # fc1 too: the out-projection wrote its operand front to back

# Flags (module-level so tests / bench can flip them)
ENABLE_FOLDING = True      # single-head cross-attention: K == V == LN(x)
ENCODER_KEY_SPLITS = 0     # 0 = auto


def _bf16_weight(w: torch.Tensor) -> torch.Tensor:
    """fp32 [N, K] -> 16-bit [N, pad8(K)] in the current operand format (pad columns zero; never read by TMA anyway)."""
    n, k = w.shape
    out = torch.zeros((n, pad8(k)), dtype=ops.dtype16(), device=w.device)
    out[:, :k] = w.detach().to(ops.dtype16())
    return out


class PreparedAttention:
    """Derived bf16 weights of one reference `Attention` module (+ fused / folded variants)."""

    def __init__(self, att, *, self_attention: bool, allow_fold: bool):
        wq, bq = att.proj_q.weight.detach(), att.proj_q.bias.detach()
        wk, bk = att.proj_k.weight.detach(), att.proj_k.bias.detach()
        wv, bv = att.proj_v.weight.detach(), att.proj_v.bias.detach()
        wf = att.final.weight.detach()
        bf = att.final.bias.detach() if att.final.bias is not None else None
        self.H = att._num_heads
        self.QK, self.Cq = wq.shape
        self.Ck = wk.shape[1]
        self.V = wv.shape[0]
        self.O = wf.shape[0]
        self.dqk = self.QK // self.H
        self.dv = self.V // self.H
        self.scale = 1.0 / math.sqrt(self.dqk)
        self.folded = bool(allow_fold and ENABLE_FOLDING and self.H == 1 and wv.shape[1] == self.Ck
                           and ops.attention_supported(self.Ck, self.Ck) and self.Ck <= 384)
        # The same key-side fold for single heads the streaming kernels do not cover (the multimodal encoder: 704 channels),
        # on the explicit S / P path: S = Q' LN(x)^T and O' = P LN(x) take the normalised input array itself as their
        # operand — K-major for S, MN-major for P.V — so the two projections of the long input array (2 x 52 GF for the
        # multimodal encoder) disappear.  Chosen per call (use_key_fold_wide): it only pays when the keys far outnumber
        # the queries; never more work than the unfolded block (Ck <= QK, Ck <= V).
        self.fold_wide = bool(allow_fold and ENABLE_FOLDING and not self.folded and self.H == 1 and wv.shape[1] == self.Ck
                              and self.Ck <= self.QK and self.Ck <= self.V and self.Ck % 8 == 0)
        self.wf, self.bf = _bf16_weight(wf), (bf.float().contiguous() if bf is not None else None)
        # Query-side fold for single-head cross-attends with MANY queries and few keys (the decoders):
        #   S_ij = (LN(q_i) Wq^T + bq) . k_j = LN(q_i) . K'_j + b'_j,    K' = k Wq = LN(z) (Wq^T Wk)^T + Wq^T bk,
        #                                                             b'_j = bq . k_j = LN(z_j) . (Wk^T bq) + bq . bk
        #   out_i = sum_j P_ij (v_j Wf^T) + bf = (P V')_i + bf,         V' = v Wf^T = LN(z) (Wf Wv)^T + Wf bv
        # so ONE projection of the (short) latent array yields both operands [K' | b' | 0.. | V'] with a single rounding,
        # the per-query projections proj_q / final disappear, the contraction shrinks from QK to Cq + 1 (b' rides on a
        # constant-one column of the query rows, which needs a free pad column: Cq % 8 != 0) and the kernel's output is
        # the block's output.  Wiped rows come out as bf, exactly as in the reference (P row = 0).
        self.qfold = None
        pair_ok = self.Cq % 8 != 0 and ops.decoder_attention_supported(self.Cq + 1, self.O)   # pio_decode_kernel
        wide_ok = self.Cq <= self.QK and self.O <= self.V    # explicit-S form: never more work than the unfolded block
        if allow_fold and ENABLE_FOLDING and self.H == 1 and wv.shape[1] == self.Ck and (pair_ok or wide_ok):
            wq64, wk64, wv64, wf64 = (t.double() for t in (wq, wk, wv, wf))
            koff = pad8(self.Cq + 1)
            n_ext = koff + self.O
            w_ext = torch.zeros((n_ext, self.Ck), dtype=torch.float64, device=wq.device)
            b_ext = torch.zeros((n_ext,), dtype=torch.float64, device=wq.device)
            w_ext[:self.Cq] = wq64.t() @ wk64
            b_ext[:self.Cq] = wq64.t() @ bk.double()
            w_ext[self.Cq] = wk64.t() @ bq.double()
            b_ext[self.Cq] = bq.double() @ bk.double()
            w_ext[koff:] = wf64 @ wv64
            b_ext[koff:] = wf64 @ bv.double()
            self.qfold = dict(w=_bf16_weight(w_ext.float()), b=b_ext.float().contiguous(), koff=koff, n=n_ext,
                              pair_ok=pair_ok)
        if self.folded or self.fold_wide:
            # S = (LN(q) Wq^T + bq) Wk . LN(x)^T  (the q.bk term is constant per row and cancels in the softmax)
            # out = (P . LN(x)) (Wf Wv)^T + (Wf bv + bf)          (rows of P sum to one)
            wq64, wk64, wv64, wf64 = (t.double() for t in (wq, wk, wv, wf))
            self.wq_fold = _bf16_weight((wk64.t() @ wq64).float())              # [Ck, Cq]
            self.bq_fold = (wk64.t() @ bq.double()).float().contiguous()        # [Ck]
            self.wo_fold = _bf16_weight((wf64 @ wv64).float())                  # [O, Ck]
            bo = wf64 @ bv.double()
            # rows the reference wipes (no valid key / masked query, :168-175) are zeroed BEFORE `final` and come out as
            # final.bias alone; the folded bias assumes a row of P that sums to one, so those rows take this part back
            self.wf_bv = bo.float().contiguous()
            if bf is not None:
                bo = bo + bf.double()
            self.bo_fold = bo.float().contiguous()
        if self.folded:
            pass
        elif self_attention:
            self.wqkv = _bf16_weight(torch.cat([wq, wk, wv], 0))
            self.bqkv = torch.cat([bq, bk, bv], 0).float().contiguous()
        else:
            self.wq, self.bq = _bf16_weight(wq), bq.float().contiguous()
            self.kv_fused = (self.QK % 8 == 0) and wk.shape[1] == wv.shape[1]   # k_in_channels may differ from v_in
            if self.kv_fused:
                self.wkv = _bf16_weight(torch.cat([wk, wv], 0))
                self.bkv = torch.cat([bk, bv], 0).float().contiguous()
            else:
                self.wk, self.bk = _bf16_weight(wk), bk.float().contiguous()
                self.wv, self.bv = _bf16_weight(wv), bv.float().contiguous()


class PreparedMLP:
    def __init__(self, mlp):
        self.w1, self.b1 = _bf16_weight(mlp.fc1.weight.detach()), mlp.fc1.bias.detach().float().contiguous()
        self.w2, self.b2 = _bf16_weight(mlp.fc2.weight.detach()), mlp.fc2.bias.detach().float().contiguous()
        self.hidden, self.cin = mlp.fc1.weight.shape
        self.cout = mlp.fc2.weight.shape[0]


class PreparedFusedLayer:
    """One SelfAttention block with both LayerNorms folded into the projections that consume them:
    LN(x) W^T = rstd (x (W diag(gamma))^T - mean colsum) + (W beta + b); the GEMM takes the raw bf16 rows and applies
    the per-row normalisation in its epilogue (pio_gemm_args.row_stats_in)."""

    def __init__(self, sa):
        att, mlp = sa.attention, sa.mlp
        d = torch.float64
        g1, b1 = sa.layer_norm1.weight.detach().to(d), sa.layer_norm1.bias.detach().to(d)
        g2, b2 = sa.layer_norm2.weight.detach().to(d), sa.layer_norm2.bias.detach().to(d)
        wqkv = torch.cat([att.proj_q.weight, att.proj_k.weight, att.proj_v.weight], 0).detach().to(d)
        bqkv = torch.cat([att.proj_q.bias, att.proj_k.bias, att.proj_v.bias], 0).detach().to(d)
        self.H = att._num_heads
        self.QK, self.C = att.proj_q.weight.shape
        self.V = att.proj_v.weight.shape[0]
        self.O = att.final.weight.shape[0]
        self.dqk, self.dv = self.QK // self.H, self.V // self.H
        self.scale = 1.0 / math.sqrt(self.dqk)
        self.eps1, self.eps2 = float(sa.layer_norm1.eps), float(sa.layer_norm2.eps)
        self.wqkv = _bf16_weight((wqkv * g1[None, :]).float())
        self.bqkv = (bqkv + wqkv @ b1).float().contiguous()
        # column sums of exactly the bf16 values the tensor core multiplies (so mean * colsum cancels consistently)
        self.cs_qkv = self.wqkv[:, :self.C].double().sum(1).float().contiguous()
        self.wf = _bf16_weight(att.final.weight.detach())
        self.bf = att.final.bias.detach().float().contiguous()
        w1, bb1 = mlp.fc1.weight.detach().to(d), mlp.fc1.bias.detach().to(d)
        self.hidden, self.cin = mlp.fc1.weight.shape
        self.cout = mlp.fc2.weight.shape[0]
        self.w1 = _bf16_weight((w1 * g2[None, :]).float())
        self.b1 = (bb1 + w1 @ b2).float().contiguous()
        self.cs_1 = self.w1[:, :self.cin].double().sum(1).float().contiguous()
        self.w2 = _bf16_weight(mlp.fc2.weight.detach())
        self.b2 = mlp.fc2.bias.detach().float().contiguous()

    def usable(self) -> bool:
        # the producer GEMMs write a raw bf16 copy in 16-column chunks; the tower keeps its width
        return (self.C % 16 == 0 and self.O == self.C and self.cin == self.C and self.cout == self.C
                and _streaming_ok(self.H, self.dqk, self.dv))


def self_attention_block_fused(pf: PreparedFusedLayer, x: Optional[torch.Tensor], xb: torch.Tensor, st: torch.Tensor, *,
                               B: int, N: int, st_mid: torch.Tensor, st_out: Optional[torch.Tensor],
                               x_lo: Optional[torch.Tensor] = None, split: bool = False, split_out: bool = False):
    """SelfAttention.forward with fused LayerNorms.  The residual stream comes in either as x fp32 [M, C] with xb = its
    16-bit rounding, or — x None — as the pair (xb, x_lo) with value xb + x_lo (pio_gemm_args.out_lo16: what the producer
    GEMMs of a large tower write instead of fp32 + raw copy, 4 bytes per element instead of 6; they are bound by HBM
    bytes).  st = per-row partial (sum, sum of squares) of the stream, [M, parts, 2]
    (ops.empty_row_stats); st_mid / st_out are such buffers for the two stream states this block produces (st_out None:
    the block's output feeds no further fused LayerNorm).  Each producer GEMM fills every slot with plain stores and each
    consumer adds a row's slots in index order, so the tower is bit-reproducible.
    split: carry the state between the two halves of the block as a pair; split_out: return the block's output as a pair.
    Returns (y fp32 [M, C] or None, 16-bit y (the rounding / the hi half) or None, lo half or None)."""
    M, C = xb.shape
    dev = xb.device
    d16 = ops.dtype16()
    nqkv = 2 * pf.QK + pf.V
    ld = pad8(nqkv)
    qkv = torch.empty((M, ld), dtype=d16, device=dev)
    ops.gemm(xb, pf.wqkv, M=M, N=nqkv, K=C, lda=xb.stride(0), bias=pf.bqkv, out_bf16=qkv, ldo16=ld,
             row_stats_in=st, ln_colsum=pf.cs_qkv, ln_channels=C, ln_eps=pf.eps1)
    o = attention(qkv, ld, 0, qkv, ld, pf.QK, qkv, ld, 2 * pf.QK, B=B, H=pf.H, Nq=N, Nk=N, dqk=pf.dqk, dv=pf.dv,
                  scale=pf.scale)
    o2 = o.view(M, -1)
    res_in = dict(residual=x, ldr=x.stride(0)) if x is not None else dict(residual_hi16=xb, residual_lo16=x_lo, ldr16=C)
    # a pair that comes in is updated in place (each element is read and rewritten by the same epilogue warp): one
    # (hi, lo) pair per tower instead of three, and the hi half — read by both projections and both producers of a
    # layer — is the working set the kernels ask L2 to keep (pio_gemm2.cu, hint_*)
    inplace = STREAM_INPLACE and split and x is None
    x1b = xb if inplace else torch.empty((M, C), dtype=d16, device=dev)
    if split:
        x1, x1l = None, (x_lo if inplace else torch.empty((M, C), dtype=d16, device=dev))
        ops.gemm(o2, pf.wf, M=M, N=C, K=pf.V, bias=pf.bf, out_bf16=x1b, ldo16=C, out_lo16=x1l, row_stats_out=st_mid,
                 **res_in)
        res_mid = dict(residual_hi16=x1b, residual_lo16=x1l, ldr16=C)
    else:
        x1 = torch.empty((M, C), dtype=torch.float32, device=dev)
        ops.gemm(o2, pf.wf, M=M, N=C, K=pf.V, bias=pf.bf, out_f32=x1, ldo32=C, out_bf16=x1b, ldo16=C,
                 row_stats_out=st_mid, **res_in)
        res_mid = dict(residual=x1, ldr=C)
    h = torch.empty((M, pad8(pf.hidden)), dtype=d16, device=dev)
    ops.gemm(x1b, pf.w1, M=M, N=pf.hidden, K=C, bias=pf.b1, act=1, out_bf16=h, ldo16=h.stride(0),
             row_stats_in=st_mid, ln_colsum=pf.cs_1, ln_channels=C, ln_eps=pf.eps2, reverse_tiles=REVERSE_FC1)
    # fc2 walks its tiles back to front: fc1 and the out-projection wrote h and x1 front to back, so their last rows are
    # what L2 still holds; and the rows fc2 writes last (the first ones) are where the next QKV projection starts
    if split_out:
        yb = x1b if (STREAM_INPLACE and split) else torch.empty((M, C), dtype=d16, device=dev)
        yl = x1l if (STREAM_INPLACE and split) else torch.empty((M, C), dtype=d16, device=dev)
        ops.gemm(h, pf.w2, M=M, N=C, K=pf.hidden, bias=pf.b2, out_bf16=yb, ldo16=C, out_lo16=yl, row_stats_out=st_out,
                 reverse_tiles=REVERSE_FC2, **res_mid)
        return None, yb, yl
    y = torch.empty((M, C), dtype=torch.float32, device=dev)
    yb = torch.empty((M, C), dtype=d16, device=dev) if st_out is not None else None
    ops.gemm(h, pf.w2, M=M, N=C, K=pf.hidden, bias=pf.b2, out_f32=y, ldo32=C,
             out_bf16=yb, ldo16=C if yb is not None else 0, row_stats_out=st_out, reverse_tiles=REVERSE_FC2, **res_mid)
    return y, yb, None


def _versions(module):
    return tuple((p.data_ptr(), p._version, p.device) for p in module.parameters())


def prepared(module, key, builder):
    """Per-module cache of derived weights, rebuilt when any parameter changes (in-place update, load_state_dict,
    .to(device))."""
    cache = module.__dict__.setdefault("_pio_cache", {})
    key = (key, ops.FP16)      # derived weights exist once per 16-bit operand format
    ver = _versions(module)
    hit = cache.get(key)
    if hit is None or hit[0] != ver:
        with torch.no_grad():
            hit = (ver, builder())
        cache[key] = hit
    return hit[1]


# ---------------------------------------------------------------------------------------------------------------
# attention cores
# ---------------------------------------------------------------------------------------------------------------

def _head_slice(t: torch.Tensor, B: int, N: int, ld: int, col0: int, d: int):
    """Return (tensor, ld, col_offset) for columns [col0, col0+d) of a [B*N, ld] bf16 matrix such that the slice
    base is 16-byte aligned; copies into a fresh padded buffer only when it is not."""
    if (col0 * 2) % 16 == 0:
        return t, ld, col0
    out = torch.zeros((B * N, pad8(d)), dtype=ops.dtype16(), device=t.device)
    out[:, :d] = t.view(B * N, ld)[:, col0:col0 + d]
    return out, pad8(d), 0


class GeneralAttentionArgs:
    """The arguments of `Attention.attend` no recipe of the reference uses (transformer_primitives.py:90): a dense
    [B, Nq, Nk] mask that is not an outer product, an additive logit bias broadcastable to [B, H, Nq, Nk], and
    return_matrix.  They are honoured on the explicit S / P path (the softmax kernel applies them)."""

    def __init__(self, *, B, H, Nq, Nk, device, attention_mask=None, attention_bias=None, return_matrix=False):
        self.dense_mask = None
        if attention_mask is not None:
            if tuple(attention_mask.shape) != (B, Nq, Nk):
                raise ValueError(f"attention_mask must be [batch, q_indices, kv_indices] = {(B, Nq, Nk)}, "
                                 f"got {tuple(attention_mask.shape)}")
            self.dense_mask = (attention_mask != 0).to(torch.uint8).contiguous()
        self.bias = None
        if attention_bias is not None:
            self.bias = torch.broadcast_to(attention_bias.to(torch.float32), (B, H, Nq, Nk))   # stride-0 view
        self.matrix = torch.empty((B, H, Nq, Nk), dtype=torch.float32, device=device) if return_matrix else None


def _materialised_attention(q, ldq, qcol, k, ldk, kcol, v, ldv, vcol, *, B, H, Nq, Nk, dqk, dv, scale,
                            key_mask, row_keep, q_bcast, general: Optional[GeneralAttentionArgs] = None):
    """Attention through explicit S and P matrices (GEMM -> softmax -> GEMM), one head at a time.  Used when the
    streaming kernel does not cover the head sizes (d > 384: the decoders, the multimodal encoder) and for the
    general attention arguments (dense mask / bias / return_matrix)."""
    dev = q.device
    ldo = pad8(H * dv)
    O = torch.empty((B, Nq, ldo), dtype=ops.dtype16(), device=dev)
    lds = (Nk + 3) // 4 * 4
    for h in range(H):
        qh, ldqh, qo = _head_slice(q, 1 if q_bcast else B, Nq, ldq, qcol + h * dqk, dqk)
        kh, ldkh, ko = _head_slice(k, B, Nk, ldk, kcol + h * dqk, dqk)
        vh, ldvh, vo = _head_slice(v, B, Nk, ldv, vcol + h * dv, dv)
        S = torch.empty((B, Nq, lds), dtype=torch.float32, device=dev)
        ops.gemm(qh.view(-1)[qo:], kh.view(-1)[ko:], M=Nq, N=Nk, K=dqk, batch=B,
                 strideA=0 if q_bcast else Nq * ldqh, strideB=Nk * ldkh, lda=ldqh, ldb=ldkh,
                 out_f32=S, ldo32=lds, strideO32=Nq * lds)
        if general is None:
            P = ops.softmax_bf16(S, Nk, scale, key_mask, row_keep)
        else:
            P = ops.softmax_bf16(S, Nk, scale, key_mask, row_keep, dense_mask=general.dense_mask,
                                 bias=general.bias[:, h] if general.bias is not None else None,
                                 probs_out=general.matrix[:, h] if general.matrix is not None else None)
        del S
        ldp = P.shape[-1]
        oh = O.view(-1)[h * dv:]
        ops.gemm(P, vh.view(-1)[vo:], M=Nq, N=dv, K=Nk, batch=B, b_mn_major=True,
                 strideA=Nq * ldp, strideB=Nk * ldvh, lda=ldp, ldb=ldvh,
                 out_bf16=oh, ldo16=ldo, strideO16=Nq * ldo)
    return O


def _streaming_ok(H, dqk, dv, same_kv=False) -> bool:
    """Whether the streaming kernels cover these head sizes (otherwise attention goes through explicit S / P)."""
    if not ops.attention_supported(dqk, dv) or not (H == 1 or dqk % 16 == 0):
        return False
    return same_kv or ((dqk + 63) // 64 <= 2 and (dv + 63) // 64 <= 3)


def _materialised_attention_h1(q, k, vT, *, B, Nq, Nk, dqk, dv, scale, key_mask, row_keep, q_bcast):
    """Single-head attention through explicit S and P with every product on the CTA-pair GEMM kernel:
    q bf16 [(1|B)*Nq, pad8(dqk)], k bf16 [B*Nk, pad8(dqk)], vT bf16 [B, dv, pad8(Nk)] (V already transposed by
    swapping the operands of its projection, so P.V has a K-major B operand too).  Returns O bf16 [B, Nq, pad8(dv)]."""
    dev = q.device
    ldq, ldk, nkp = q.shape[-1], k.shape[-1], vT.shape[-1]
    lds = (Nk + 3) // 4 * 4
    S = torch.empty((B, Nq, lds), dtype=torch.float32, device=dev)
    ops.gemm(q, k, M=Nq, N=Nk, K=dqk, batch=B, strideA=0 if q_bcast else Nq * ldq, strideB=Nk * ldk, lda=ldq, ldb=ldk,
             out_f32=S, ldo32=lds, strideO32=Nq * lds)
    P = ops.softmax_bf16(S, Nk, scale, key_mask, row_keep)          # [B, Nq, pad8(Nk)]
    del S
    ldo = pad8(dv)
    ldp = P.shape[-1]
    tiles = B * ((Nq + 255) // 256) * ((dv + 255) // 256)
    if B == 1 and tiles * 4 <= 74 and Nk >= 16384:
        # few query rows, very long key axis (the multimodal encoder: 784 x 704 outputs over 52,097 keys = 12 tiles of
        # the CTA-pair kernel): split the contraction over the key axis so that every SM pair gets a tile, and add the
        # partial products with the partial-merge kernel (all maxima 0, all sums 1 / splits: a plain sum)
        splits = min(32, 74 // tiles, Nk // 4096)
        chunk = -(-Nk // splits) + 63 & ~63          # keys per split, a multiple of the 64-column K chunk
        splits = -(-Nk // chunk)
        Op = torch.empty((splits, 1, 1, Nq, dv), dtype=torch.float32, device=dev)
        # the last split is shorter: zero-padded columns of P (ldp) and of V^T (nkp) beyond Nk contribute nothing, but the
        # split must not read past the rows' pitch -> give it its own launch
        full = splits - 1 if splits * chunk != Nk else splits
        if full > 0:
            ops.gemm(P, vT, M=Nq, N=dv, K=chunk, batch=full, strideA=chunk, strideB=chunk, lda=ldp, ldb=nkp,
                     out_f32=Op, ldo32=dv, strideO32=Nq * dv)
        if full < splits:
            k0 = full * chunk
            ops.gemm(P.view(-1)[k0:], vT.view(-1)[k0:], M=Nq, N=dv, K=Nk - k0, lda=ldp, ldb=nkp,
                     out_f32=Op[full], ldo32=dv)
        ml = torch.zeros((2, splits, 1, 1, Nq), dtype=torch.float32, device=dev)
        ml[1].fill_(1.0 / splits)
        return ops.attention_combine(Op, ml[0], ml[1])
    O = torch.empty((B, Nq, ldo), dtype=ops.dtype16(), device=dev)
    ops.gemm(P, vT, M=Nq, N=dv, K=Nk, batch=B, strideA=Nq * ldp, strideB=dv * nkp, lda=ldp, ldb=nkp,
             out_bf16=O, ldo16=ldo, strideO16=Nq * ldo)
    return O


def _pick_splits(B, H, Nq, Nk, dqk, dv, same_kv, sm_count: int = 148) -> int:
    """Key-axis splits inside one GPU: fill the SMs when (batch x heads x query tiles) is small, and trim the
    partial last wave otherwise (256 CTAs on 148 SMs waste 14 % of the machine; 4 x 256 waste 1 %)."""
    bn = ops._lib.load().pio_attention_key_tile(dqk, dv, 1 if same_kv else 0)
    tiles = (Nk + bn - 1) // bn
    ctas = B * H * ((Nq + 127) // 128)
    # a handful of query tiles over a few thousand keys (the language encoder: 8 heads x 256 latents over 2048 bytes):
    # unsplit, 8 CTAs of the two-tile kernel walk 32 key tiles each (160 us); split, every SM gets four tiles
    few = Nk >= 1024 and ctas * 2 <= sm_count
    min_tiles = 4 if (few and Nk < 4096) else 8
    if ENCODER_KEY_SPLITS > 0:
        cands = [min(ENCODER_KEY_SPLITS, tiles)]
    else:
        if Nk < 4096 and not few:
            return 1
        cands = range(1, 33)
    best, best_eff = 1, 0.0
    for s in cands:
        if s > tiles or tiles // s < min_tiles:
            continue
        if s > 1 and (s - 1) * ((tiles + s - 1) // s) >= tiles:
            continue  # would leave an empty split
        waves = -(-ctas * s // sm_count)
        eff = ctas * s / (waves * sm_count)
        if eff > best_eff + 0.02:
            best, best_eff = s, eff
    return best


def attention(q, ldq, qcol, k, ldk, kcol, v, ldv, vcol, *, B, H, Nq, Nk, dqk, dv, scale, key_mask=None,
              row_keep=None, q_bcast=False, partial=False, num_splits=None,
              general: Optional[GeneralAttentionArgs] = None):
    """Dispatch between the streaming kernel and the materialised path.  q/k/v are flat bf16 [rows, ld] matrices;
    *col give the first column of head 0.  Returns O bf16 [B, Nq, pad8(H*dv)] (or partials)."""
    same_kv = (k.data_ptr() == v.data_ptr()) and kcol == vcol and ldk == ldv and dqk == dv
    flash_ok = (ops.attention_supported(dqk, dv) and (H == 1 or dqk % 16 == 0)
                and all((c * 2) % 16 == 0 for c in (qcol, kcol, vcol)))
    if flash_ok and not same_kv and not ((dqk + 63) // 64 <= 2 and (dv + 63) // 64 <= 3):
        flash_ok = False
    if general is not None:
        flash_ok = False
    if not flash_ok:
        if partial:
            raise RuntimeError("perceiverio_pytorch_b200: key-sharded attention needs head sizes covered by the "
                               f"streaming kernel (got dqk={dqk}, dv={dv})")
        return _materialised_attention(q, ldq, qcol, k, ldk, kcol, v, ldv, vcol, B=B, H=H, Nq=Nq, Nk=Nk, dqk=dqk,
                                       dv=dv, scale=scale, key_mask=key_mask, row_keep=row_keep, q_bcast=q_bcast,
                                       general=general)
    if num_splits is None:
        num_splits = _pick_splits(B, H, Nq, Nk, dqk, dv, same_kv)
    qv, kv_, vv = q.view(-1)[qcol:], k.view(-1)[kcol:], v.view(-1)[vcol:]
    if same_kv:
        vv = kv_
    return ops.attention_fwd(qv, kv_, vv, B=B, H=H, Nq=Nq, Nk=Nk, dqk=dqk, dv=dv,
                             strideQ=0 if q_bcast else Nq * ldq, strideK=Nk * ldk, strideV=Nk * ldv,
                             ldq=ldq, ldk=ldk, ldv=ldv, scale=scale, key_mask=key_mask, row_keep=row_keep,
                             num_splits=num_splits, partial=partial)


# ---------------------------------------------------------------------------------------------------------------
# blocks
# ---------------------------------------------------------------------------------------------------------------

class PreparedTail:
    """A decoder whose final projection has only a handful of outputs (optical flow: 322 -> 2, perceiver.py:179):
    final(x + fc2(h)) = x Wfin^T + h (Wfin W2)^T + (Wfin b2 + bfin), so fc2, its fp32 output array and the separate head
    collapse into one fp32 pass over x and h (pio_linear_f32 with a second operand; all weights stay fp32)."""

    def __init__(self, mlp, final_layer):
        d = torch.float64
        wfin, w2 = final_layer.weight.detach().to(d), mlp.fc2.weight.detach().to(d)
        b = wfin @ mlp.fc2.bias.detach().to(d)
        if final_layer.bias is not None:
            b = b + final_layer.bias.detach().to(d)
        self.wfin = self._pitch4(wfin.float())
        self.w2 = self._pitch4((wfin @ w2).float())
        self.bias = b.float().contiguous()

    @staticmethod
    def _pitch4(w: torch.Tensor) -> torch.Tensor:
        """[n, k] view of a buffer whose row pitch is a multiple of 4 floats (16-byte aligned rows: vector loads)."""
        n, k = w.shape
        buf = torch.zeros((n, (k + 3) // 4 * 4), dtype=torch.float32, device=w.device)
        buf[:, :k] = w
        return buf[:, :k]


def mlp_block(pm: PreparedMLP, x_f32: torch.Tensor, ln_w, ln_b, *, want_bf16_out=False, stats_out=None, xn=None,
              tail: Optional[PreparedTail] = None):
    """x + fc2(gelu(fc1(LN(x)))) on a flat fp32 [M, C] matrix.  Returns (fp32 [M, cout], bf16 copy or None).
    `stats_out` (ops.empty_row_stats(M, cout)) additionally receives the per-row partial (sum, sum of squares) of the
    result — the producer side of the fused LayerNorm of the next block.  `xn`: LN(x) when a producer kernel already
    wrote it.  `tail`: return final_layer(x + fc2(...)) [M, n_out] instead (see PreparedTail)."""
    if xn is None:
        xn = ops.layernorm_bf16(x_f32, ln_w, ln_b)
    _, h = ops.linear(xn, pm.cin, pm.w1, pm.hidden, pm.b1, act=1)
    if tail is not None:
        return ops.linear_f32(x_f32, tail.wfin, tail.bias, x2=h, w2=tail.w2), None
    if stats_out is None:
        return ops.linear(h, pm.hidden, pm.w2, pm.cout, pm.b2, residual=x_f32, want_f32=True, want_bf16=want_bf16_out)
    m = h.shape[0]
    y32 = ops.empty_f32_rows(m, pm.cout, h.device)
    y16 = torch.empty((m, pad8(pm.cout)), dtype=ops.dtype16(), device=h.device)
    ops.gemm(h, pm.w2, M=m, N=pm.cout, K=pm.hidden, bias=pm.b2, residual=x_f32, ldr=x_f32.stride(0),
             out_f32=y32, ldo32=y32.stride(0), out_bf16=y16, ldo16=y16.stride(0), row_stats_out=stats_out)
    return y32, y16


def mlp_only(pm: PreparedMLP, x_bf16: torch.Tensor):
    """fc2(gelu(fc1(x))) without LayerNorm / residual (the bare reference `MLP.forward`)."""
    _, h = ops.linear(x_bf16, pm.cin, pm.w1, pm.hidden, pm.b1, act=1)
    y32, _ = ops.linear(h, pm.hidden, pm.w2, pm.cout, pm.b2, want_f32=True, want_bf16=False)
    return y32


def self_attention_block(pa: PreparedAttention, pm: PreparedMLP, x: torch.Tensor, ln1, ln2,
                         key_mask=None, row_keep=None, general: Optional[GeneralAttentionArgs] = None) -> torch.Tensor:
    """SelfAttention.forward (transformer_primitives.py:275-297) on x fp32 [B, N, C] (contiguous)."""
    B, N, C = x.shape
    x2 = x.reshape(B * N, C)
    xn = ops.layernorm_bf16(x2, ln1.weight, ln1.bias)
    nqkv = 2 * pa.QK + pa.V
    _, qkv = ops.linear(xn, C, pa.wqkv, nqkv, pa.bqkv)
    ld = pad8(nqkv)
    o = attention(qkv, ld, 0, qkv, ld, pa.QK, qkv, ld, 2 * pa.QK, B=B, H=pa.H, Nq=N, Nk=N, dqk=pa.dqk, dv=pa.dv,
                  scale=pa.scale, key_mask=key_mask, row_keep=row_keep, general=general)
    x1, _ = ops.linear(o.view(B * N, -1), pa.V, pa.wf, pa.O, pa.bf, residual=x2, want_f32=True, want_bf16=False)
    y, _ = mlp_block(pm, x1, ln2.weight, ln2.bias)
    return y.view(B, N, -1)


def cross_attention_core(pa: PreparedAttention, qn, kvn, *, B, Nq, Nk, q_bcast, key_mask, row_keep,
                         partial=False, num_splits=None, general: Optional[GeneralAttentionArgs] = None):
    """Projection(s) + attention for a cross-attend.  qn: bf16 [(1|B)*Nq, pad8(Cq)], kvn: bf16 [B*Nk, pad8(Ck)].
    Returns the attention output O bf16 [B, Nq, ld] *before* the output projection, and its logical width."""
    if pa.folded:
        _, qf = ops.linear(qn, pa.Cq, pa.wq_fold, pa.Ck, pa.bq_fold)
        ldq = pad8(pa.Ck)
        ldk = kvn.shape[-1]
        o = attention(qf, ldq, 0, kvn, ldk, 0, kvn, ldk, 0, B=B, H=1, Nq=Nq, Nk=Nk, dqk=pa.Ck, dv=pa.Ck,
                      scale=pa.scale, key_mask=key_mask, row_keep=row_keep, q_bcast=q_bcast, partial=partial,
                      num_splits=num_splits)
        return o, pa.Ck
    _, q = ops.linear(qn, pa.Cq, pa.wq, pa.QK, pa.bq)
    if pa.H == 1 and not partial and general is None and not _streaming_ok(1, pa.dqk, pa.dv):
        # wide single head (the decoders: d = 512 / 1024; the multimodal encoder: 704): explicit S / P, all three
        # products K-major.  V^T [B, V, Nk] comes straight out of its projection with the operands swapped
        # (W_v as the A operand, the LayerNorm'd key/value array as B, bias per output row).
        if pa.kv_fused:
            wk, bk, wv, bv = pa.wkv[:pa.QK], pa.bkv[:pa.QK], pa.wkv[pa.QK:], pa.bkv[pa.QK:]
        else:
            wk, bk, wv, bv = pa.wk, pa.bk, pa.wv, pa.bv
        _, k = ops.linear(kvn, pa.Ck, wk, pa.QK, bk)
        nkp = pad8(Nk)
        ldkv = kvn.shape[-1]
        vT = torch.empty((B, pa.V, nkp), dtype=ops.dtype16(), device=kvn.device)
        if nkp != Nk:
            vT[:, :, Nk:].zero_()       # P's pad columns are zero as well; keep the products finite
        ops.gemm(wv, kvn, M=pa.V, N=Nk, K=pa.Ck, batch=B, strideA=0, strideB=Nk * ldkv, lda=wv.stride(0), ldb=ldkv,
                 bias=bv, bias_mode=2, out_bf16=vT, ldo16=nkp, strideO16=pa.V * nkp)
        o = _materialised_attention_h1(q, k, vT, B=B, Nq=Nq, Nk=Nk, dqk=pa.dqk, dv=pa.dv, scale=pa.scale,
                                       key_mask=key_mask, row_keep=row_keep, q_bcast=q_bcast)
        return o, pa.V
    if pa.kv_fused:
        n = pa.QK + pa.V
        _, kv = ops.linear(kvn, pa.Ck, pa.wkv, n, pa.bkv)
        ld = pad8(n)
        k, ldk, kcol, v, ldv, vcol = kv, ld, 0, kv, ld, pa.QK
    else:
        _, k = ops.linear(kvn, pa.Ck, pa.wk, pa.QK, pa.bk)
        _, v = ops.linear(kvn, pa.Ck, pa.wv, pa.V, pa.bv)
        ldk, kcol, ldv, vcol = pad8(pa.QK), 0, pad8(pa.V), 0
    o = attention(q, pad8(pa.QK), 0, k, ldk, kcol, v, ldv, vcol, B=B, H=pa.H, Nq=Nq, Nk=Nk, dqk=pa.dqk, dv=pa.dv,
                  scale=pa.scale, key_mask=key_mask, row_keep=row_keep, q_bcast=q_bcast, partial=partial,
                  num_splits=num_splits, general=general)
    return o, pa.V


# the query-side fold pays once the per-forward projection of the latent array is small next to the per-query work
QFOLD_MIN_QUERIES = 1024


def use_query_fold(pa: PreparedAttention, Nq: int, Nk: int) -> bool:
    return pa.qfold is not None and pa.qfold["pair_ok"] and Nq >= QFOLD_MIN_QUERIES and Nq >= 2 * Nk


def use_query_fold_explicit(pa: PreparedAttention, Nq: int, Nk: int) -> bool:
    """The same fold for head sizes the decoder kernel does not cover (the classification decoder: 1024 channels), on
    the explicit S / P path: proj_q and final disappear, S = LN(q) K'^T and out = P V' + b_f (+ residual) come straight
    out of the two attention GEMMs."""
    return pa.qfold is not None and not pa.qfold["pair_ok"] and not pa.folded and Nq >= Nk >= 64


def cross_attention_query_fold_explicit(pa: PreparedAttention, q_src, ln_q, kvn, *, B, Nq, Nk, q_bcast, key_mask, row_keep,
                                        residual):
    """q_src fp32 [(1|B)*Nq, Cq] (the un-normalised queries), kvn 16-bit [B*Nk, pad8(Ck)].  Returns the fp32 block output
    [B*Nq, O] before the MLP."""
    f = pa.qfold
    dev = kvn.device
    ldq = f["koff"]                                  # pad8(Cq + 1): room for the constant-one column that carries b'
    qn = ops.layernorm_bf16(q_src, ln_q.weight, ln_q.bias, eps=ln_q.eps, ld=ldq)
    qn[:, pa.Cq] = 1.0
    koff, ldkv = f["koff"], kvn.shape[-1]
    _, kext = ops.linear(kvn, pa.Ck, f["w"][:koff], koff, f["b"][:koff])          # [B*Nk, koff]: K' | b' | 0..
    nkp = pad8(Nk)
    vT = torch.empty((B, pa.O, nkp), dtype=ops.dtype16(), device=dev)            # V'^T, operands of its projection swapped
    if nkp != Nk:
        vT[:, :, Nk:].zero_()
    ops.gemm(f["w"][koff:], kvn, M=pa.O, N=Nk, K=pa.Ck, batch=B, strideA=0, strideB=Nk * ldkv, lda=f["w"].stride(0),
             ldb=ldkv, bias=f["b"][koff:], bias_mode=2, out_bf16=vT, ldo16=nkp, strideO16=pa.O * nkp)
    lds = (Nk + 3) // 4 * 4
    S = torch.empty((B, Nq, lds), dtype=torch.float32, device=dev)
    ops.gemm(qn, kext, M=Nq, N=Nk, K=pa.Cq + 1, batch=B, strideA=0 if q_bcast else Nq * ldq, strideB=Nk * koff, lda=ldq,
             ldb=koff, out_f32=S, ldo32=lds, strideO32=Nq * lds)
    P = ops.softmax_bf16(S, Nk, pa.scale, key_mask, row_keep)
    del S
    ldp = P.shape[-1]
    x = ops.empty_f32_rows(B * Nq, pa.O, dev)
    ldx = x.stride(0)
    if residual is not None:
        assert residual.stride(2) == 1
    ops.gemm(P, vT, M=Nq, N=pa.O, K=Nk, batch=B, strideA=Nq * ldp, strideB=pa.O * nkp, lda=ldp, ldb=nkp, bias=pa.bf,
             residual=residual, ldr=residual.stride(1) if residual is not None else 0,
             strideR=(residual.stride(0) if B > 1 else 0) if residual is not None else 0,
             out_f32=x, ldo32=ldx, strideO32=Nq * ldx)
    return x


def cross_attention_query_fold(pa: PreparedAttention, qn, kvn, *, B, Nq, Nk, q_bcast, key_mask, row_keep, residual,
                               ln=None):
    """Attention + output projection (+ query residual) of a single-head cross-attend through the query-side fold and
    the query-tiled decoder kernel (pio_decoder_attention_fwd).  qn: 16-bit [(1|B)*Nq, pad8(Cq)] (its first pad column
    becomes the constant one that carries the per-key logit bias), kvn: 16-bit [B*Nk, pad8(Ck)].
    Returns the fp32 block output [B*Nq, O] before the MLP — and, with ln = (gamma, beta, eps), its LayerNorm as 16-bit
    rows, written by the same kernel."""
    f = pa.qfold
    qn[:, pa.Cq] = 1.0
    _, kvp = ops.linear(kvn, pa.Ck, f["w"], f["n"], f["b"])      # [B*Nk, pad8(n)]: K' | b' | 0.. | V'
    ld = kvp.shape[-1]
    ldq = qn.shape[-1]
    if residual is not None:
        assert residual.stride(2) == 1
    return ops.decoder_attention(qn, kvp, kvp.view(-1)[f["koff"]:], B=B, Nq=Nq, Nk=Nk, dqk=pa.Cq + 1, dv=pa.O,
                                 ldq=ldq, ldk=ld, ldv=ld, strideQ=0 if q_bcast else Nq * ldq, strideK=Nk * ld,
                                 strideV=Nk * ld, scale=pa.scale, key_mask=key_mask, row_keep=row_keep, bias=pa.bf,
                                 residual=residual, ldr=residual.stride(1) if residual is not None else 0,
                                 strideR=(residual.stride(0) if B > 1 else 0) if residual is not None else 0, ln=ln)


# the key-side fold on the explicit path: long input arrays only (the projections it removes scale with Nk)
KFOLD_WIDE_MIN_KEYS = 4096


def use_key_fold_wide(pa: PreparedAttention, Nq: int, Nk: int) -> bool:
    return pa.fold_wide and Nk >= KFOLD_WIDE_MIN_KEYS and Nk >= 4 * Nq


def cross_attention_key_fold_wide(pa: PreparedAttention, qn, kvn, *, B, Nq, Nk, q_bcast, key_mask, row_keep):
    """Attention of a wide single-head encoder cross-attend through the key-side fold (PreparedAttention.fold_wide):
    qn 16-bit [(1|B)*Nq, pad8(Cq)], kvn 16-bit [B*Nk, pad8(Ck)] = LN(x).  S = Q' kvn^T on the CTA-pair kernel (kvn is its
    K-major B operand), row softmax, O' = P kvn with kvn as the MN-major B operand of the single-CTA kernel — no
    transposed copy, no K / V projection.  Returns (O' 16-bit [B, Nq, pad8(Ck)], Ck) for cross_attention_out(folded=True)."""
    dev = kvn.device
    _, qf = ops.linear(qn, pa.Cq, pa.wq_fold, pa.Ck, pa.bq_fold)
    ldq, ldk = qf.shape[-1], kvn.shape[-1]
    lds = (Nk + 3) // 4 * 4
    S = torch.empty((B, Nq, lds), dtype=torch.float32, device=dev)
    ops.gemm(qf, kvn, M=Nq, N=Nk, K=pa.Ck, batch=B, strideA=0 if q_bcast else Nq * ldq, strideB=Nk * ldk, lda=ldq,
             ldb=ldk, out_f32=S, ldo32=lds, strideO32=Nq * lds)
    P = ops.softmax_bf16(S, Nk, pa.scale, key_mask, row_keep)          # [B, Nq, pad8(Nk)], pad columns zero
    del S
    ldp, dv = P.shape[-1], pa.Ck
    tiles = B * ((Nq + 127) // 128) * ((dv + 255) // 256)
    if B == 1 and tiles * 4 <= 148 and Nk >= 16384:
        # few query rows, very long key axis: split the contraction over the key axis so that every SM gets a tile, and add
        # the partial products with the partial-merge kernel (all maxima 0, all sums 1 / splits: a plain sum)
        splits = min(32, 148 // tiles, Nk // 4096)
        chunk = -(-Nk // splits) + 63 & ~63          # keys per split, a multiple of the 64-row K chunk
        splits = -(-Nk // chunk)
        Op = torch.empty((splits, 1, 1, Nq, dv), dtype=torch.float32, device=dev)
        full = splits - 1 if splits * chunk != Nk else splits      # the last split is shorter: its own launch
        if full > 0:
            ops.gemm(P, kvn, M=Nq, N=dv, K=chunk, batch=full, b_mn_major=True, strideA=chunk, strideB=chunk * ldk,
                     lda=ldp, ldb=ldk, out_f32=Op, ldo32=dv, strideO32=Nq * dv)
        if full < splits:
            k0 = full * chunk
            ops.gemm(P.view(-1)[k0:], kvn.view(-1)[k0 * ldk:], M=Nq, N=dv, K=Nk - k0, b_mn_major=True, lda=ldp, ldb=ldk,
                     out_f32=Op[full], ldo32=dv)
        ml = torch.zeros((2, splits, 1, 1, Nq), dtype=torch.float32, device=dev)
        ml[1].fill_(1.0 / splits)
        return ops.attention_combine(Op, ml[0], ml[1]), dv
    ldo = pad8(dv)
    O = torch.empty((B, Nq, ldo), dtype=ops.dtype16(), device=dev)
    ops.gemm(P, kvn, M=Nq, N=dv, K=Nk, batch=B, b_mn_major=True, strideA=Nq * ldp, strideB=Nk * ldk, lda=ldp, ldb=ldk,
             out_bf16=O, ldo16=ldo, strideO16=Nq * ldo)
    return O, dv


def cross_attention_out(pa: PreparedAttention, o: torch.Tensor, width: int, *, B, Nq, residual, row_keep=None,
                        folded: Optional[bool] = None):
    """Output projection (+ query residual) of a cross-attend: returns fp32 [B*Nq, O].  `row_keep` (u8 [B, Nq]) marks the
    rows the attention kernel did not wipe; it only matters on the folded path (see PreparedAttention.wf_bv)."""
    folded = pa.folded if folded is None else folded
    if folded:
        w, b = pa.wo_fold, pa.bo_fold
    else:
        w, b = pa.wf, pa.bf
    o2 = o.view(B * Nq, -1)
    y = ops.empty_f32_rows(B * Nq, pa.O, o.device)
    ldy = y.stride(0)
    if residual is None:
        ops.gemm(o2, w, M=B * Nq, N=pa.O, K=width, bias=b, out_f32=y, ldo32=ldy)
    else:
        # residual is the un-normalised query [B, Nq, Cq], possibly a stride-0 batch broadcast
        assert residual.stride(2) == 1
        ops.gemm(o2, w, M=Nq, N=pa.O, K=width, batch=B, strideA=Nq * o2.shape[1], strideB=0, bias=b,
                 residual=residual, ldr=residual.stride(1), strideR=residual.stride(0) if B > 1 else 0,
                 out_f32=y, ldo32=ldy, strideO32=Nq * ldy)
    if folded and row_keep is not None:
        wiped = (row_keep.reshape(B * Nq, 1) == 0).to(torch.float32)
        y.addcmul_(wiped, pa.wf_bv[None, :], value=-1.0)
    return y


def _as_u8(mask: Optional[torch.Tensor]):
    if mask is None:
        return None
    return mask.to(torch.uint8).contiguous()


def cross_attention_block(pa: PreparedAttention, pm: PreparedMLP, inputs_q: torch.Tensor, inputs_kv: torch.Tensor,
                          ln_q, ln_kv, ln2, *, use_query_residual: bool, key_mask=None, row_keep=None,
                          want_bf16_out=False, shard=None, stats_out=None,
                          general: Optional[GeneralAttentionArgs] = None, tail: Optional[PreparedTail] = None):
    """CrossAttention.forward (transformer_primitives.py:371-406).

    inputs_q fp32 [B, Nq, Cq] (batch stride may be 0), inputs_kv fp32 [B, Nk, Ck].  `shard`, if given, is a
    parallel.KeyShard describing a key-axis shard of inputs_kv across ranks."""
    B, Nq, Cq = inputs_q.shape
    Nk, Ck = inputs_kv.shape[1], inputs_kv.shape[2]
    q_bcast = B > 1 and inputs_q.stride(0) == 0
    kvn = None
    if isinstance(inputs_kv, PositionedInput):
        # features + batch-invariant position table: normalised without building the concatenated array (inputs.py)
        cf, cp = inputs_kv.features.shape[2], inputs_kv.pos.shape[1]
        if ops.layernorm_concat_supported(B, Nk, cf, cp):
            kvn = ops.layernorm_concat_bf16(inputs_kv.features, inputs_kv.pos, ln_kv.weight, ln_kv.bias, eps=ln_kv.eps)
        else:
            inputs_kv = inputs_kv.dense()
    if kvn is None:
        if inputs_kv.stride(2) != 1 or inputs_kv.stride(0) != Nk * inputs_kv.stride(1):
            inputs_kv = inputs_kv.contiguous()
        kvn = ops.layernorm_bf16(inputs_kv.view(B * Nk, Ck) if inputs_kv.is_contiguous()
                                 else inputs_kv.reshape(B * Nk, Ck), ln_kv.weight, ln_kv.bias)
    q_src = inputs_q[0] if q_bcast else (inputs_q if inputs_q.is_contiguous() else inputs_q.contiguous()).view(B * Nq, Cq)
    if q_src.stride(-1) != 1:
        q_src = q_src.contiguous()
    km, rk = _as_u8(key_mask), _as_u8(row_keep)
    if use_query_residual:
        res = inputs_q if inputs_q.stride(2) == 1 else inputs_q.contiguous()
    else:
        res = None
    if general is None and shard is None and use_query_fold_explicit(pa, Nq, Nk):
        x = cross_attention_query_fold_explicit(pa, q_src, ln_q, kvn, B=B, Nq=Nq, Nk=Nk, q_bcast=q_bcast, key_mask=km,
                                                row_keep=rk, residual=res)
        y32, y16 = mlp_block(pm, x, ln2.weight, ln2.bias, want_bf16_out=want_bf16_out, stats_out=stats_out, tail=tail)
        return y32.view(B, Nq, -1), y16
    qn = ops.layernorm_bf16(q_src, ln_q.weight, ln_q.bias)
    if general is None and shard is None and use_query_fold(pa, Nq, Nk):
        x, xn = cross_attention_query_fold(pa, qn, kvn, B=B, Nq=Nq, Nk=Nk, q_bcast=q_bcast, key_mask=km, row_keep=rk,
                                           residual=res, ln=(ln2.weight, ln2.bias, ln2.eps))
        y32, y16 = mlp_block(pm, x, ln2.weight, ln2.bias, want_bf16_out=want_bf16_out, stats_out=stats_out, xn=xn,
                             tail=tail)
        return y32.view(B, Nq, -1), y16
    folded = None
    if general is not None:
        assert shard is None and not pa.folded
        o, width = cross_attention_core(pa, qn, kvn, B=B, Nq=Nq, Nk=Nk, q_bcast=q_bcast, key_mask=km, row_keep=rk,
                                        general=general)
    elif shard is None and use_key_fold_wide(pa, Nq, Nk):
        o, width = cross_attention_key_fold_wide(pa, qn, kvn, B=B, Nq=Nq, Nk=Nk, q_bcast=q_bcast, key_mask=km,
                                                 row_keep=rk)
        folded = True
    elif shard is None:
        o, width = cross_attention_core(pa, qn, kvn, B=B, Nq=Nq, Nk=Nk, q_bcast=q_bcast, key_mask=km, row_keep=rk)
    else:
        parts, width = cross_attention_core(pa, qn, kvn, B=B, Nq=Nq, Nk=Nk, q_bcast=q_bcast, key_mask=km,
                                            row_keep=None, partial=True,
                                            num_splits=shard.local_splits if shard.local_splits > 0 else None)
        # the wipe of rows without any valid key (on any rank) comes out of the merged sums: no reduction of the mask
        o, rk = shard.combine(parts, row_keep=rk, want_alive=True)
    x = cross_attention_out(pa, o, width, B=B, Nq=Nq, residual=res, row_keep=rk, folded=folded)
    y32, y16 = mlp_block(pm, x, ln2.weight, ln2.bias, want_bf16_out=want_bf16_out, stats_out=stats_out, tail=tail)
    return y32.view(B, Nq, -1), y16
