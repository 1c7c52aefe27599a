"""Validation-precision lowering of the Perceiver IO blocks (BASELINE.json north_star: <= 1e-4 against the fp32
reference).  Enabled with `perceiverio_pytorch_b200.set_precision("bf16x3")`.

Same kernels, same C ABI as the bf16 path, but every product is evaluated with **bf16 x 2 split operands**: a value v
is carried as hi = bf16(v) and lo = bf16(v - hi) (16 mantissa bits together), and

    a . b  ~=  a_hi b_hi + a_lo b_hi + a_hi b_lo          (fp32 accumulation in TMEM; the lo.lo term is ~2^-18)

is ONE tcgen05 GEMM over a three-times longer contraction: the A operand is laid out [hi | lo | hi], the B operand
[hi | hi | lo] (pio_layernorm_args.split / pio_softmax_args.split produce those layouts on the device).  Plain TF32
would not reach 1e-4 (10 mantissa bits -> 1e-3 on the 26..48-layer towers).  All activations
between kernels stay fp32; attention goes through explicit S and P matrices (GEMM -> fp32 softmax -> GEMM), no
folding, no streaming kernel.  This mode is ~4-6x slower than the bf16 path and exists to validate it.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import ops
from .ops import BF16, pad8


def split_weight(w: torch.Tensor) -> torch.Tensor:
    """fp32 [N, K] -> bf16 [N, 3 * pad8(K)] laid out [hi | hi | lo] (B side of a split product).  One-time host prep."""
    n, k = w.shape
    kp = pad8(k)
    w = w.detach().float()
    hi = w.to(BF16)
    lo = (w - hi.float()).to(BF16)
    out = torch.zeros((n, 3 * kp), dtype=BF16, device=w.device)
    out[:, :k] = hi
    out[:, kp:kp + k] = hi
    out[:, 2 * kp:2 * kp + k] = lo
    return out


class Weights:
    """Split weights of an nn.Linear."""

    def __init__(self, lin):
        self.w3 = split_weight(lin.weight)
        self.b = lin.bias.detach().float().contiguous() if lin.bias is not None else None
        self.n, self.k = lin.weight.shape


def linear(x: torch.Tensor, wt: Weights, *, ln=None, act: int = 0, residual: Optional[torch.Tensor] = None):
    """fp32 [M, K] (uniform row stride) -> fp32 [M, N]:  act(LN(x) @ W^T + b) (+ residual)."""
    m = x.shape[0]
    a3 = ops.layernorm_bf16(x, ln.weight if ln is not None else None, ln.bias if ln is not None else None,
                            normalize=ln is not None, split=1)
    y = torch.empty((m, wt.n), dtype=torch.float32, device=x.device)
    ops.gemm(a3, wt.w3, M=m, N=wt.n, K=3 * pad8(wt.k), bias=wt.b, act=act,
             residual=residual, ldr=residual.stride(0) if residual is not None else 0, out_f32=y, ldo32=wt.n)
    return y


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, B, H, Nq, Nk, dqk, dv, scale,
              key_mask=None, row_keep=None, q_bcast=False, general=None) -> torch.Tensor:
    """q fp32 [(1|B)*Nq, H*dqk], k [B*Nk, H*dqk], v [B*Nk, H*dv] -> fp32 [B*Nq, H*dv] (heads merged head-major).
    `general`: engine.GeneralAttentionArgs (dense mask / additive bias / probabilities out), applied by the softmax."""
    dev = q.device
    o = torch.empty((B * Nq, H * dv), dtype=torch.float32, device=dev)
    dqp, dvp, nkp = pad8(dqk), pad8(dv), pad8(Nk)
    lds = (Nk + 3) // 4 * 4
    for h in range(H):
        qa = ops.layernorm_bf16(q[:, h * dqk:(h + 1) * dqk], None, None, normalize=False, split=1)   # [.., 3 dqp]
        kb = ops.layernorm_bf16(k[:, h * dqk:(h + 1) * dqk], None, None, normalize=False, split=2)
        vb = ops.layernorm_bf16(v[:, h * dv:(h + 1) * dv], None, None, normalize=False, split=2)     # [hi | hi | lo]
        S = torch.empty((B, Nq, lds), dtype=torch.float32, device=dev)
        ops.gemm(qa, kb, M=Nq, N=Nk, K=3 * dqp, batch=B, strideA=0 if q_bcast else Nq * 3 * dqp, strideB=Nk * 3 * dqp,
                 lda=3 * dqp, ldb=3 * dqp, out_f32=S, ldo32=lds, strideO32=Nq * lds)
        if general is None:
            P3 = ops.softmax_bf16(S, Nk, scale, key_mask, row_keep, split=True)                        # [hi | lo | hi]
        else:
            P3 = ops.softmax_bf16(S, Nk, scale, key_mask, row_keep, split=True, dense_mask=general.dense_mask,
                                  bias=general.bias[:, h] if general.bias is not None else None,
                                  probs_out=general.matrix[:, h] if general.matrix is not None else None)
        del S
        oh = o.view(-1)[h * dv:]
        ldo = H * dv
        flat_p, flat_v = P3.view(-1), vb.view(-1)
        # O = P_hi V_hi + P_lo V_hi + P_hi V_lo, accumulated through the fp32 residual input of the epilogue
        for i, (pseg, vseg) in enumerate(((0, 0), (1, 0), (0, 2))):
            ops.gemm(flat_p[pseg * nkp:], flat_v[vseg * dvp:], M=Nq, N=dv, K=Nk, batch=B, b_mn_major=True,
                     strideA=Nq * 3 * nkp, strideB=Nk * 3 * dvp, lda=3 * nkp, ldb=3 * dvp,
                     residual=oh if i > 0 else None, ldr=ldo, strideR=Nq * ldo,
                     out_f32=oh, ldo32=ldo, strideO32=Nq * ldo)
    return o


def _prepared(module, key, builder):
    from . import engine
    return engine.prepared(module, key, builder)


class _Att:
    def __init__(self, att):
        self.q, self.k, self.v, self.f = Weights(att.proj_q), Weights(att.proj_k), Weights(att.proj_v), Weights(att.final)
        self.H = att._num_heads
        self.dqk = self.q.n // self.H
        self.dv = self.v.n // self.H
        self.scale = 1.0 / math.sqrt(self.dqk)


class _Mlp:
    def __init__(self, mlp):
        self.fc1, self.fc2 = Weights(mlp.fc1), Weights(mlp.fc2)


def _u8(mask):
    return None if mask is None else mask.to(torch.uint8).contiguous()


def self_attention_block(mod, x: torch.Tensor, key_mask=None, row_keep=None, general=None) -> torch.Tensor:
    """SelfAttention.forward (transformer_primitives.py:275-297) on x fp32 [B, N, C]."""
    pa = _prepared(mod.attention, "v_att", lambda: _Att(mod.attention))
    pm = _prepared(mod.mlp, "v_mlp", lambda: _Mlp(mod.mlp))
    B, N, C = x.shape
    x2 = x.reshape(B * N, C)
    q = linear(x2, pa.q, ln=mod.layer_norm1)
    k = linear(x2, pa.k, ln=mod.layer_norm1)
    v = linear(x2, pa.v, ln=mod.layer_norm1)
    o = attention(q, k, v, B=B, H=pa.H, Nq=N, Nk=N, dqk=pa.dqk, dv=pa.dv, scale=pa.scale,
                  key_mask=_u8(key_mask), row_keep=_u8(row_keep), general=general)
    x1 = linear(o, pa.f, residual=x2)
    h = linear(x1, pm.fc1, ln=mod.layer_norm2, act=1)
    y = linear(h, pm.fc2, residual=x1)
    return y.view(B, N, -1)


def cross_attention_block(mod, inputs_q: torch.Tensor, inputs_kv: torch.Tensor, *, key_mask=None, row_keep=None,
                          general=None):
    """CrossAttention.forward (transformer_primitives.py:371-406): fp32 [B, Nq, Cq] x [B, Nk, Ck] -> fp32 [B, Nq, Cq]."""
    pa = _prepared(mod.attention, "v_att", lambda: _Att(mod.attention))
    pm = _prepared(mod.mlp, "v_mlp", lambda: _Mlp(mod.mlp))
    B, Nq, Cq = inputs_q.shape
    Nk, Ck = inputs_kv.shape[1], inputs_kv.shape[2]
    q_bcast = B > 1 and inputs_q.stride(0) == 0
    kv2 = inputs_kv.contiguous().view(B * Nk, Ck)
    q2 = inputs_q[0] if q_bcast else inputs_q.contiguous().view(B * Nq, Cq)
    if q2.stride(-1) != 1:
        q2 = q2.contiguous()
    q = linear(q2, pa.q, ln=mod.layer_norm_q)
    k = linear(kv2, pa.k, ln=mod.layer_norm_kv)
    v = linear(kv2, pa.v, ln=mod.layer_norm_kv)
    o = attention(q, k, v, B=B, H=pa.H, Nq=Nq, Nk=Nk, dqk=pa.dqk, dv=pa.dv, scale=pa.scale,
                  key_mask=_u8(key_mask), row_keep=_u8(row_keep), q_bcast=q_bcast, general=general)
    res = None
    if mod._use_query_residual:
        res = inputs_q.expand(B, Nq, Cq).contiguous().view(B * Nq, Cq)
    x = linear(o, pa.f, residual=res)
    h = linear(x, pm.fc1, ln=mod.layer_norm2, act=1)
    y = linear(h, pm.fc2, residual=x)
    return y.view(B, Nq, -1)


def attention_module(mod, inputs_q, inputs_k, inputs_v, key_mask=None, row_keep=None, general=None):
    """Bare Attention.forward (transformer_primitives.py:90-115)."""
    pa = _prepared(mod, "v_att", lambda: _Att(mod))
    B, Nq, _ = inputs_q.shape
    Nk = inputs_k.shape[1]
    q = linear(inputs_q.contiguous().view(B * Nq, -1), pa.q)
    k = linear(inputs_k.contiguous().view(B * Nk, -1), pa.k)
    v = linear(inputs_v.contiguous().view(B * Nk, -1), pa.v)
    o = attention(q, k, v, B=B, H=pa.H, Nq=Nq, Nk=Nk, dqk=pa.dqk, dv=pa.dv, scale=pa.scale,
                  key_mask=_u8(key_mask), row_keep=_u8(row_keep), general=general)
    return linear(o, pa.f).view(B, Nq, -1)


def mlp_module(mod, x):
    pm = _prepared(mod, "v_mlp", lambda: _Mlp(mod))
    shape = x.shape
    h = linear(x.contiguous().view(-1, shape[-1]), pm.fc1, act=1)
    return linear(h, pm.fc2).view(*shape[:-1], -1)


def final_layer(lin, y32: torch.Tensor) -> torch.Tensor:
    wt = _prepared(lin, "v_w", lambda: Weights(lin))
    B, Nq, C = y32.shape
    return linear(y32.view(B * Nq, C), wt).view(B, Nq, -1)
