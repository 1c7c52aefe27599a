"""Drop-in replacements for `perceiver_io/transformer_primitives.py` of JOBR0/PerceiverIO_Pytorch.

Same class names, constructor kwargs, sub-module / parameter names (hence the same state_dict) and forward
signatures as the reference (transformer_primitives.py:10-15, :18-180, :183-216, :219-297, :300-406); the forward
passes run on hand-written sm_100a kernels through the C ABI in include/pio_b200.h.  Inference only (the reference
has no training code); dropout must be inactive (eval mode or p = 0), as in every recipe of the reference.

There is no CPU path: calling a forward with CPU tensors raises.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import engine, ops, validate


# ---------------------------------------------------------------------------------------------------------------
# init helpers (re-statement of the documented timm initialisers the reference imports; init fidelity does not
# matter for parity because both sides load the same state_dict)
# ---------------------------------------------------------------------------------------------------------------

def variance_scaling_(tensor: torch.Tensor, scale: float = 1.0, mode: str = "fan_in",
                      distribution: str = "truncated_normal"):
    fan_out, fan_in = tensor.shape[0], tensor.shape[1]
    denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": 0.5 * (fan_in + fan_out)}[mode]
    variance = scale / denom
    with torch.no_grad():
        if distribution == "truncated_normal":
            std = math.sqrt(variance) / .87962566103423978
            return nn.init.trunc_normal_(tensor, std=std, a=-2 * std, b=2 * std)
        if distribution == "normal":
            return tensor.normal_(std=math.sqrt(variance))
        bound = math.sqrt(3 * variance)
        return tensor.uniform_(-bound, bound)


def lecun_normal_(tensor: torch.Tensor):
    return variance_scaling_(tensor, 1.0, "fan_in", "truncated_normal")


def make_cross_attention_mask(query_mask: torch.Tensor, kv_mask: torch.Tensor) -> torch.Tensor:
    """[B, Nq] x [B, Nk] -> dense [B, Nq, Nk] mask (transformer_primitives.py:10-15).  The two rank-1 factors are
    attached to the result so that the kernels can use them directly instead of the dense matrix."""
    batch_size, query_len = query_mask.shape
    _, key_len = kv_mask.shape
    mask = torch.einsum("bi,bj->bij", (query_mask, kv_mask))
    assert mask.shape == (batch_size, query_len, key_len)
    mask._pio_factors = (query_mask, kv_mask, _version_of(mask), _version_of(query_mask), _version_of(kv_mask))
    return mask


def _version_of(t: torch.Tensor):
    """In-place modification counter of a tensor (None for inference tensors, which do not track one)."""
    try:
        return t._version
    except RuntimeError:
        return None


def _factor_mask(attention_mask: Optional[torch.Tensor]):
    """Dense [B, Nq, Nk] mask -> (row_keep [B, Nq], key_mask [B, Nk]) in the kernels' terms.

    The reference wipes a row when its whole mask row is False (:168-175); for an outer-product mask that is
    `query_mask[i] == 0 or no key valid`.  Masks that are not an outer product are not produced by any caller in
    the reference (perceiver.py:99-102, :171-175); they return None (the caller then takes the general path)."""
    if attention_mask is None:
        return None, None
    fac = getattr(attention_mask, "_pio_factors", None)
    if fac is not None and (fac[2], fac[3], fac[4]) != (_version_of(attention_mask), _version_of(fac[0]),
                                                        _version_of(fac[1])):
        fac = None      # the mask (or a factor) was edited in place after make_cross_attention_mask: factor it again
    if fac is None:
        m = attention_mask != 0
        qm, km = m.any(dim=2), m.any(dim=1)
        if not torch.equal(qm[:, :, None] & km[:, None, :], m):
            return None
    else:
        qm, km = fac[0].to(torch.bool), fac[1].to(torch.bool)
    row_keep = qm & km.any(dim=1, keepdim=True)
    return row_keep, km


def _route_mask(attention_mask, attention_bias, return_matrix, *, B, H, Nq, Nk, device):
    """Decide between the factored fast path and the general path of `Attention.attend`.

    Returns (row_keep, key_mask, general): general is None on the fast path (no bias, no return_matrix, and the mask
    — if any — is an outer product), otherwise an engine.GeneralAttentionArgs carrying the dense mask / bias /
    matrix buffer (and row_keep / key_mask are None: the dense mask subsumes them)."""
    if attention_bias is None and not return_matrix:
        fac = _factor_mask(attention_mask)
        if fac is not None:
            return fac[0], fac[1], None
    return None, None, engine.GeneralAttentionArgs(B=B, H=H, Nq=Nq, Nk=Nk, device=device,
                                                   attention_mask=attention_mask, attention_bias=attention_bias,
                                                   return_matrix=return_matrix)


def _check_inference(module: nn.Module, *probs):
    if module.training and any(p > 0 for p in probs):
        raise RuntimeError("perceiverio_pytorch_b200 is inference-only: call .eval() or use dropout_prob=0")


class Attention(nn.Module):
    """Multi-headed {cross, self}-attention (reference: transformer_primitives.py:18-180)."""

    def __init__(self, q_in_channels: int, k_in_channels: int = None, v_in_channels: int = None, num_heads: int = 8,
                 init_scale: float = 1.0, with_final_bias: bool = True, final_init_scale_multiplier: float = 1.,
                 dropout_prob: float = 0.0, qk_out_channels: int = None, v_out_channels: int = None,
                 output_channels: int = None):
        super().__init__()
        self._num_heads = num_heads
        final_init_scale = final_init_scale_multiplier * init_scale
        if qk_out_channels is None:
            qk_out_channels = q_in_channels
        if v_out_channels is None:
            v_out_channels = qk_out_channels
        if output_channels is None:
            output_channels = v_out_channels
        self._qk_channels_per_head = qk_out_channels // num_heads
        self._v_channels_per_head = v_out_channels // num_heads
        if qk_out_channels % num_heads != 0:
            raise ValueError(f"qk_out_channels ({qk_out_channels}) must be divisible by"
                             f" num_heads ({num_heads}).")
        if v_out_channels % num_heads != 0:
            raise ValueError(f"v_channels ({v_out_channels}) must be divisible by"
                             f" num_heads ({num_heads}).")
        self.proj_q = nn.Linear(q_in_channels, qk_out_channels, bias=True)
        self.proj_k = nn.Linear(k_in_channels, qk_out_channels, bias=True)
        self.proj_v = nn.Linear(v_in_channels, v_out_channels, bias=True)
        for lin in (self.proj_q, self.proj_k, self.proj_v):
            variance_scaling_(lin.weight, scale=init_scale)
            nn.init.constant_(lin.bias, 0)
        self._dropout_prob = dropout_prob
        self.dropout = nn.Dropout(dropout_prob)
        self.final = nn.Linear(v_out_channels, output_channels, bias=with_final_bias)
        variance_scaling_(self.final.weight, scale=final_init_scale)
        nn.init.constant_(self.final.bias, 0)

    def forward(self, inputs_q, inputs_k, inputs_v, attention_mask=None, attention_bias=None, return_matrix=False):
        _check_inference(self, self._dropout_prob)
        ops._need_cuda(inputs_q, inputs_k, inputs_v)
        B, Nq, Cq = inputs_q.shape
        Nk = inputs_k.shape[1]
        if B == 0 or Nq == 0:   # empty batch / no queries: nothing to launch
            y = inputs_q.new_empty(B, Nq, self.final.out_features)
            return (inputs_q.new_empty(B, self._num_heads, Nq, Nk), y) if return_matrix else y
        pa = engine.prepared(self, "plain", lambda: engine.PreparedAttention(self, self_attention=False,
                                                                             allow_fold=False))
        row_keep, key_mask, general = _route_mask(attention_mask, attention_bias, return_matrix, B=B, H=pa.H, Nq=Nq,
                                                  Nk=Nk, device=inputs_q.device)
        if engine.PRECISION == "bf16x3":
            y = validate.attention_module(self, inputs_q, inputs_k, inputs_v, key_mask, row_keep, general)
            return (general.matrix, y) if return_matrix else y
        qn = ops.layernorm_bf16(inputs_q.contiguous().view(B * Nq, Cq), None, None, normalize=False)
        kn = ops.layernorm_bf16(inputs_k.contiguous().view(B * Nk, -1), None, None, normalize=False)
        _, q = ops.linear(qn, pa.Cq, pa.wq, pa.QK, pa.bq)
        same_kv = inputs_v is inputs_k
        vn = kn if same_kv else ops.layernorm_bf16(inputs_v.contiguous().view(B * Nk, -1), None, None,
                                                   normalize=False)
        if pa.kv_fused and same_kv:
            n = pa.QK + pa.V
            _, kv = ops.linear(kn, pa.Ck, pa.wkv, n, pa.bkv)
            k, ldk, kcol, v, ldv, vcol = kv, ops.pad8(n), 0, kv, ops.pad8(n), pa.QK
        else:
            if pa.kv_fused:
                wk, bk = pa.wkv[:pa.QK], pa.bkv[:pa.QK]
                wv, bv = pa.wkv[pa.QK:], pa.bkv[pa.QK:]
            else:
                wk, bk, wv, bv = pa.wk, pa.bk, pa.wv, pa.bv
            _, k = ops.linear(kn, pa.Ck, wk, pa.QK, bk)
            _, v = ops.linear(vn, self.proj_v.in_features, wv, pa.V, bv)
            ldk, kcol, ldv, vcol = ops.pad8(pa.QK), 0, ops.pad8(pa.V), 0
        o = engine.attention(q, ops.pad8(pa.QK), 0, k, ldk, kcol, v, ldv, vcol, B=B, H=pa.H, Nq=Nq, Nk=Nk,
                             dqk=pa.dqk, dv=pa.dv, scale=pa.scale, key_mask=engine._as_u8(key_mask),
                             row_keep=engine._as_u8(row_keep), general=general)
        y, _ = ops.linear(o.view(B * Nq, -1), pa.V, pa.wf, pa.O, pa.bf, want_f32=True, want_bf16=False)
        y = y.contiguous().view(B, Nq, -1)
        return (general.matrix, y) if return_matrix else y


class MLP(nn.Module):
    """Transformer-style dense block (reference: transformer_primitives.py:183-216)."""

    def __init__(self, in_channels: int, out_channels: int = None, widening_factor: int = 4,
                 dropout_prob: float = 0.0, init_scale: float = 1.):
        super().__init__()
        out_channels = out_channels or in_channels
        self.fc1 = nn.Linear(in_channels, widening_factor * in_channels)
        variance_scaling_(self.fc1.weight, scale=init_scale)
        nn.init.constant_(self.fc1.bias, 0)
        self.fc2 = nn.Linear(widening_factor * in_channels, out_channels)
        variance_scaling_(self.fc2.weight, scale=init_scale)
        nn.init.constant_(self.fc2.bias, 0)
        self._dropout_prob = dropout_prob
        self.dropout = nn.Dropout(dropout_prob)

    def forward(self, x):
        _check_inference(self, self._dropout_prob)
        if engine.PRECISION == "bf16x3":
            ops._need_cuda(x)
            return validate.mlp_module(self, x)
        pm = engine.prepared(self, "mlp", lambda: engine.PreparedMLP(self))
        shape = x.shape
        xb = ops.layernorm_bf16(x.contiguous().view(-1, shape[-1]), None, None, normalize=False)
        return engine.mlp_only(pm, xb).contiguous().view(*shape[:-1], -1)


class SelfAttention(nn.Module):
    """Self-attention block incl. dense block (reference: transformer_primitives.py:219-297)."""

    def __init__(self, in_channels: int, widening_factor: int = 4, dropout_prob: float = 0.0,
                 dropout_attn_prob: float = 0.0, num_heads: int = 8, att_init_scale: float = 1.0,
                 dense_init_scale: float = 1.0, qk_channels: int = None, v_channels: int = None):
        super().__init__()
        if qk_channels is None:
            qk_channels = in_channels
        if v_channels is None:
            v_channels = qk_channels
        self.mlp = MLP(in_channels=v_channels, widening_factor=widening_factor, dropout_prob=dropout_prob,
                       init_scale=dense_init_scale)
        self.attention = Attention(q_in_channels=in_channels, k_in_channels=in_channels, v_in_channels=in_channels,
                                   num_heads=num_heads, init_scale=att_init_scale, qk_out_channels=qk_channels,
                                   v_out_channels=v_channels, dropout_prob=dropout_attn_prob)
        self.layer_norm1 = nn.LayerNorm(in_channels)
        self.layer_norm2 = nn.LayerNorm(v_channels)
        self._dropout_probs = (dropout_prob, dropout_attn_prob)
        self.dropout = nn.Dropout(dropout_prob)

    def forward(self, inputs, *, attention_mask=None, attention_bias=None, return_matrix: bool = False):
        _check_inference(self, *self._dropout_probs)
        ops._need_cuda(inputs)
        B, N, _ = inputs.shape
        if B == 0 or N == 0:   # empty batch: nothing to launch
            y = inputs.new_empty(inputs.shape)
            return (inputs.new_empty(B, self.attention._num_heads, N, N), y) if return_matrix else y
        row_keep, key_mask, general = _route_mask(attention_mask, attention_bias, return_matrix, B=B,
                                                  H=self.attention._num_heads, Nq=N, Nk=N, device=inputs.device)
        if engine.PRECISION == "bf16x3":
            y = validate.self_attention_block(self, inputs.contiguous(), key_mask, row_keep, general)
            return (general.matrix, y) if return_matrix else y
        pa = engine.prepared(self.attention, "self", lambda: engine.PreparedAttention(self.attention,
                                                                                     self_attention=True,
                                                                                     allow_fold=False))
        pm = engine.prepared(self.mlp, "mlp", lambda: engine.PreparedMLP(self.mlp))
        x = inputs if inputs.is_contiguous() else inputs.contiguous()
        y = engine.self_attention_block(pa, pm, x, self.layer_norm1, self.layer_norm2,
                                        key_mask=engine._as_u8(key_mask), row_keep=engine._as_u8(row_keep),
                                        general=general)
        y = y if y.is_contiguous() else y.contiguous()
        return (general.matrix, y) if return_matrix else y


class CrossAttention(nn.Module):
    """Cross-attention block incl. dense block (reference: transformer_primitives.py:300-406)."""

    def __init__(self, q_in_channels: int, kv_in_channels: int, widening_factor: int = 1, dropout_prob: float = 0.0,
                 dropout_attn_prob: float = 0.0, num_heads: int = 8, attn_init_scale: float = 1.0,
                 mlp_init_scale: float = 1.0, shape_for_attn: str = "kv", use_query_residual: bool = True,
                 qk_channels: int = None, v_channels: int = None):
        super().__init__()
        self._use_query_residual = use_query_residual
        output_channels = q_in_channels
        if qk_channels is None:
            if shape_for_attn == "q":
                qk_channels = q_in_channels
            elif shape_for_attn == "kv":
                qk_channels = kv_in_channels
            else:
                raise ValueError(f"Unknown value {shape_for_attn} for "
                                 "shape_for_attention.")
        if v_channels is None:
            v_channels = qk_channels
        self.attention = Attention(q_in_channels=q_in_channels, k_in_channels=kv_in_channels,
                                   v_in_channels=kv_in_channels, num_heads=num_heads, init_scale=attn_init_scale,
                                   dropout_prob=dropout_attn_prob, qk_out_channels=qk_channels,
                                   v_out_channels=v_channels, output_channels=output_channels)
        self.mlp = MLP(in_channels=output_channels, widening_factor=widening_factor, dropout_prob=dropout_prob,
                       init_scale=mlp_init_scale)
        self.layer_norm_q = nn.LayerNorm(q_in_channels)
        self.layer_norm_kv = nn.LayerNorm(kv_in_channels)
        self.layer_norm2 = nn.LayerNorm(output_channels)
        self._dropout_probs = (dropout_prob, dropout_attn_prob)
        self.dropout = nn.Dropout(dropout_prob)

    def forward(self, inputs_q, inputs_kv, *, attention_mask=None, attention_bias=None, return_matrix: bool = False):
        ops._need_cuda(inputs_q, inputs_kv)
        if inputs_q.shape[0] == 0 or inputs_q.shape[1] == 0:   # empty batch / no queries: nothing to launch
            y = inputs_q.new_empty(inputs_q.shape)
            m = inputs_q.new_empty(inputs_q.shape[0], self.attention._num_heads, inputs_q.shape[1], inputs_kv.shape[1])
            return (m, y) if return_matrix else y
        row_keep, key_mask, general = _route_mask(attention_mask, attention_bias, return_matrix,
                                                  B=inputs_q.shape[0], H=self.attention._num_heads,
                                                  Nq=inputs_q.shape[1], Nk=inputs_kv.shape[1], device=inputs_q.device)
        y, _ = self._forward_factored(inputs_q, inputs_kv, key_mask=key_mask, row_keep=row_keep, general=general)
        y = y if y.is_contiguous() else y.contiguous()   # odd widths are carried with a 16-byte row pitch inside
        return (general.matrix, y) if return_matrix else y

    def _forward_factored(self, inputs_q, inputs_kv, *, key_mask=None, row_keep=None, want_bf16_out=False,
                          shard=None, stats_out=None, general=None, tail=None):
        _check_inference(self, *self._dropout_probs)
        if engine.PRECISION == "bf16x3":
            ops._need_cuda(inputs_q, inputs_kv)
            if shard is not None:
                raise RuntimeError("perceiverio_pytorch_b200: the validation precision runs unsharded")
            return validate.cross_attention_block(self, inputs_q, inputs_kv, key_mask=key_mask, row_keep=row_keep,
                                                  general=general), None
        if general is not None:   # explicit S / P path: no K/V folding
            pa = engine.prepared(self.attention, "plain", lambda: engine.PreparedAttention(self.attention,
                                                                                          self_attention=False,
                                                                                          allow_fold=False))
        else:
            pa = engine.prepared(self.attention, "cross", lambda: engine.PreparedAttention(self.attention,
                                                                                          self_attention=False,
                                                                                          allow_fold=True))
        pm = engine.prepared(self.mlp, "mlp", lambda: engine.PreparedMLP(self.mlp))
        return engine.cross_attention_block(pa, pm, inputs_q, inputs_kv, self.layer_norm_q, self.layer_norm_kv,
                                            self.layer_norm2, use_query_residual=self._use_query_residual,
                                            key_mask=key_mask, row_keep=row_keep, want_bf16_out=want_bf16_out,
                                            shard=shard, stats_out=stats_out, general=general, tail=tail)
