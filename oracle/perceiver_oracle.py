"""CPU oracle for the Perceiver IO attention-stack forward.

TEST INFRASTRUCTURE ONLY.  This file is a plain-tensor restatement of the hot path of
JOBR0/PerceiverIO_Pytorch (`perceiver_io/transformer_primitives.py` and the encoder / decoder
drivers in `perceiver_io/perceiver.py`).  It exists so the CUDA path can be checked without the
reference being present (the GPU box has no /root/reference).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
it; the product package `perceiverio_pytorch_b200` never does.

Parity pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against *outputs of the reference itself run in the build container*
(`oracle/make_golden.py` imports the reference through `oracle/ref_shim.py` and writes
`tests/golden/*.npz`; `tests/test_oracle_vs_reference.py` re-runs the live comparison whenever
/root/reference is mounted, `tests/test_oracle_golden.py` checks the committed vectors everywhere).

Every function takes explicit tensors / a flat ``params`` mapping with the reference's state_dict key
names; nothing here is an ``nn.Module``.  All arithmetic is done in the dtype of the
inputs (fp32 by default, fp64 for tight checks).
"""
from __future__ import annotations

import math
from typing import Mapping, Optional

import torch

Tensor = torch.Tensor
LN_EPS = 1e-5  # torch.nn.LayerNorm default, used by transformer_primitives.py:270-271,365-367


def make_cross_attention_mask(query_mask: Tensor, kv_mask: Tensor) -> Tensor:
    """Outer product of two boolean masks -> [B, Nq, Nk] (transformer_primitives.py:10-15)."""
    b, nq = query_mask.shape
    b2, nk = kv_mask.shape
    assert b == b2
    return query_mask[:, :, None] & kv_mask[:, None, :] if query_mask.dtype == torch.bool \
        else query_mask[:, :, None] * kv_mask[:, None, :]


def layer_norm(x: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """LayerNorm over the channel axis, biased variance, eps 1e-5 (nn.LayerNorm semantics)."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + LN_EPS) * weight + bias


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """y = x W^T + b with W stored [out, in] (nn.Linear)."""
    y = x @ weight.t()
    return y if bias is None else y + bias


def gelu_erf(x: Tensor) -> Tensor:
    """Exact (erf) GELU, the F.gelu default used at transformer_primitives.py:214."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def attend(q: Tensor, k: Tensor, v: Tensor, attention_mask: Optional[Tensor] = None,
           attention_bias: Optional[Tensor] = None, return_matrix: bool = False):
    """Multi-head attention core (transformer_primitives.py:117-180).

    q [B,Nq,H,Dqk], k [B,Nk,H,Dqk], v [B,Nk,H,Dv]; mask [B,Nq,Nk] (True = attend); bias broadcastable to
    [B,H,Nq,Nk].  Order of operations follows the reference: logits, additive bias BEFORE the scale (:143-144), the
    1/sqrt(Dqk) scale (:146-147), masked positions replaced by -1e30 (:149-156), softmax (:158), P@V (:163),
    head-major merge (:164-166) and finally rows whose mask row is entirely False are forced to zero (:168-175).
    With return_matrix the un-wiped probabilities [B,H,Nq,Nk] are returned first (:177-178): a fully masked row is
    uniform there.
    """
    b, nq, h, dqk = q.shape
    dv = v.shape[-1]
    qh = q.permute(0, 2, 1, 3)
    kh = k.permute(0, 2, 1, 3)
    vh = v.permute(0, 2, 1, 3)
    logits = qh @ kh.transpose(-2, -1)
    if attention_bias is not None:
        logits = logits + attention_bias
    logits = logits * (1.0 / math.sqrt(dqk))
    if attention_mask is not None:
        m = attention_mask.to(torch.bool)[:, None, :, :]
        logits = torch.where(m, logits, torch.full((), -1e30, dtype=logits.dtype))
    probs = torch.softmax(logits, dim=-1)
    out = (probs @ vh).permute(0, 2, 1, 3).reshape(b, nq, h * dv)
    if attention_mask is not None:
        wipe = ~(attention_mask.to(torch.bool).any(dim=2, keepdim=True))
        out = torch.where(wipe, torch.zeros((), dtype=out.dtype), out)
    if return_matrix:
        return probs, out
    return out


def attention(p: Mapping[str, Tensor], prefix: str, num_heads: int, xq: Tensor, xk: Tensor, xv: Tensor,
              attention_mask: Optional[Tensor] = None, attention_bias: Optional[Tensor] = None,
              return_matrix: bool = False):
    """`Attention.forward` (transformer_primitives.py:90-115): q/k/v Linear, attend, final Linear."""
    q = linear(xq, p[prefix + "proj_q.weight"], p[prefix + "proj_q.bias"])
    k = linear(xk, p[prefix + "proj_k.weight"], p[prefix + "proj_k.bias"])
    v = linear(xv, p[prefix + "proj_v.weight"], p[prefix + "proj_v.bias"])
    b, nq, qk = q.shape
    nk = k.shape[1]
    vc = v.shape[-1]
    q = q.reshape(b, nq, num_heads, qk // num_heads)
    k = k.reshape(b, nk, num_heads, qk // num_heads)
    v = v.reshape(b, nk, num_heads, vc // num_heads)
    o = attend(q, k, v, attention_mask, attention_bias, return_matrix)
    matrix = None
    if return_matrix:
        matrix, o = o
    y = linear(o, p[prefix + "final.weight"], p.get(prefix + "final.bias"))
    return (matrix, y) if return_matrix else y


def mlp(p: Mapping[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """`MLP.forward` (transformer_primitives.py:212-216): fc2(gelu(fc1(x)))."""
    h = gelu_erf(linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    return linear(h, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def self_attention(p: Mapping[str, Tensor], prefix: str, num_heads: int, x: Tensor,
                   attention_mask: Optional[Tensor] = None, attention_bias: Optional[Tensor] = None,
                   return_matrix: bool = False):
    """`SelfAttention.forward` (transformer_primitives.py:275-297)."""
    xn = layer_norm(x, p[prefix + "layer_norm1.weight"], p[prefix + "layer_norm1.bias"])
    a = attention(p, prefix + "attention.", num_heads, xn, xn, xn, attention_mask, attention_bias, return_matrix)
    matrix = None
    if return_matrix:
        matrix, a = a
    x = x + a
    xn2 = layer_norm(x, p[prefix + "layer_norm2.weight"], p[prefix + "layer_norm2.bias"])
    y = x + mlp(p, prefix + "mlp.", xn2)
    return (matrix, y) if return_matrix else y


def cross_attention(p: Mapping[str, Tensor], prefix: str, num_heads: int, use_query_residual: bool,
                    inputs_q: Tensor, inputs_kv: Tensor, attention_mask: Optional[Tensor] = None,
                    attention_bias: Optional[Tensor] = None, return_matrix: bool = False):
    """`CrossAttention.forward` (transformer_primitives.py:371-406).

    The query residual uses the un-normalised queries (:396-399)."""
    kvn = layer_norm(inputs_kv, p[prefix + "layer_norm_kv.weight"], p[prefix + "layer_norm_kv.bias"])
    qn = layer_norm(inputs_q, p[prefix + "layer_norm_q.weight"], p[prefix + "layer_norm_q.bias"])
    a = attention(p, prefix + "attention.", num_heads, qn, kvn, kvn, attention_mask, attention_bias, return_matrix)
    matrix = None
    if return_matrix:
        matrix, a = a
    x = inputs_q + a if use_query_residual else a
    xn2 = layer_norm(x, p[prefix + "layer_norm2.weight"], p[prefix + "layer_norm2.bias"])
    y = x + mlp(p, prefix + "mlp.", xn2)
    return (matrix, y) if return_matrix else y


def encoder_latents(p: Mapping[str, Tensor], prefix: str, batch: int) -> Tensor:
    """`PerceiverEncoder.latents` (perceiver.py:94-96): the latent array broadcast over the batch."""
    z = p[prefix + "latent_pos_enc.pos_embs"]
    return z[None].expand(batch, *z.shape)


def encoder_forward(p: Mapping[str, Tensor], prefix: str, *, num_blocks: int, num_self_attends_per_block: int,
                    num_cross_attend_heads: int, num_self_attend_heads: int, use_query_residual: bool = True,
                    inputs: Tensor, latents: Optional[Tensor] = None, input_mask: Optional[Tensor] = None,
                    stop_after_cross_attend: bool = False) -> Tensor:
    """`PerceiverEncoder.forward` (perceiver.py:98-107): one cross-attend, then num_blocks x the shared
    list of self-attends."""
    if latents is None:
        latents = encoder_latents(p, prefix, inputs.shape[0])
    mask = None
    if input_mask is not None:
        ones = torch.ones(latents.shape[:2], dtype=torch.bool)
        mask = make_cross_attention_mask(ones, input_mask.to(torch.bool))
    z = cross_attention(p, prefix + "cross_attend.", num_cross_attend_heads, use_query_residual,
                        latents, inputs, mask)
    if stop_after_cross_attend:
        return z
    for _ in range(num_blocks):
        for i in range(num_self_attends_per_block):
            z = self_attention(p, f"{prefix}self_attends.{i}.", num_self_attend_heads, z)
    return z


def decoder_forward(p: Mapping[str, Tensor], prefix: str, *, num_heads: int, use_query_residual: bool,
                    final_project: bool, query: Tensor, latents: Tensor,
                    query_mask: Optional[Tensor] = None) -> Tensor:
    """`PerceiverDecoder.forward` (perceiver.py:166-180)."""
    mask = None
    if query_mask is not None:
        ones = torch.ones(latents.shape[:2], dtype=torch.bool)
        mask = make_cross_attention_mask(query_mask.to(torch.bool), ones)
    out = cross_attention(p, prefix + "decoding_cross_attn.", num_heads, use_query_residual,
                          query, latents, mask)
    if final_project:
        out = linear(out, p[prefix + "final_layer.weight"], p[prefix + "final_layer.bias"])
    return out


# ---------------------------------------------------------------------------------------------
# Input-side glue of the pixels recipe: what the reference's ImagePreprocessor builds as the
# encoder input.  Restated so that the CPU baseline covers the same boundary as the fused CUDA input path.
# ---------------------------------------------------------------------------------------------

def linear_positions(index_dims) -> Tensor:
    """`build_linear_positions` (position_encoding.py:70-89): grid in [-1, 1]^d, shape [prod(index_dims), d]."""
    ranges = [torch.linspace(-1.0, 1.0, steps=int(n), dtype=torch.float32) for n in index_dims]
    grid = torch.meshgrid(*ranges, indexing="ij")
    return torch.stack(grid, dim=-1).reshape(-1, len(index_dims))


def fourier_features(pos: Tensor, num_bands: int, max_resolution, concat_pos: bool = True,
                     sine_only: bool = False) -> Tensor:
    """`generate_fourier_features` (position_encoding.py:19-67): [n, d] positions -> [n, d (+ 2) * ...] features ordered
    [pos, sin(pi f x) for every dim and band, cos(pi f x) ...]; bands are linspace(1, res / 2, num_bands)."""
    freq = torch.stack([torch.linspace(1.0, res / 2, steps=num_bands) for res in max_resolution], dim=0)
    per_pos = (pos[:, :, None] * freq[None, :, :]).reshape(pos.shape[0], -1)
    if sine_only:
        feats = torch.sin(math.pi * per_pos)
    else:
        feats = torch.cat([torch.sin(math.pi * per_pos), torch.cos(math.pi * per_pos)], dim=-1)
    return torch.cat([pos, feats], dim=-1) if concat_pos else feats


def image_inputs_pixels(images: Tensor, num_bands: int = 64, max_resolution=(224, 224),
                        spatial_downsample: int = 1) -> Tensor:
    """`ImagePreprocessor(prep_type="pixels", concat)` (io_processors/preprocessors.py:239-258 and :180-199): channels
    last, crude strided downsampling, flatten the index dims, concatenate the (batch-broadcast) Fourier features.
    images [B, C, H, W] -> [B, H' * W', C + n_pos]."""
    x = images.movedim(-3, -1)[:, ::spatial_downsample, ::spatial_downsample]
    b, h, w, c = x.shape
    feats = x.reshape(b, h * w, c)
    pos = fourier_features(linear_positions((h, w)), num_bands, max_resolution)
    return torch.cat([feats, pos[None].expand(b, -1, -1).to(feats.dtype)], dim=-1)


# ---------------------------------------------------------------------------------------------
# Key-axis sharding identity used by the multi-GPU encoder path.  The oracle states it
# on the CPU so the gloo tests can check the host-side combine.
# ---------------------------------------------------------------------------------------------

def attend_partial(q: Tensor, k: Tensor, v: Tensor, key_mask: Optional[Tensor] = None):
    """Un-normalised attention over one key shard: returns (O_unnorm [B,H,Nq,Dv], m [B,H,Nq], l [B,H,Nq]).

    Masked keys contribute nothing; a shard with no valid key returns m=-inf, l=0, O=0."""
    dqk = q.shape[-1]
    qh, kh, vh = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    s = (qh @ kh.transpose(-2, -1)) * (1.0 / math.sqrt(dqk))
    if key_mask is not None:
        s = torch.where(key_mask.to(torch.bool)[:, None, None, :], s, torch.full((), float("-inf"), dtype=s.dtype))
    m = s.max(dim=-1).values
    m_safe = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    pexp = torch.exp(s - m_safe[..., None])
    return pexp @ vh, m, pexp.sum(dim=-1)


def combine_partials(parts):
    """Merge [(O, m, l), ...] from disjoint key shards into the normalised attention output
    [B,H,Nq,Dv]; rows with no valid key anywhere come out as zeros."""
    m_all = torch.stack([p[1] for p in parts]).max(dim=0).values
    m_safe = torch.where(torch.isinf(m_all), torch.zeros_like(m_all), m_all)
    o = torch.zeros_like(parts[0][0])
    l = torch.zeros_like(parts[0][2])
    for (oi, mi, li) in parts:
        w = torch.exp(torch.where(torch.isinf(mi), torch.full_like(mi, float("-inf")), mi) - m_safe)
        o = o + oi * w[..., None]
        l = l + li * w
    return torch.where(l[..., None] > 0, o / l[..., None].clamp_min(1e-38), torch.zeros_like(o))


# ---------------------------------------------------------------------------------------------
# FLOP model: reference-algorithm FLOPs of one attention block and of a whole model.
# ---------------------------------------------------------------------------------------------

def attn_block_flops(nq, nk, cq, ck, qk, v, o, hidden=None):
    hidden = o if hidden is None else hidden
    return (2 * nq * cq * qk + 2 * nk * ck * (qk + v) + 2 * nq * nk * (qk + v) + 2 * nq * v * o
            + 2 * nq * (o * hidden + hidden * o))
