"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python oracle/make_golden.py
Each fixture stores the module's parameters (reference state_dict names), the seeded inputs and the
output of the unmodified reference module in fp32 on CPU.  Fixtures are kept tiny (tens of KB) so they
can be committed; they travel to the GPU box, the reference does not.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _build(cls, **kwargs):
    """Construct a reference module and remember the constructor kwargs (stored in the fixture so the
    B200 drop-in can be built with the very same call)."""
    m = cls(**kwargs).eval()
    m._ctor = dict(cls=cls.__name__, kwargs=kwargs)
    return m


ONLY = set(sys.argv[1:])   # optional: regenerate just the named fixtures


def _save(name, module, meta, inputs, output, **extra_outputs):
    if ONLY and name not in ONLY:
        return
    meta = dict(meta, ctor=json.dumps(module._ctor))
    arrays = {f"param::{k}": v.detach().numpy() for k, v in module.state_dict().items()}
    arrays.update({f"input::{k}": (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                   for k, v in inputs.items()})
    arrays["output"] = output.detach().numpy()
    arrays.update({f"output::{k}": v.detach().numpy() for k, v in extra_outputs.items()})
    arrays.update({f"meta::{k}": np.asarray(v) for k, v in meta.items()})
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, out absmax {float(output.abs().max()):.4f}")


def main():
    ref = ref_shim.load_reference()
    assert ref is not None, "reference tree not mounted"
    P, R = ref.primitives, ref.perceiver
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    with torch.inference_mode():
        # 1. single-head cross-attend with an odd channel count (the cls/flow encoder pattern)
        m = ref_shim.perturb_parameters(_build(P.CrossAttention, q_in_channels=64, kv_in_channels=37, num_heads=1), 1)
        q, kv = torch.randn(2, 48, 64), torch.randn(2, 300, 37)
        _save("xattn_h1_odd", m, dict(kind="cross", num_heads=1, use_query_residual=1),
              dict(q=q, kv=kv), m(q, kv))

        # 2. multi-head masked cross-attend, distinct qk / v widths (the language encoder pattern)
        m = ref_shim.perturb_parameters(_build(P.CrossAttention, q_in_channels=64, kv_in_channels=40, num_heads=4,
                                                         qk_channels=32, v_channels=80), 2)
        q, kv = torch.randn(2, 24, 64), torch.randn(2, 200, 40)
        kmask = torch.ones(2, 200, dtype=torch.bool)
        kmask[0, 150:] = False
        kmask[1, 3:] = False
        mask = P.make_cross_attention_mask(torch.ones(2, 24, dtype=torch.bool), kmask)
        _save("xattn_h4_keymask", m, dict(kind="cross", num_heads=4, use_query_residual=1),
              dict(q=q, kv=kv, key_mask=kmask), m(q, kv, attention_mask=mask))

        # 3. query-masked cross-attend without query residual (the language decoder pattern);
        #    sample 1 has every query masked -> rows become final.bias + MLP of it
        m = ref_shim.perturb_parameters(_build(P.CrossAttention, q_in_channels=48, kv_in_channels=64, num_heads=4,
                                                         qk_channels=32, v_channels=48,
                                                         use_query_residual=False), 3)
        q, kv = torch.randn(2, 130, 48), torch.randn(2, 20, 64)
        qmask = torch.ones(2, 130, dtype=torch.bool)
        qmask[0, 100:] = False
        qmask[1, :] = False
        mask = P.make_cross_attention_mask(qmask, torch.ones(2, 20, dtype=torch.bool))
        _save("xattn_h4_querymask", m, dict(kind="cross", num_heads=4, use_query_residual=0),
              dict(q=q, kv=kv, query_mask=qmask), m(q, kv, attention_mask=mask))

        # 4. self-attention block, 8 heads, qk != v (language tower pattern)
        m = ref_shim.perturb_parameters(_build(P.SelfAttention, in_channels=64, widening_factor=1, num_heads=8,
                                                        qk_channels=32, v_channels=64), 4)
        x = torch.randn(2, 40, 64)
        _save("selfattn_h8", m, dict(kind="self", num_heads=8), dict(x=x), m(x))

        # 5. peaky softmax: query projection scaled x16 so the running max moves
        m = ref_shim.perturb_parameters(_build(P.CrossAttention, q_in_channels=32, kv_in_channels=24, num_heads=1), 5)
        m.attention.proj_q.weight.mul_(16.0)
        q, kv = torch.randn(1, 16, 32), 2.0 * torch.randn(1, 700, 24)
        _save("xattn_peaky", m, dict(kind="cross", num_heads=1, use_query_residual=1),
              dict(q=q, kv=kv), m(q, kv))

        # 6. whole encoder: masked input, 2 blocks x 2 shared self-attends
        enc = ref_shim.perturb_parameters(_build(R.PerceiverEncoder, num_input_channels=37, num_self_attends_per_block=2,
                                                             num_blocks=2, num_latents=40, num_latent_channels=64,
                                                             num_cross_attend_heads=1, num_self_attend_heads=4), 6)
        x = torch.randn(2, 260, 37)
        imask = torch.ones(2, 260, dtype=torch.bool)
        imask[1, 200:] = False
        _save("encoder_small", enc, dict(kind="encoder", num_blocks=2, num_self_attends_per_block=2,
                                         num_cross_attend_heads=1, num_self_attend_heads=4, use_query_residual=1),
              dict(inputs=x, input_mask=imask), enc(x, enc.latents(x), input_mask=imask))

        # 7. whole decoder with final projection and a query residual
        dec = ref_shim.perturb_parameters(_build(R.PerceiverDecoder, query_channels=50, final_project_out_channels=10,
                                                             num_latent_channels=64, use_query_residual=True,
                                                             num_heads=1), 7)
        query, lat = torch.randn(2, 70, 50), torch.randn(2, 40, 64)
        _save("decoder_small", dec, dict(kind="decoder", num_heads=1, use_query_residual=1, final_project=1),
              dict(query=query, latents=lat), dec(query, lat))

        # 8. decoder without final projection, query mask (language decoder)
        dec = ref_shim.perturb_parameters(_build(R.PerceiverDecoder, query_channels=48, final_project_out_channels=48,
                                                             num_latent_channels=64, qk_channels=32, v_channels=48,
                                                             use_query_residual=False, num_heads=4,
                                                             final_project=False), 8)
        query, lat = torch.randn(2, 33, 48), torch.randn(2, 40, 64)
        qmask = torch.ones(2, 33, dtype=torch.bool)
        qmask[0, 20:] = False
        _save("decoder_querymask", dec, dict(kind="decoder", num_heads=4, use_query_residual=0, final_project=0),
              dict(query=query, latents=lat, query_mask=qmask), dec(query, lat, query_mask=qmask))

        # ---- general attention arguments (transformer_primitives.py:90, :143-144, :149-156, :177-178).  No recipe of
        # the reference passes them, but they are part of the operator interface.  Own seeds: adding cases never
        # changes the fixtures above.
        # 9. bare Attention: additive bias [B,H,Nq,Nk], dense (non-outer-product) mask with a fully masked row,
        #    return_matrix
        torch.manual_seed(109)
        m = ref_shim.perturb_parameters(_build(P.Attention, q_in_channels=32, k_in_channels=24, v_in_channels=24,
                                                num_heads=4, qk_out_channels=32, v_out_channels=48,
                                                output_channels=40), 9)
        q, kv = torch.randn(2, 20, 32), torch.randn(2, 50, 24)
        bias = torch.randn(2, 4, 20, 50)
        dmask = torch.rand(2, 20, 50) > 0.3
        dmask[0, 3, :] = False
        dmask[1, :, 45:] = False
        mat, out = m(q, kv, kv, attention_mask=dmask, attention_bias=bias, return_matrix=True)
        _save("attn_bias_densemask_matrix", m, dict(kind="attention", num_heads=4, return_matrix=1),
              dict(q=q, kv=kv, bias=bias, dense_mask=dmask), out, matrix=mat)

        # 10. SelfAttention: broadcast bias [1,1,N,N] (a relative-position table), causal mask, return_matrix
        torch.manual_seed(110)
        m = ref_shim.perturb_parameters(_build(P.SelfAttention, in_channels=64, widening_factor=1, num_heads=4), 10)
        x = torch.randn(2, 40, 64)
        bias = torch.randn(1, 1, 40, 40)
        causal = torch.tril(torch.ones(40, 40, dtype=torch.bool))[None].expand(2, 40, 40).contiguous()
        mat, out = m(x, attention_mask=causal, attention_bias=bias, return_matrix=True)
        _save("selfattn_bias_causal_matrix", m, dict(kind="self", num_heads=4, return_matrix=1),
              dict(x=x, bias=bias, dense_mask=causal), out, matrix=mat)

        # 11. CrossAttention: dense random mask only (single head, odd channel count), no matrix
        torch.manual_seed(111)
        m = ref_shim.perturb_parameters(_build(P.CrossAttention, q_in_channels=64, kv_in_channels=37, num_heads=1), 11)
        q, kv = torch.randn(2, 24, 64), torch.randn(2, 150, 37)
        dmask = torch.rand(2, 24, 150) > 0.5
        dmask[1, 7, :] = False
        _save("xattn_densemask", m, dict(kind="cross", num_heads=1, use_query_residual=1),
              dict(q=q, kv=kv, dense_mask=dmask), m(q, kv, attention_mask=dmask))

        # 12. CrossAttention: per-head bias broadcast over the batch [1,H,Nq,Nk], return_matrix, no mask
        torch.manual_seed(112)
        m = ref_shim.perturb_parameters(_build(P.CrossAttention, q_in_channels=48, kv_in_channels=64, num_heads=4,
                                                qk_channels=32, v_channels=48, use_query_residual=False), 12)
        q, kv = torch.randn(2, 30, 48), torch.randn(2, 70, 64)
        bias = 2.0 * torch.randn(1, 4, 30, 70)
        mat, out = m(q, kv, attention_bias=bias, return_matrix=True)
        _save("xattn_bias_matrix", m, dict(kind="cross", num_heads=4, use_query_residual=0, return_matrix=1),
              dict(q=q, kv=kv, bias=bias), out, matrix=mat)

        # ---- rows the reference wipes, on SINGLE-head blocks (the drop-in folds the K/V projections of those onto the
        # query side, which assumes rows of P that sum to one: wiped rows are the exception, transformer_primitives.py:168-175)
        # 13. single-head cross-attend with a query mask and a query residual
        torch.manual_seed(113)
        m = ref_shim.perturb_parameters(_build(P.CrossAttention, q_in_channels=48, kv_in_channels=40, num_heads=1), 13)
        for lin in (m.attention.proj_v, m.attention.final):
            lin.bias.mul_(20.0)      # make final.weight @ proj_v.bias (the term a wiped row must not receive) visible
        q, kv = torch.randn(2, 70, 48), torch.randn(2, 90, 40)
        qmask = torch.ones(2, 70, dtype=torch.bool)
        qmask[0, 50:] = False
        qmask[1, ::3] = False
        mask = P.make_cross_attention_mask(qmask, torch.ones(2, 90, dtype=torch.bool))
        _save("xattn_h1_querymask", m, dict(kind="cross", num_heads=1, use_query_residual=1),
              dict(q=q, kv=kv, query_mask=qmask), m(q, kv, attention_mask=mask))

        # 14. encoder with a single-head cross-attend where sample 1 has EVERY input masked
        torch.manual_seed(114)
        enc = ref_shim.perturb_parameters(_build(R.PerceiverEncoder, num_input_channels=37, num_self_attends_per_block=1,
                                                 num_blocks=1, num_latents=40, num_latent_channels=64,
                                                 num_cross_attend_heads=1, num_self_attend_heads=4), 14)
        enc.cross_attend.attention.proj_v.bias.mul_(20.0)
        x = torch.randn(2, 150, 37)
        imask = torch.ones(2, 150, dtype=torch.bool)
        imask[0, 100:] = False
        imask[1, :] = False
        _save("encoder_h1_sample_fully_masked", enc,
              dict(kind="encoder", num_blocks=1, num_self_attends_per_block=1, num_cross_attend_heads=1,
                   num_self_attend_heads=4, use_query_residual=1),
              dict(inputs=x, input_mask=imask), enc(x, enc.latents(x), input_mask=imask))

        # 15. single-head decoder (latent width <= 384) with a query mask, no query residual, final projection
        torch.manual_seed(115)
        dec = ref_shim.perturb_parameters(_build(R.PerceiverDecoder, query_channels=50, final_project_out_channels=10,
                                                 num_latent_channels=64, use_query_residual=False, num_heads=1), 15)
        dec.decoding_cross_attn.attention.proj_v.bias.mul_(20.0)
        query, lat = torch.randn(2, 70, 50), torch.randn(2, 40, 64)
        qmask = torch.ones(2, 70, dtype=torch.bool)
        qmask[0, 33:] = False
        _save("decoder_h1_querymask", dec, dict(kind="decoder", num_heads=1, use_query_residual=0, final_project=1),
              dict(query=query, latents=lat, query_mask=qmask), dec(query, lat, query_mask=qmask))


if __name__ == "__main__":
    main()
