"""Host-side weight algebra, checked on CPU in fp64 against the oracle (no kernels involved):
* the single-head K/V folding of `engine.PreparedAttention` — scores and outputs computed from the
  folded matrices equal the reference attention block;
* the LayerNorm folding of `engine.PreparedFusedLayer` (section 4.7) — rstd (x W'^T - mean colsum) + b' equals
  LayerNorm(x) W^T + b;
* the key-split chooser keeps every split non-empty."""
import math

import torch

import perceiverio_pytorch_b200 as pio
from perceiverio_pytorch_b200 import engine
from oracle import perceiver_oracle as O


def _perturb(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif "layer_norm" in name:
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))


def test_single_head_folding_reproduces_the_attention_block():
    torch.manual_seed(0)
    ca = pio.CrossAttention(q_in_channels=48, kv_in_channels=37, num_heads=1).double().eval()
    _perturb(ca, 1)
    pa = engine.PreparedAttention(ca.attention, self_attention=False, allow_fold=True)
    assert pa.folded
    # the folded matrices are stored as bf16 for the tensor cores; redo the same algebra in fp64 for the identity
    att = ca.attention
    wq, bq = att.proj_q.weight.detach(), att.proj_q.bias.detach()
    wk = att.proj_k.weight.detach()
    wv, bv = att.proj_v.weight.detach(), att.proj_v.bias.detach()
    wf, bf = att.final.weight.detach(), att.final.bias.detach()
    wq_fold, bq_fold = wk.t() @ wq, wk.t() @ bq
    wo_fold, bo_fold = wf @ wv, wf @ bv + bf
    assert torch.allclose(pa.wq_fold[:, :48].double(), wq_fold, atol=2e-2, rtol=2e-2)      # bf16 copies of the same
    assert torch.allclose(pa.wo_fold[:, :37].double(), wo_fold, atol=2e-2, rtol=2e-2)
    qn, kvn = torch.randn(2, 9, 48, dtype=torch.float64), torch.randn(2, 50, 37, dtype=torch.float64)
    # reference attention (oracle) on the already normalised inputs
    p = {k: v.detach() for k, v in att.state_dict().items()}
    want = O.attention(p, "", 1, qn, kvn, kvn)
    # folded: S = (qn Wq'^T + bq') . kvn^T  (the q.bk term is constant per row), out = (P kvn) Wo'^T + bo'
    qf = qn @ wq_fold.t() + bq_fold
    s = (qf @ kvn.transpose(1, 2)) / math.sqrt(att.proj_q.weight.shape[0])
    got = (torch.softmax(s, -1) @ kvn) @ wo_fold.t() + bo_fold
    assert torch.allclose(got, want, atol=1e-10, rtol=1e-10)


def test_layernorm_folding_reproduces_the_projections():
    torch.manual_seed(1)
    sa = pio.SelfAttention(in_channels=64, widening_factor=1, num_heads=4).double().eval()
    _perturb(sa, 2)
    pf = engine.PreparedFusedLayer(sa.float())
    sa = sa.double()
    x = torch.randn(10, 64, dtype=torch.float64) * 2 + 0.3
    mean, var = x.mean(-1, keepdim=True), x.var(-1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + sa.layer_norm1.eps)
    with torch.no_grad():
        ln = torch.nn.functional.layer_norm(x, (64,), sa.layer_norm1.weight, sa.layer_norm1.bias, sa.layer_norm1.eps)
        want = torch.cat([sa.attention.proj_q(ln), sa.attention.proj_k(ln), sa.attention.proj_v(ln)], -1)
    w = pf.wqkv[:, :64].double()            # W diag(gamma), bf16-rounded
    got = rstd * (x @ w.t() - mean * pf.cs_qkv.double()) + pf.bqkv.double()
    # only the bf16 rounding of W' separates the two (relative 2^-9 per weight)
    assert float((got - want).abs().max() / want.abs().max()) < 5e-3
    # with the un-rounded W' the identity is exact
    wq = torch.cat([sa.attention.proj_q.weight, sa.attention.proj_k.weight, sa.attention.proj_v.weight], 0).detach()
    bq = torch.cat([sa.attention.proj_q.bias, sa.attention.proj_k.bias, sa.attention.proj_v.bias], 0).detach()
    w_exact = wq * sa.layer_norm1.weight.detach()[None, :]
    exact = rstd * (x @ w_exact.t() - mean * w_exact.sum(1)) + (bq + wq @ sa.layer_norm1.bias.detach())
    assert torch.allclose(exact, want, atol=1e-10, rtol=1e-10)
    assert torch.allclose(pf.bqkv.double(), bq + wq @ sa.layer_norm1.bias.detach(), atol=1e-6)


def test_key_split_chooser_never_leaves_a_split_empty():
    for B, H, Nq, Nk in [(1, 1, 512, 50176), (64, 1, 512, 50176), (1, 1, 2048, 182528), (8, 1, 512, 1 << 20),
                         (1, 1, 256, 5000), (3, 1, 130, 4097)]:
        s = engine._pick_splits(B, H, Nq, Nk, 261, 261, True)
        bn = 64
        tiles = (Nk + bn - 1) // bn
        assert 1 <= s <= 32
        per = (tiles + s - 1) // s
        assert (s - 1) * per < tiles, (B, Nq, Nk, s)


def test_key_split_chooser_splits_short_inputs_only_when_the_query_tiles_leave_the_sms_idle():
    """Inputs of 1024 .. 4095 keys are split only when (batch x heads x query tiles) would occupy less than half of the SMs
    (the language encoder: 8 heads x 256 latents over 2048 bytes); towers with many query tiles and short self-attends
    keep the unsplit two-tile kernel."""
    from perceiverio_pytorch_b200 import _lib
    lang = engine._pick_splits(1, 8, 256, 2048, 32, 160, False)
    assert lang > 1
    bn = _lib.load().pio_attention_key_tile(32, 160, 0)
    tiles = (2048 + bn - 1) // bn
    assert tiles // lang >= 4 and (lang - 1) * ((tiles + lang - 1) // lang) < tiles
    assert engine._pick_splits(1, 16, 2048, 2048, 32, 32, False) == 1      # flow tower: 256 query tiles
    assert engine._pick_splits(1, 8, 784, 784, 64, 64, False) == 1         # multimodal tower: < 1024 keys
    assert engine._pick_splits(1, 8, 256, 256, 32, 160, False) == 1        # language tower
    assert engine._pick_splits(64, 8, 512, 512, 128, 128, False) == 1      # classification tower
    assert engine._pick_splits(1, 8, 2048, 256, 32, 96, False) == 1        # language decoder


def test_postprocessor_linear_composes_with_final_layer():
    """N3 host algebra: post(final(x)) as one affine map (perceiver._ComposedFinal), in fp64 against the two layers
    applied in sequence; the composed weights follow in-place parameter updates and stay out of the state_dict."""
    import perceiverio_pytorch_b200 as pio
    from perceiverio_pytorch_b200 import inputs as pin_mod
    torch.manual_seed(5)
    dec = pio.PerceiverDecoder(query_channels=48, final_project_out_channels=10, num_latent_channels=64).eval()
    post = torch.nn.Linear(10, 7)
    with torch.no_grad():
        dec.final_layer.bias.normal_()
        post.bias.normal_()
    keys = list(dec.state_dict().keys())
    fin, n_out = dec._final(post)
    assert n_out == 7 and list(dec.state_dict().keys()) == keys
    x = torch.randn(33, 48, dtype=torch.float64)
    want = post.double()(dec.final_layer.double()(x)).detach()
    got = x @ fin.weight.double().t() + fin.bias.double()
    assert float((got - want).abs().max()) <= 1e-6 * float(want.abs().max())
    with torch.no_grad():
        dec.final_layer.weight.mul_(0.5)
    want2 = post(dec.final_layer(x)).detach()
    got2 = x @ fin.weight.double().t() + fin.bias.double()
    assert float((got2 - want2).abs().max()) <= 1e-6 * float(want2.abs().max())
    assert dec._final(post)[0] is fin and dec._final(None)[0] is dec.final_layer
    assert not dec.fuses_post_linear(torch.nn.Linear(11, 3)) and dec.fuses_post_linear(post)
    no_proj = pio.PerceiverDecoder(query_channels=48, final_project_out_channels=10, num_latent_channels=64,
                                   final_project=False)
    assert not no_proj.fuses_post_linear(post)

    # which postprocessors qualify (duck-typed on the reference's class names, postprocessors.py:165-208)
    class ClassificationPostprocessor(torch.nn.Module):
        def __init__(self, project):
            super().__init__()
            self._project = project
            if project:
                self.linear = torch.nn.Linear(10, 7)

    class ProjectionPostprocessor(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.projection = torch.nn.Linear(10, 4)

    class FlowPostprocessor(torch.nn.Module):
        pass

    c = ClassificationPostprocessor(True)
    assert pin_mod._post_linear({"__default": c}, dec) is c.linear
    assert pin_mod._post_linear({"__default": ClassificationPostprocessor(False)}, dec) is None
    pr = ProjectionPostprocessor()
    assert pin_mod._post_linear({"__default": pr}, dec) is pr.projection
    assert pin_mod._post_linear({"__default": FlowPostprocessor()}, dec) is None
    assert pin_mod._post_linear({"a": c, "b": pr}, dec) is None and pin_mod._post_linear(None, dec) is None
    assert pin_mod._post_linear({"__default": c}, torch.nn.Linear(2, 2)) is None      # a decoder that is not ours
