"""GPU parity tests proper: the CUDA path (through the C ABI) vs the committed reference golden vectors and vs
the CPU oracle on seeded inputs.  Tolerances (BASELINE.json north_star): bf16 path max|d|/max|ref| <= 1e-2."""
import json

import pytest
import torch

from golden_util import golden_names, load_golden, load_golden_matrix, rel_err

pytestmark = pytest.mark.gpu

BF16_TOL = 1e-2
# xattn_peaky scales proj_q x16 so logits reach |75|: bf16 operand rounding (2^-9 relative) alone moves logits by
# ~0.15, i.e. probabilities by ~15 %.  A CPU emulation of the *reference algorithm* with bf16-rounded operands gives
# 2.1e-2 on this case (unfolded) / 2.08e-2 (folded), and the kernel reproduces that, so the case is held to 3e-2:
# it exists to exercise the online-softmax rescale path, not the 1e-2 budget of the real configs.
CASE_TOL = {"xattn_peaky": 3e-2}


def _build_ours(meta):
    import perceiverio_pytorch_b200 as pio
    ctor = json.loads(meta["ctor"])
    return getattr(pio, ctor["cls"])(**ctor["kwargs"]).eval()


def _run_ours(module, inputs, meta):
    import perceiverio_pytorch_b200 as pio
    x = {k: v.cuda() for k, v in inputs.items()}
    kind = meta["kind"]
    general = {}
    if "bias" in x:
        general["attention_bias"] = x["bias"]
    if meta.get("return_matrix", 0):
        general["return_matrix"] = True
    with torch.inference_mode():
        if kind == "attention":
            return module(x["q"], x["kv"], x["kv"], attention_mask=x.get("dense_mask"), **general)
        if kind == "cross":
            b, nq, nk = x["q"].shape[0], x["q"].shape[1], x["kv"].shape[1]
            mask = x.get("dense_mask")
            if "key_mask" in x:
                mask = pio.make_cross_attention_mask(torch.ones(b, nq, dtype=torch.bool, device="cuda"), x["key_mask"])
            if "query_mask" in x:
                mask = pio.make_cross_attention_mask(x["query_mask"], torch.ones(b, nk, dtype=torch.bool, device="cuda"))
            return module(x["q"], x["kv"], attention_mask=mask, **general)
        if kind == "self":
            return module(x["x"], attention_mask=x.get("dense_mask"), **general)
        if kind == "encoder":
            return module(x["inputs"], module.latents(x["inputs"]), input_mask=x.get("input_mask"))
        if kind == "decoder":
            return module(x["query"], x["latents"], query_mask=x.get("query_mask"))
    raise ValueError(kind)


@pytest.mark.parametrize("name", golden_names())
def test_cuda_path_matches_reference_golden(name):
    params, inputs, meta, expected = load_golden(name)
    m = _build_ours(meta)
    m.load_state_dict(params, strict=True)
    m = m.cuda()
    got = _run_ours(m, inputs, meta)
    expected_matrix = load_golden_matrix(name)
    if expected_matrix is not None:     # return_matrix fixtures: (probabilities [B,H,Nq,Nk], block output)
        matrix, got = got
        matrix = matrix.float().cpu()
        assert matrix.shape == expected_matrix.shape
        # probabilities are compared absolutely (they live in [0, 1]; bf16 operand rounding moves logits by ~1e-2)
        assert float((matrix - expected_matrix).abs().max()) <= 2e-2, name
        assert float((matrix.sum(-1) - 1).abs().max()) <= 1e-4
    got = got.float().cpu()
    assert got.shape == expected.shape
    emax, el2 = rel_err(got, expected)
    assert emax <= CASE_TOL.get(name, BF16_TOL), (name, emax, el2)


@pytest.mark.parametrize("name", golden_names())
def test_cuda_path_matches_reference_golden_with_fp16_operands(name):
    """The same fixtures in the fp16 operand mode (engine.set_precision("fp16") / module.precision): 8x less operand
    rounding than bf16, so the bound tightens from 1e-2 to 2e-3."""
    from perceiverio_pytorch_b200 import engine
    params, inputs, meta, expected = load_golden(name)
    m = _build_ours(meta)
    m.load_state_dict(params, strict=True)
    m = m.cuda()
    with engine.precision_scope("fp16"):
        got = _run_ours(m, inputs, meta)
    assert engine.PRECISION == "bf16"
    expected_matrix = load_golden_matrix(name)
    if expected_matrix is not None:
        matrix, got = got
        assert float((matrix.float().cpu() - expected_matrix).abs().max()) <= 4e-3, name
    emax, el2 = rel_err(got.float().cpu(), expected)
    assert emax <= (6e-3 if name == "xattn_peaky" else 2e-3), (name, emax, el2)


def test_dense_mask_without_factors_is_factored():
    """A dense outer-product mask built by the caller (not via our helper) is still honoured."""
    import perceiverio_pytorch_b200 as pio
    params, inputs, meta, expected = load_golden("xattn_h4_keymask")
    m = _build_ours(meta)
    m.load_state_dict(params, strict=True)
    m = m.cuda()
    q, kv, km = inputs["q"].cuda(), inputs["kv"].cuda(), inputs["key_mask"].cuda()
    dense = torch.ones(q.shape[0], q.shape[1], 1, dtype=torch.bool, device="cuda") & km[:, None, :]
    with torch.inference_mode():
        got = m(q, kv, attention_mask=dense).cpu()
    assert rel_err(got, expected)[0] <= BF16_TOL


def test_cpu_tensors_fail_loudly():
    import perceiverio_pytorch_b200 as pio
    m = pio.SelfAttention(in_channels=64, widening_factor=1, num_heads=8).eval()
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 8, 64))


def _oracle_enc_dec(enc, dec, enc_cfg, dec_cfg, inputs, query, input_mask=None, query_mask=None):
    from oracle import perceiver_oracle as O
    pe = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    pd = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
    z = O.encoder_forward(pe, "", inputs=inputs, input_mask=input_mask, **enc_cfg)
    out = O.decoder_forward(pd, "", query=query, latents=z, query_mask=query_mask, **dec_cfg)
    return z, out


def _perturb(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if name.endswith("bias"):
                prm.copy_((0.1 if "layer_norm" in name else 0.02) * torch.randn(prm.shape, generator=g))
            elif "layer_norm" in name:
                prm.copy_(1.0 + 0.1 * torch.randn(prm.shape, generator=g))


CONFIGS = {
    # reduced-depth versions of the BASELINE.json configs (same widths / head structure, fewer layers and tokens so
    # the CPU oracle finishes in seconds)
    "language": dict(enc=dict(num_input_channels=768, num_self_attends_per_block=3, num_blocks=1, num_latents=256,
                              num_latent_channels=1280, qk_channels=256, v_channels=1280, num_cross_attend_heads=8,
                              num_self_attend_heads=8),
                     dec=dict(query_channels=768, final_project_out_channels=768, num_latent_channels=1280,
                              qk_channels=256, v_channels=768, num_heads=8, use_query_residual=False,
                              final_project=False),
                     B=1, Nk=2048, Nq=2048, masks=True),
    "classification": dict(enc=dict(num_input_channels=261, num_self_attends_per_block=2, num_blocks=2,
                                    num_latents=512, num_latent_channels=1024),
                           dec=dict(query_channels=1024, final_project_out_channels=1000, num_latent_channels=1024,
                                    use_query_residual=True),
                           B=2, Nk=6000, Nq=1000, masks=False),
    "flow": dict(enc=dict(num_input_channels=322, num_self_attends_per_block=2, num_blocks=1, num_latents=2048,
                          num_latent_channels=512, num_self_attend_heads=16),
                 dec=dict(query_channels=322, final_project_out_channels=2, num_latent_channels=512,
                          use_query_residual=False),
                 B=1, Nk=5000, Nq=5000, masks=False),
    "multimodal": dict(enc=dict(num_input_channels=704, num_self_attends_per_block=2, num_blocks=1, num_latents=784,
                                num_latent_channels=512),
                       dec=dict(query_channels=1026, final_project_out_channels=512, num_latent_channels=512,
                                use_query_residual=False),
                       B=1, Nk=3000, Nq=1200, masks=False),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_config_shapes_match_oracle(name):
    import perceiverio_pytorch_b200 as pio
    cfg = CONFIGS[name]
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(**cfg["enc"]).eval()
    dec = pio.PerceiverDecoder(**cfg["dec"]).eval()
    _perturb(enc, 1)
    _perturb(dec, 2)
    B, Nk, Nq = cfg["B"], cfg["Nk"], cfg["Nq"]
    inputs = torch.randn(B, Nk, cfg["enc"]["num_input_channels"])
    query = torch.randn(B, Nq, cfg["dec"]["query_channels"])
    imask = qmask = None
    if cfg["masks"]:
        imask = torch.zeros(B, Nk, dtype=torch.bool)
        imask[:, :1500] = True
        qmask = imask[:, :Nq].clone()
    e = cfg["enc"]
    enc_cfg = dict(num_blocks=e["num_blocks"], num_self_attends_per_block=e["num_self_attends_per_block"],
                   num_cross_attend_heads=e.get("num_cross_attend_heads", 1),
                   num_self_attend_heads=e.get("num_self_attend_heads", 8), use_query_residual=True)
    d = cfg["dec"]
    dec_cfg = dict(num_heads=d.get("num_heads", 1), use_query_residual=d["use_query_residual"],
                   final_project=d.get("final_project", True))
    z_ref, out_ref = _oracle_enc_dec(enc, dec, enc_cfg, dec_cfg, inputs, query, imask, qmask)
    enc, dec = enc.cuda(), dec.cuda()
    with torch.inference_mode():
        xi = inputs.cuda()
        z = enc(xi, enc.latents(xi), input_mask=imask.cuda() if imask is not None else None)
        out = dec(query.cuda(), z, query_mask=qmask.cuda() if qmask is not None else None)
    ez = rel_err(z.cpu(), z_ref)
    eo = rel_err(out.cpu(), out_ref)
    print(f"{name}: latents max {ez[0]:.3e} l2 {ez[1]:.3e}; output max {eo[0]:.3e} l2 {eo[1]:.3e}")
    assert ez[0] <= BF16_TOL and eo[0] <= BF16_TOL, (ez, eo)


def test_key_sharded_encoder_matches_unsharded_single_rank():
    """The key-axis shard path (partial attention -> packed exchange -> LSE combine) with world size 1 and two local
    key splits must reproduce the plain path and the oracle (the collective itself is covered by the gloo tests and by
    tools/encoder_sweep.py --check under torchrun)."""
    import perceiverio_pytorch_b200 as pio
    from perceiverio_pytorch_b200 import parallel
    from oracle import perceiver_oracle as O
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(num_input_channels=261, num_self_attends_per_block=1, num_blocks=1, num_latents=512,
                               num_latent_channels=1024).eval()
    _perturb(enc, 3)
    B, Nk = 2, 9000 + 37
    x = torch.randn(B, Nk, 261)
    mask = torch.rand(B, Nk) > 0.2
    mask[1, Nk // 2:] = False
    p = {k: v.detach() for k, v in enc.state_dict().items()}
    ref = O.encoder_forward(p, "", num_blocks=1, num_self_attends_per_block=1, num_cross_attend_heads=1,
                            num_self_attend_heads=8, use_query_residual=True, inputs=x, input_mask=mask)
    enc = enc.cuda()
    with torch.inference_mode():
        xc, mc = x.cuda(), mask.cuda()
        plain = enc(xc, enc.latents(xc), input_mask=mc)
        parallel.shard_encoder_keys(enc, group=None, local_splits=2)
        sharded = enc(xc, enc.latents(xc), input_mask=mc)
    assert rel_err(sharded.cpu(), plain.cpu())[0] <= 5e-3
    assert rel_err(sharded.cpu(), ref)[0] <= BF16_TOL
    # a sample without ANY valid key: the wipe comes out of the merged partial sums (pio_combine_args.row_alive), not
    # out of a reduction of the mask over the ranks
    params, inputs, meta, expected = load_golden("encoder_h1_sample_fully_masked")
    m = _build_ours(meta)
    m.load_state_dict(params, strict=True)
    m = m.cuda()
    parallel.shard_encoder_keys(m, group=None, local_splits=2)
    with torch.inference_mode():
        got = m(inputs["inputs"].cuda(), m.latents(inputs["inputs"].cuda()), input_mask=inputs["input_mask"].cuda())
    assert rel_err(got.cpu(), expected)[0] <= BF16_TOL


# ---------------------------------------------------------------------------------------------------------------
# Full-size checks through size-independent properties (the CPU oracle would need minutes and tens of GB there):
# the hot path carries no positional state of its own (positions are input features), so
#   * the encoder is invariant under a permutation of the input (key) axis,
#   * the decoder is equivariant under a permutation of the query axis,
#   * samples of a batch do not interact.
# These hold exactly in exact arithmetic; on the bf16 path a permutation only changes the summation order inside
# P.V and the tile / split a key lands in, so the bound is a fraction of the 1e-2 budget.
# ---------------------------------------------------------------------------------------------------------------
FULL = {
    "classification_pixels": dict(C=261, Nk=50176, lat=512, ch=1024, heads=8, B=2),
    "flow": dict(C=322, Nk=182528, lat=2048, ch=512, heads=16, B=1),
    "multimodal": dict(C=704, Nk=52097, lat=784, ch=512, heads=8, B=1),
}


@pytest.mark.parametrize("name", sorted(FULL))
def test_full_size_encoder_is_key_permutation_invariant(name):
    import perceiverio_pytorch_b200 as pio
    f = FULL[name]
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(num_input_channels=f["C"], num_self_attends_per_block=1, num_blocks=1,
                               num_latents=f["lat"], num_latent_channels=f["ch"],
                               num_self_attend_heads=f["heads"]).eval()
    _perturb(enc, 5)
    enc = enc.cuda()
    x = torch.randn(f["B"], f["Nk"], f["C"], device="cuda")
    perm = torch.randperm(f["Nk"], device="cuda")
    with torch.inference_mode():
        z0 = enc(x, enc.latents(x))
        z1 = enc(x[:, perm].contiguous(), enc.latents(x))
        # batch independence: sample 0 alone gives the same latents as sample 0 inside the batch
        z2 = enc(x[:1].contiguous(), enc.latents(x[:1]))
    assert torch.isfinite(z0).all()
    assert rel_err(z1, z0)[0] <= 4e-3, rel_err(z1, z0)
    assert rel_err(z2, z0[:1])[0] <= 4e-3, rel_err(z2, z0[:1])


def test_full_size_flow_decoder_is_query_permutation_equivariant():
    """182,528 output queries x 322 channels attending over 2048 x 512 latents (the optical-flow decoder)."""
    import perceiverio_pytorch_b200 as pio
    torch.manual_seed(0)
    dec = pio.PerceiverDecoder(query_channels=322, final_project_out_channels=2, num_latent_channels=512,
                               use_query_residual=False).eval()
    _perturb(dec, 6)
    dec = dec.cuda()
    nq = 182528
    query = torch.randn(1, nq, 322, device="cuda")
    lat = torch.randn(1, 2048, 512, device="cuda")
    perm = torch.randperm(nq, device="cuda")
    with torch.inference_mode():
        y0 = dec(query, lat)
        y1 = dec(query[:, perm].contiguous(), lat)
    assert y0.shape == (1, nq, 2)
    assert rel_err(y1, y0[:, perm])[0] <= 2e-3, rel_err(y1, y0[:, perm])


def test_full_depth_language_config_matches_oracle():
    """BASELINE.json configs[0] at full size: 2048 UTF-8 bytes, 256 latents x 1280 channels, 26 self-attends, masks."""
    import perceiverio_pytorch_b200 as pio
    cfg = CONFIGS["language"]
    e = dict(cfg["enc"], num_self_attends_per_block=26)
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(**e).eval()
    dec = pio.PerceiverDecoder(**cfg["dec"]).eval()
    _perturb(enc, 1)
    _perturb(dec, 2)
    inputs = torch.randn(1, 2048, 768)
    query = torch.randn(1, 2048, 768)
    imask = torch.zeros(1, 2048, dtype=torch.bool)
    imask[:, :1500] = True
    enc_cfg = dict(num_blocks=1, num_self_attends_per_block=26, num_cross_attend_heads=8, num_self_attend_heads=8,
                   use_query_residual=True)
    dec_cfg = dict(num_heads=8, use_query_residual=False, final_project=False)
    z_ref, out_ref = _oracle_enc_dec(enc, dec, enc_cfg, dec_cfg, inputs, query, imask, imask.clone())
    enc, dec = enc.cuda(), dec.cuda()
    with torch.inference_mode():
        xi = inputs.cuda()
        z = enc(xi, enc.latents(xi), input_mask=imask.cuda())
        out = dec(query.cuda(), z, query_mask=imask.cuda())
    ez, eo = rel_err(z.cpu(), z_ref), rel_err(out.cpu(), out_ref)
    print(f"language full depth: latents max {ez[0]:.3e} l2 {ez[1]:.3e}; output max {eo[0]:.3e} l2 {eo[1]:.3e}")
    assert ez[0] <= BF16_TOL and eo[0] <= BF16_TOL, (ez, eo)


def test_fused_layernorm_tower_matches_oracle_and_unfused_path():
    """Latent arrays of >= 2048 rows take the tower with the LayerNorms folded into the projections (no LayerNorm
    kernel, statistics accumulated by the producing GEMMs): same 1e-2 bound against the oracle, and close to the
    unfused CUDA path."""
    import perceiverio_pytorch_b200 as pio
    from perceiverio_pytorch_b200 import engine
    cfg = CONFIGS["classification"]
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(**dict(cfg["enc"], num_self_attends_per_block=3, num_blocks=2)).eval()
    _perturb(enc, 7)
    B, Nk = 8, 3000
    inputs = torch.randn(B, Nk, 261)
    pe = {k: v.detach() for k, v in enc.state_dict().items()}
    from oracle import perceiver_oracle as O
    z_ref = O.encoder_forward(pe, "", num_blocks=2, num_self_attends_per_block=3, num_cross_attend_heads=1,
                              num_self_attend_heads=8, use_query_residual=True, inputs=inputs)
    enc = enc.cuda()
    x = inputs.cuda()
    assert B * 512 >= engine.FUSE_LN_MIN_ROWS
    with torch.inference_mode():
        n0 = pio._lib.launch_count() if hasattr(pio, "_lib") else None
        z_fused = enc(x, enc.latents(x))
        engine.FUSE_LN = False
        try:
            z_plain = enc(x, enc.latents(x))
        finally:
            engine.FUSE_LN = True
    ef, ep = rel_err(z_fused.cpu(), z_ref), rel_err(z_plain.cpu(), z_ref)
    print(f"fused tower: max {ef[0]:.3e} l2 {ef[1]:.3e}; unfused: max {ep[0]:.3e} l2 {ep[1]:.3e}; "
          f"fused vs unfused {rel_err(z_fused, z_plain)[0]:.3e}")
    assert ef[0] <= BF16_TOL, ef
    # two independent bf16 evaluations, each ~5e-3 from the oracle: their mutual distance is bounded by the sum
    assert rel_err(z_fused, z_plain)[0] <= 8e-3


def test_encode_once_latent_cache():
    """cache_latents: a second call with equal (but re-built) inputs returns the cached latents without launching the
    encoder; different inputs, an in-place change or a parameter update invalidate it."""
    import perceiverio_pytorch_b200 as pio
    from perceiverio_pytorch_b200 import _lib
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(num_input_channels=64, num_self_attends_per_block=1, num_blocks=1, num_latents=128,
                               num_latent_channels=256).eval().cuda()
    enc.cache_latents = True
    x = torch.randn(2, 500, 64, device="cuda")
    with torch.inference_mode():
        z0 = enc(x, enc.latents(x))
        n0 = _lib.launch_count()
        z1 = enc(x.clone(), enc.latents(x))              # same content, new tensor
        assert _lib.launch_count() <= n0 + 4 and z1 is z0  # only the content hash (pio_hash_words) was launched
        z2 = enc(x + 1.0, enc.latents(x))                 # different content
        assert _lib.launch_count() > n0 and not torch.equal(z2, z0)
        n1 = _lib.launch_count()
        enc.cross_attend.attention.proj_q.bias.add_(0.5)  # parameter update
        enc(x + 1.0, enc.latents(x))
        assert _lib.launch_count() > n1 + 4               # more than the hash launches: the encoder ran
        # a PositionedInput is hashed as its two parts (no dense array is built for the cache)
        feats, table = torch.randn(2, 500, 3, device="cuda"), torch.randn(500, 61, device="cuda")
        za = enc(pio.PositionedInput(feats, table), enc.latents(x))
        n2 = _lib.launch_count()
        zb = enc(pio.PositionedInput(feats.clone(), table.clone()), enc.latents(x))
        assert zb is za and _lib.launch_count() <= n2 + 4  # only pio_hash_words launches


def test_empty_batch_and_empty_query_sets_return_empty_outputs():
    """The reference's modules accept an empty batch / an empty query array (every op is a no-op on them); the drop-in
    returns tensors of the same shapes without launching anything."""
    import perceiverio_pytorch_b200 as pio
    from perceiverio_pytorch_b200 import _lib
    enc = pio.PerceiverEncoder(num_input_channels=37, num_self_attends_per_block=1, num_blocks=1, num_latents=16,
                               num_latent_channels=64, num_self_attend_heads=4).eval().cuda()
    dec = pio.PerceiverDecoder(query_channels=48, final_project_out_channels=10, num_latent_channels=64).eval().cuda()
    sa = pio.SelfAttention(in_channels=64, widening_factor=1, num_heads=4).eval().cuda()
    n0 = _lib.launch_count()
    with torch.inference_mode():
        x0 = torch.zeros(0, 100, 37, device="cuda")
        z0 = enc(x0, enc.latents(x0))
        assert z0.shape == (0, 16, 64)
        assert dec(torch.zeros(0, 7, 48, device="cuda"), z0).shape == (0, 7, 10)
        z = torch.randn(2, 16, 64, device="cuda")
        assert dec(torch.zeros(2, 0, 48, device="cuda"), z).shape == (2, 0, 10)
        assert sa(torch.zeros(0, 16, 64, device="cuda")).shape == (0, 16, 64)
        m, y = sa(torch.zeros(0, 16, 64, device="cuda"), return_matrix=True)
        assert m.shape == (0, 4, 16, 16) and y.shape == (0, 16, 64)
    assert _lib.launch_count() == n0


@pytest.mark.parametrize("nk,nq", [(1, 1), (63, 129), (257, 5), (1000, 1)])
def test_ragged_sizes_match_oracle(nk, nq):
    """Key / query counts that are not multiples of any tile (down to a single key and a single query)."""
    import perceiverio_pytorch_b200 as pio
    from oracle import perceiver_oracle as O
    torch.manual_seed(nk * 1000 + nq)
    enc = pio.PerceiverEncoder(num_input_channels=37, num_self_attends_per_block=1, num_blocks=1, num_latents=24,
                               num_latent_channels=64, num_self_attend_heads=4).eval()
    dec = pio.PerceiverDecoder(query_channels=50, final_project_out_channels=10, num_latent_channels=64).eval()
    _perturb(enc, 3)
    _perturb(dec, 4)
    inputs, query = torch.randn(3, nk, 37), torch.randn(3, nq, 50)
    z_ref, out_ref = _oracle_enc_dec(enc, dec, dict(num_blocks=1, num_self_attends_per_block=1, num_cross_attend_heads=1,
                                                    num_self_attend_heads=4, use_query_residual=True),
                                     dict(num_heads=1, use_query_residual=False, final_project=True), inputs, query)
    enc, dec = enc.cuda(), dec.cuda()
    with torch.inference_mode():
        xi = inputs.cuda()
        z = enc(xi, enc.latents(xi))
        out = dec(query.cuda(), z)
    assert rel_err(z.cpu(), z_ref)[0] <= BF16_TOL
    assert rel_err(out.cpu(), out_ref)[0] <= BF16_TOL


@pytest.mark.parametrize("B,Nq,Nk,masked", [(2, 40, 5000, True), (1, 200, 17000, False), (1, 130, 16500, True)])
def test_wide_single_head_encoder_takes_the_key_side_fold_on_the_explicit_path(B, Nq, Nk, masked, monkeypatch):
    """A single-head cross-attend over a long input array whose width the streaming kernels do not cover (> 384 channels:
    the multimodal encoder's 704): K and V projections folded onto the query side, S = Q' LN(x)^T, O' = P LN(x) with the
    normalised inputs as the MN-major operand (engine.cross_attention_key_fold_wide) — against the oracle, incl. a key
    mask, a sample without any valid key (rows wiped: `final.bias` only) and the key-split P.V of the batch-1 case."""
    import perceiverio_pytorch_b200 as pio
    from perceiverio_pytorch_b200 import engine
    from oracle import perceiver_oracle as O
    torch.manual_seed(Nk)
    Cq, Ck = 96, 448
    ca = pio.CrossAttention(q_in_channels=Cq, kv_in_channels=Ck, num_heads=1, use_query_residual=True).eval()
    _perturb(ca, 21)
    q, kv = torch.randn(B, Nq, Cq), torch.randn(B, Nk, Ck)
    mask = None
    if masked:
        km = torch.rand(B, Nk) > 0.3
        if B > 1:
            km[1] = False                  # no valid key at all for sample 1
        mask = O.make_cross_attention_mask(torch.ones(B, Nq, dtype=torch.bool), km)
    sd = {k: v.detach() for k, v in ca.state_dict().items()}
    ref = O.cross_attention(sd, "", 1, True, q, kv, mask)
    calls = []
    real = engine.cross_attention_key_fold_wide
    monkeypatch.setattr(engine, "cross_attention_key_fold_wide", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    ca = ca.cuda()
    with torch.inference_mode():
        pmask = pio.make_cross_attention_mask(torch.ones(B, Nq, dtype=torch.bool, device="cuda"), km.cuda()) if masked else None
        got = ca(q.cuda(), kv.cuda(), attention_mask=pmask)
        monkeypatch.setattr(engine, "KFOLD_WIDE_MIN_KEYS", 10 ** 9)
        plain = ca(q.cuda(), kv.cuda(), attention_mask=pmask)
    assert calls == [1]
    assert rel_err(got.cpu(), ref)[0] <= BF16_TOL, rel_err(got.cpu(), ref)
    assert rel_err(plain.cpu(), ref)[0] <= BF16_TOL
    if masked and B > 1:     # wiped rows: q + final.bias (no trace of the folded value bias), then the MLP — on both paths
        assert rel_err(got[1].cpu(), ref[1])[0] <= 4e-3
        assert rel_err(got[1].cpu(), plain[1].cpu())[0] <= 1e-3


@pytest.mark.parametrize("n_out,n_post,mode", [(40, 24, "bf16"), (40, 24, "fp16"), (2, 5, "bf16"), (1000, 1000, "bf16"),
                                               (40, 24, "bf16x3")])
def test_decoder_absorbs_the_postprocessor_linear(n_out, n_post, mode):
    """`final_layer` and the postprocessor's own Linear (postprocessors.py:176-187, :200-208)
    composed into one projection inside the decoder — against the oracle's decoder followed by the Linear in fp32, on
    the wide head (tensor cores), the narrow head (folded into the MLP's second layer) and the validation mode."""
    import perceiverio_pytorch_b200 as pio
    from perceiverio_pytorch_b200 import engine
    from oracle import perceiver_oracle as O
    torch.manual_seed(n_out + n_post)
    C, Cq, Nl, Nq = 128, 96, 40, 300
    dec = pio.PerceiverDecoder(query_channels=Cq, final_project_out_channels=n_out, num_latent_channels=C,
                               use_query_residual=True).eval()
    post = torch.nn.Linear(n_out, n_post)
    _perturb(dec, 11)
    with torch.no_grad():
        post.bias.normal_(0, 0.1)
    query, latents = torch.randn(2, Nq, Cq), torch.randn(2, Nl, C)
    sd = {k: v.detach() for k, v in dec.state_dict().items()}
    ref = O.decoder_forward(sd, "", query=query, latents=latents, query_mask=None, num_heads=1, use_query_residual=True,
                            final_project=True)
    ref = torch.nn.functional.linear(ref, post.weight.detach(), post.bias.detach())
    keys_before = list(dec.state_dict().keys())
    dec, post = dec.cuda(), post.cuda()
    with torch.inference_mode(), engine.precision_scope(mode):
        got = dec(query.cuda(), latents.cuda(), post_linear=post)
        plain = post(dec(query.cuda(), latents.cuda()))
    assert list(dec.state_dict().keys()) == keys_before       # the composed weights are no part of the module tree
    assert got.shape == ref.shape == (2, Nq, n_post)
    tol = 1e-4 if mode == "bf16x3" else BF16_TOL
    assert rel_err(got.cpu(), ref)[0] <= tol, rel_err(got.cpu(), ref)
    assert rel_err(plain.cpu(), ref)[0] <= tol
    # derived weights follow in-place updates of either layer
    with torch.no_grad():
        post.weight.mul_(2.0)
        post.bias.mul_(2.0)
    with torch.inference_mode(), engine.precision_scope(mode):
        again = dec(query.cuda(), latents.cuda(), post_linear=post)
    assert rel_err(again.cpu(), 2.0 * ref)[0] <= tol
    with pytest.raises(ValueError, match="post_linear"):
        dec(query.cuda(), latents.cuda(), post_linear=torch.nn.Linear(n_out + 1, 3).cuda())


def test_standalone_mlp_matches_oracle():
    """The bare `MLP.forward` (transformer_primitives.py:212-216), as a caller of the module API would use it:
    widening factors 1 and 4, an odd channel count, a 4-D input."""
    import perceiverio_pytorch_b200 as pio
    from oracle import perceiver_oracle as O
    for cin, cout, wf, shape in ((64, None, 4, (3, 50, 64)), (322, 322, 1, (2, 7, 9, 322)), (1024, 512, 1, (300, 1024))):
        torch.manual_seed(cin)
        m = pio.MLP(in_channels=cin, out_channels=cout, widening_factor=wf).eval()
        _perturb(m, 9)
        x = torch.randn(*shape)
        ref = O.mlp({k: v.detach() for k, v in m.state_dict().items()}, "", x)
        with torch.inference_mode():
            got = m.cuda()(x.cuda())
        assert got.shape == ref.shape
        assert rel_err(got.cpu(), ref)[0] <= BF16_TOL, (cin, rel_err(got.cpu(), ref))


def test_attention_with_different_key_and_value_input_widths():
    """`Attention(q_in, k_in, v_in)` allows k_in_channels != v_in_channels (transformer_primitives.py:73-75)."""
    import perceiverio_pytorch_b200 as pio
    from oracle import perceiver_oracle as O
    torch.manual_seed(4)
    m = pio.Attention(q_in_channels=32, k_in_channels=24, v_in_channels=40, num_heads=4, qk_out_channels=32,
                      v_out_channels=48, output_channels=40).eval()
    _perturb(m, 10)
    q, k, v = torch.randn(2, 20, 32), torch.randn(2, 50, 24), torch.randn(2, 50, 40)
    ref = O.attention({kk: vv.detach() for kk, vv in m.state_dict().items()}, "", 4, q, k, v, None)
    with torch.inference_mode():
        got = m.cuda()(q.cuda(), k.cuda(), v.cuda())
    assert rel_err(got.cpu(), ref)[0] <= BF16_TOL


def test_mask_edited_in_place_after_construction_is_refactored():
    """`make_cross_attention_mask` attaches the two rank-1 factors for the kernels; editing the dense mask in place
    afterwards must not leave stale factors in use."""
    import perceiverio_pytorch_b200 as pio
    from oracle import perceiver_oracle as O
    torch.manual_seed(5)
    m = pio.CrossAttention(q_in_channels=48, kv_in_channels=40, num_heads=4, qk_channels=32, v_channels=48).eval()
    _perturb(m, 11)
    q, kv = torch.randn(2, 30, 48), torch.randn(2, 60, 40)
    km = torch.ones(2, 60, dtype=torch.bool)
    mask = pio.make_cross_attention_mask(torch.ones(2, 30, dtype=torch.bool), km)
    mask[:, :, 40:] = False          # in-place edit: keys 40.. masked for every query (still an outer product)
    ref = O.cross_attention({k: v.detach() for k, v in m.state_dict().items()}, "", 4, True, q, kv, mask.clone())
    dev_mask = _to_cuda_keep_factors(mask)     # built outside inference mode: inference tensors carry no version counter
    with torch.inference_mode():
        got = m.cuda()(q.cuda(), kv.cuda(), attention_mask=dev_mask)
    assert rel_err(got.cpu(), ref)[0] <= BF16_TOL


def _to_cuda_keep_factors(mask):
    """Move a mask to the GPU the way a caller holding it on the device would have built it there: the attribute does
    not survive .cuda(), so rebuild and repeat the edit on the device."""
    import perceiverio_pytorch_b200 as pio
    b, nq, nk = mask.shape
    m = pio.make_cross_attention_mask(torch.ones(b, nq, dtype=torch.bool, device="cuda"),
                                      torch.ones(b, nk, dtype=torch.bool, device="cuda"))
    m[:, :, 40:] = False
    return m


def test_fused_layernorm_tower_falls_back_on_rows_with_a_large_common_offset():
    """The fused form rounds x to bf16 before the mean subtraction; a residual stream with a large common offset
    (|mean| >> std) must be detected from the produced statistics and keep the LayerNorm kernels."""
    import warnings
    import perceiverio_pytorch_b200 as pio
    from oracle import perceiver_oracle as O
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(num_input_channels=261, num_self_attends_per_block=2, num_blocks=1, num_latents=512,
                               num_latent_channels=1024).eval()
    _perturb(enc, 12)
    with torch.no_grad():
        enc.latent_pos_enc.pos_embs.add_(4.0)        # every latent row: mean 4, std ~0.02 .. 1
    B, Nk = 8, 2000
    x = torch.randn(B, Nk, 261)
    ref = O.encoder_forward({k: v.detach() for k, v in enc.state_dict().items()}, "", num_blocks=1,
                            num_self_attends_per_block=2, num_cross_attend_heads=1, num_self_attend_heads=8,
                            use_query_residual=True, inputs=x)
    enc = enc.cuda()
    xc = x.cuda()
    with torch.inference_mode():
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            z = enc(xc, enc.latents(xc))
        assert any("LayerNorm kernels" in str(i.message) for i in w), [str(i.message) for i in w]
        assert enc.fused_layernorm_offset > 1.0 and enc._fuse_ln_checked[1] is False
        n0 = pio._lib.launch_count()
        z2 = enc(xc, enc.latents(xc))                # second call: straight to the unfused tower, no second warning
    assert torch.equal(z, z2)
    assert rel_err(z.cpu(), ref)[0] <= BF16_TOL, rel_err(z.cpu(), ref)


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_module_on_second_device_while_first_is_current():
    """A drop-in module moved with .to('cuda:1') works while cuda:0 is the current device (per-device kernel
    attributes, launches on the tensors' device), and mixing devices in one call fails loudly."""
    import perceiverio_pytorch_b200 as pio
    from oracle import perceiver_oracle as O
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(num_input_channels=261, num_self_attends_per_block=1, num_blocks=1, num_latents=128,
                               num_latent_channels=256, num_self_attend_heads=4).eval()
    _perturb(enc, 13)
    x = torch.randn(2, 3000, 261)
    ref = O.encoder_forward({k: v.detach() for k, v in enc.state_dict().items()}, "", num_blocks=1,
                            num_self_attends_per_block=1, num_cross_attend_heads=1, num_self_attend_heads=4,
                            use_query_residual=True, inputs=x)
    torch.cuda.set_device(0)
    enc = enc.to("cuda:1")
    with torch.inference_mode():
        z = enc(x.to("cuda:1"), enc.latents(x.to("cuda:1")))
        assert z.device == torch.device("cuda:1") and torch.cuda.current_device() == 0
        assert rel_err(z.cpu(), ref)[0] <= BF16_TOL
        with pytest.raises(RuntimeError, match="different devices"):
            enc.cross_attend(enc.latents(x.to("cuda:1")), x.to("cuda:0"))


@pytest.mark.parametrize("name", ["decoder_small", "decoder_h1_querymask"])
@pytest.mark.parametrize("mode", ["bf16", "fp16"])
def test_query_fold_decoder_kernel_on_reference_goldens(name, mode, monkeypatch):
    """The query-side fold + CTA-pair decoder kernel (engine.cross_attention_query_fold -> pio_decoder_attention_fwd),
    forced onto the small single-head decoder fixtures (normally it takes over at >= 1024 queries): query residual,
    query mask with wiped rows, final projection."""
    from perceiverio_pytorch_b200 import engine, ops
    params, inputs, meta, expected = load_golden(name)
    m = _build_ours(meta)
    m.load_state_dict(params, strict=True)
    m = m.cuda()
    calls = []
    real = ops.decoder_attention
    monkeypatch.setattr(ops, "decoder_attention", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    monkeypatch.setattr(engine, "use_query_fold", lambda pa, nq, nk: pa.qfold is not None)
    with engine.precision_scope(mode):
        got = _run_ours(m, inputs, meta)
    assert calls, "the decoder kernel was not used"
    emax, el2 = rel_err(got.float().cpu(), expected)
    assert emax <= (BF16_TOL if mode == "bf16" else 2e-3), (name, mode, emax, el2)


@pytest.mark.parametrize("name", ["decoder_small", "decoder_h1_querymask"])
@pytest.mark.parametrize("mode", ["bf16", "fp16"])
def test_query_fold_on_the_explicit_path_on_reference_goldens(name, mode, monkeypatch):
    """The same fold without the decoder kernel (engine.cross_attention_query_fold_explicit: S = LN(q) K'^T, softmax,
    out = P V' + b_f + residual — what the 1024-channel classification decoder takes), forced onto the fixtures."""
    from perceiverio_pytorch_b200 import engine
    params, inputs, meta, expected = load_golden(name)
    m = _build_ours(meta)
    m.load_state_dict(params, strict=True)
    m = m.cuda()
    calls = []
    real = engine.cross_attention_query_fold_explicit
    monkeypatch.setattr(engine, "cross_attention_query_fold_explicit", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    monkeypatch.setattr(engine, "use_query_fold", lambda pa, nq, nk: False)
    monkeypatch.setattr(engine, "use_query_fold_explicit", lambda pa, nq, nk: pa.qfold is not None)
    with engine.precision_scope(mode):
        got = _run_ours(m, inputs, meta)
    assert calls, "the explicit query-fold path was not used"
    emax, el2 = rel_err(got.float().cpu(), expected)
    assert emax <= (BF16_TOL if mode == "bf16" else 2e-3), (name, mode, emax, el2)
