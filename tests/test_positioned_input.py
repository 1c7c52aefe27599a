"""Input-side glue kept on the device: `PositionedInput`, the Fourier table restatement and
the fused concat + LayerNorm kernel."""
import pytest
import torch

import perceiverio_pytorch_b200 as pio


def test_positioned_input_stands_for_the_concatenation():
    torch.manual_seed(0)
    img = torch.randn(2, 3, 4, 6)
    feats = img.movedim(-3, -1).reshape(2, 24, 3)          # a view of the NCHW image
    table = torch.randn(24, 10)
    p = pio.PositionedInput(feats, table)
    assert p.shape == (2, 24, 13) and p.dtype == torch.float32 and p.dim() == 3 and not p.is_cuda
    dense = p.dense()
    assert torch.equal(dense[..., :3], feats) and torch.equal(dense[1, :, 3:], table)
    s = p[:, 4:10]                                        # what `restructure` (perceiver.py:370-387) does
    assert isinstance(s, pio.PositionedInput) and torch.equal(s.dense(), dense[:, 4:10])
    with pytest.raises(TypeError):
        p[:, ::2]
    with pytest.raises(ValueError):
        pio.PositionedInput(feats, torch.randn(23, 10))


def test_fourier_table_matches_live_reference():
    from oracle import ref_shim
    ref = ref_shim.load_reference()
    if ref is None:
        pytest.skip("reference tree not mounted")
    import importlib
    pe = importlib.import_module("perceiver_io.position_encoding")
    for index_dims, bands, kw in [((6, 5), 4, {}), ((224, 224), 64, {}), ((3, 8, 7), 5, dict(sine_only=True)),
                                  ((12,), 6, dict(concat_pos=False, max_resolution=(30,)))]:
        want = pe.FourierPositionEncoding(index_dims=list(index_dims), num_bands=bands, **kw)(batch_size=1)[0]
        got = pio.fourier_position_table(index_dims, bands, **kw)
        assert got.shape == want.shape
        assert torch.equal(got, want), (index_dims, float((got - want).abs().max()))


def test_fourier_table_golden():
    """Known-answer values of the ImageNet recipe's table (224 x 224 positions, 64 bands -> 258 channels), taken from
    the function the test above pins bit-for-bit to the live reference."""
    t = pio.fourier_position_table((224, 224), 64)
    assert t.shape == (50176, 258)
    assert float(t[0, 0]) == -1.0 and float(t[0, 1]) == -1.0 and float(t[-1, 0]) == 1.0
    assert abs(float(t.double().sum()) - 14834.376819364727) < 1e-3
    assert abs(float(t.double().abs().sum()) - 8239033.747733984) < 1e-1
    assert abs(float(t[12345, 200]) + 0.9973764419555664) < 1e-6
    assert abs(float(t[40000, 77]) + 0.050288937985897064) < 1e-6


def test_whole_wrapper_glue_on_live_reference():
    """`perceiver_io_forward` reproduces PerceiverIO.forward's glue (perceiver.py:287-325) around a PositionedInput.
    On CPU the reference's own encoder stands in for ours (fed with the densified input), so the test pins the glue:
    the preprocessor features, the cached position table, the query construction and the postprocessing."""
    from oracle import ref_shim
    ns = ref_shim.load_wrappers()
    if ns is None:
        pytest.skip("reference tree not mounted")
    from perceiverio_pytorch_b200 import inputs as pin_mod
    cp = ns.classification
    torch.manual_seed(0)
    for prep in (cp.PrepType.FOURIER_POS_PIXEL, cp.PrepType.LEARNED_POS_1X1CONV):
        model = cp.ClassificationPerceiver(num_classes=10, img_size=(16, 16), prep_type=prep, num_self_attends_per_block=1,
                                           num_blocks=1, num_latents=8, num_latent_channels=32).eval()
        ref_shim.perturb_parameters(model, 5)
        img = torch.randn(2, 3, 16, 16)
        with torch.inference_mode():
            want = model(img)
        enc = model.perceiver._encoder
        seen = []

        class _Densify(torch.nn.Module):
            def __init__(self, inner):
                super().__init__()
                self.inner = inner

            def latents(self, inputs):
                return self.inner.latents(inputs)

            def forward(self, inputs, latents, *, input_mask=None):
                seen.append(type(inputs).__name__)
                if isinstance(inputs, pio.PositionedInput):
                    inputs = inputs.dense()
                return self.inner(inputs, latents, input_mask=input_mask)

        model.perceiver._encoder = _Densify(enc)
        with torch.inference_mode():
            got = pin_mod.perceiver_io_forward(model.perceiver, img)
            got2 = pin_mod.perceiver_io_forward(model.perceiver, img)     # second call: cached table
        assert seen == ["PositionedInput", "PositionedInput"], seen
        assert got.shape == want.shape
        assert torch.allclose(got, want, atol=1e-6, rtol=1e-6) and torch.equal(got, got2)
        with torch.inference_mode():   # N3: decode only the query row the postprocessor keeps
            pruned = pin_mod.perceiver_io_forward(model.perceiver, img, only_needed_queries=True)
        # same modules, one query row instead of 1000: only the BLAS re-association differs
        assert pruned.shape == want.shape
        assert float((pruned - want).abs().max()) <= 2e-5 * float(want.abs().max()), float((pruned - want).abs().max())


def _densifying_encoder(enc, seen):
    class _Densify(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def latents(self, inputs):
            return self.inner.latents(inputs)

        def forward(self, inputs, latents, *, input_mask=None):
            seen.append(type(inputs).__name__)
            if isinstance(inputs, pio.PositionedInput):
                inputs = inputs.dense()
            return self.inner(inputs, latents, input_mask=input_mask)
    return _Densify(enc)


def test_fused_input_glue_on_the_other_wrappers():
    """swap_hot_path(..., fuse_input=True) routes every PerceiverIO.forward through `perceiver_io_forward`; `model(...)`
    must keep working for all four wrappers.  Optical flow: its FlowQuery hands the preprocessed inputs back as the
    decoder query (output_queries.py:73-76, :129-139), so the glue gives it the dense array while the encoder still gets
    the two parts.  Language / multimodal (embedding preprocessor, several modalities) fall back to the reference's own
    forward.  On CPU the reference's encoder / decoder stand in for ours: the test pins the glue."""
    from oracle import ref_shim
    ns = ref_shim.load_wrappers()
    if ns is None:
        pytest.skip("reference tree not mounted")
    import functools
    from perceiverio_pytorch_b200 import inputs as pin_mod
    torch.manual_seed(0)
    # ---- optical flow
    flow = ref_shim.perturb_parameters(
        ns.flow.FlowPerceiver(img_size=(16, 24), num_latents=32, num_self_attends_per_block=1).eval(), 3)
    a, b = torch.randn(1, 3, 16, 24), torch.randn(1, 3, 16, 24)
    with torch.inference_mode():
        want = flow(a, b, test_mode=False)
    seen = []
    flow.perceiver._encoder = _densifying_encoder(flow.perceiver._encoder, seen)
    flow.perceiver.forward = functools.partial(pin_mod.perceiver_io_forward, flow.perceiver)
    with torch.inference_mode():
        got = flow(a, b, test_mode=False)
    assert seen == ["PositionedInput"], seen
    assert got.shape == want.shape and torch.allclose(got, want, atol=1e-6, rtol=1e-5)
    # the reference's `pos=` keyword is accepted and forwarded (own forward)
    with torch.inference_mode():
        assert flow.perceiver(torch.randn(1, 2, 27, 16, 24), pos=None).shape[0] == 1
    # ---- language: an embedding preprocessor is not an image modality -> the module's own forward
    lang = ref_shim.perturb_parameters(ns.language.LanguagePerceiver(num_self_attends_per_block=1, num_latents=16,
                                                                     num_latent_channels=64, max_seq_len=64).eval(), 4)
    tok = torch.randint(6, 262, (1, 64))
    msk = torch.ones(1, 64, dtype=torch.bool)
    msk[:, 40:] = False
    with torch.inference_mode():
        want = lang(tok, msk)
    lang.perceiver.forward = functools.partial(pin_mod.perceiver_io_forward, lang.perceiver)
    with torch.inference_mode():
        got = lang(tok, msk)
    assert torch.equal(got, want)
    # ---- multimodal: several modalities with channel padding and modality masking -> the module's own forward
    mm = ref_shim.perturb_parameters(
        ns.multimodal.MultiModalPerceiver(img_size=(16, 16), num_frames=2, num_classes=20, audio_samples_per_frame=64,
                                          num_self_attends_per_block=1, num_latents=16, num_latent_channels=512).eval(), 5)
    images, audio = torch.rand(1, 2, 3, 16, 16), 0.1 * torch.randn(1, 128, 1)
    with torch.inference_mode():
        want = mm(images, audio, n_chunks=2)
    mm.perceiver.forward = functools.partial(pin_mod.perceiver_io_forward, mm.perceiver)
    with torch.inference_mode():
        got = mm(images, audio, n_chunks=2)
    assert all(torch.equal(got[k], want[k]) for k in want)


def test_positioned_image_input_matches_reference_preprocessors():
    """Every single-modality image preprocessor configuration with a concatenated position encoding: the two parts of
    the PositionedInput, densified, are exactly what the reference's ImagePreprocessor returns."""
    from oracle import ref_shim
    if ref_shim.load_wrappers() is None:
        pytest.skip("reference tree not mounted")
    import importlib
    prep_mod = importlib.import_module("perceiver_io.io_processors.preprocessors")
    pe = importlib.import_module("perceiver_io.position_encoding")
    four = dict(position_encoding_type=pe.PosEncodingType.FOURIER,
                fourier_position_encoding_kwargs=dict(concat_pos=True, max_resolution=(32, 32), num_bands=8,
                                                      sine_only=False))
    train = dict(position_encoding_type=pe.PosEncodingType.TRAINABLE,
                 trainable_position_encoding_kwargs=dict(init_scale=0.02, num_channels=24))
    torch.manual_seed(0)
    cases = [
        (dict(img_size=(32, 32), prep_type="pixels", spatial_downsample=1, **four), (2, 3, 32, 32)),
        (dict(img_size=(32, 32), prep_type="pixels", spatial_downsample=2, **four), (2, 3, 32, 32)),
        (dict(img_size=(32, 32), prep_type="conv1x1", spatial_downsample=1, num_channels=16, **train), (2, 3, 32, 32)),
        (dict(img_size=(32, 32), prep_type="conv", spatial_downsample=4, num_channels=16, **four), (2, 3, 32, 32)),
        # the optical-flow recipe: 3x3 patches of two frames stacked in the channel axis, projected (flow_perceiver.py:47-66)
        (dict(img_size=(16, 24), input_channels=27, prep_type="patches", spatial_downsample=1,
              temporal_downsample=2, conv_after_patching=True, num_channels=64, **four), (2, 2, 27, 16, 24)),
    ]
    for kw, shape in cases:
        prep = prep_mod.ImagePreprocessor(**kw).eval()
        x = torch.randn(*shape)
        with torch.inference_mode():
            want, want_nopos = prep(x)
            pin = pio.positioned_image_input(prep, x)
        assert pin is not None, kw["prep_type"]
        assert pin.shape == want.shape
        assert torch.equal(pin.dense(), want), kw["prep_type"]
        assert torch.equal(pin.features, want_nopos)
    add = prep_mod.ImagePreprocessor(img_size=(32, 32), prep_type="pixels", spatial_downsample=1, concat_or_add_pos="add",
                                     position_encoding_type=pe.PosEncodingType.TRAINABLE,
                                     trainable_position_encoding_kwargs=dict(init_scale=0.02, num_channels=3))
    assert pio.positioned_image_input(add, torch.randn(1, 3, 32, 32)) is None     # not a concatenation: left to the reference


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,Cf,Cp", [(3, 8, 12, 3, 258), (2, 16, 16, 3, 10), (1, 4, 8, 64, 258), (5, 6, 6, 7, 29)])
def test_fused_concat_layernorm_matches_dense_path(B, H, W, Cf, Cp):
    from perceiverio_pytorch_b200 import ops
    torch.manual_seed(B * 100 + Cf)
    img = torch.randn(B, Cf, H, W, device="cuda") * 2 + 0.3
    feats = img.movedim(-3, -1).reshape(B, H * W, Cf)
    table = torch.randn(H * W, Cp, device="cuda")
    C = Cf + Cp
    g, b = 1 + 0.1 * torch.randn(C, device="cuda"), 0.1 * torch.randn(C, device="cuda")
    assert ops.layernorm_concat_supported(B, H * W, Cf, Cp)
    y = ops.layernorm_concat_bf16(feats, table, g, b)
    dense = torch.cat([feats, table[None].expand(B, -1, -1)], -1)
    ref = torch.nn.functional.layer_norm(dense, (C,), g, b, 1e-5).reshape(B * H * W, C)
    err = float((y[:, :C].float() - ref).abs().max() / ref.abs().max())
    assert err < 8e-3, err
    if y.shape[1] > C:
        assert float(y[:, C:].float().abs().max()) == 0.0
    y2 = ops.layernorm_bf16(dense.reshape(B * H * W, C), g, b)
    # same values as the dense kernel up to the last bf16 bit of a few elements (the statistics are combined differently)
    d = (y.float() - y2.float()).abs()
    assert float(d.max()) <= 0.04 and float((d > 0).float().mean()) < 0.02


@pytest.mark.gpu
def test_encoder_accepts_positioned_input():
    """The ImageNet-pixels pattern: encoder on a PositionedInput == encoder on the dense concatenation."""
    torch.manual_seed(1)
    enc = pio.PerceiverEncoder(num_input_channels=261, num_self_attends_per_block=2, num_blocks=1, num_latents=128,
                               num_latent_channels=256, num_self_attend_heads=4).eval().cuda()
    img = torch.randn(4, 3, 32, 32, device="cuda")
    table = pio.fourier_position_table((32, 32), 64, device="cuda")
    pin = pio.PositionedInput(img.movedim(-3, -1).reshape(4, 1024, 3), table)
    with torch.inference_mode():
        z_fused = enc(pin, enc.latents(pin))
        dense = pin.dense()
        z_dense = enc(dense, enc.latents(dense))
    from oracle import perceiver_oracle as O
    pe = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
    z_ref = O.encoder_forward(pe, "", num_blocks=1, num_self_attends_per_block=2, num_cross_attend_heads=1,
                              num_self_attend_heads=4, use_query_residual=True, inputs=dense.cpu())
    scale = float(z_ref.abs().max())
    assert float((z_fused.cpu() - z_ref).abs().max()) / scale <= 1e-2
    assert float((z_dense.cpu() - z_ref).abs().max()) / scale <= 1e-2
    # two bf16 evaluations whose LayerNorm outputs differ in the last bit of a few elements
    assert float((z_fused - z_dense).abs().max()) / scale < 5e-3
