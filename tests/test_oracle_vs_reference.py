"""Live comparison of the CPU oracle with the unmodified reference (imported through oracle/ref_shim.py).
Skipped when /root/reference is not mounted (e.g. on the GPU box).  Tolerance 1e-5 max-norm, fp32."""
import pytest
import torch

from golden_util import rel_err
from oracle import perceiver_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


def _hot_path_io(perceiver_io_module):
    """Hook the reference PerceiverIO's _encoder/_decoder to capture the hot path's inputs/outputs."""
    rec = {}

    def enc_hook(mod, args, kwargs, out):
        rec["enc"] = (args, kwargs, out)

    def dec_hook(mod, args, kwargs, out):
        rec["dec"] = (args, kwargs, out)

    h1 = perceiver_io_module._encoder.register_forward_hook(enc_hook, with_kwargs=True)
    h2 = perceiver_io_module._decoder.register_forward_hook(dec_hook, with_kwargs=True)
    return rec, (h1, h2)


def _check_wrapper_hot_path(pio, rec, enc_cfg, dec_cfg):
    p = {k: v for k, v in pio.state_dict().items()}
    (inputs, latents), ekw, enc_out = rec["enc"]
    got = O.encoder_forward(p, "_encoder.", inputs=inputs, latents=latents, input_mask=ekw.get("input_mask"), **enc_cfg)
    emax, el2 = rel_err(got, enc_out)
    assert emax < 1e-5, ("encoder", emax, el2)
    (query, lat), dkw, dec_out = rec["dec"]
    got = O.decoder_forward(p, "_decoder.", query=query, latents=lat, query_mask=dkw.get("query_mask"), **dec_cfg)
    emax, el2 = rel_err(got, dec_out)
    assert emax < 1e-5, ("decoder", emax, el2)


def test_language_wrapper_hot_path():
    ns = ref_shim.load_wrappers()
    torch.manual_seed(0)
    model = ref_shim.perturb_parameters(ns.language.LanguagePerceiver(num_self_attends_per_block=3).eval())
    tokens = torch.randint(6, 262, (1, 2048))
    mask = torch.zeros(1, 2048, dtype=torch.bool)
    mask[:, :1500] = True
    rec, hooks = _hot_path_io(model.perceiver)
    with torch.inference_mode():
        model(tokens, mask)
    _check_wrapper_hot_path(model.perceiver, rec,
                            dict(num_blocks=1, num_self_attends_per_block=3, num_cross_attend_heads=8,
                                 num_self_attend_heads=8, use_query_residual=True),
                            dict(num_heads=8, use_query_residual=False, final_project=False))


def test_classification_wrapper_hot_path():
    ns = ref_shim.load_wrappers()
    torch.manual_seed(0)
    model = ref_shim.perturb_parameters(
        ns.classification.ClassificationPerceiver(num_self_attends_per_block=2, num_blocks=2).eval())
    img = torch.randn(1, 3, 224, 224)
    rec, hooks = _hot_path_io(model.perceiver)
    with torch.inference_mode():
        model(img)
    _check_wrapper_hot_path(model.perceiver, rec,
                            dict(num_blocks=2, num_self_attends_per_block=2, num_cross_attend_heads=1,
                                 num_self_attend_heads=8, use_query_residual=True),
                            dict(num_heads=1, use_query_residual=True, final_project=True))


def test_flow_wrapper_hot_path_small_image():
    ns = ref_shim.load_wrappers()
    torch.manual_seed(0)
    model = ref_shim.perturb_parameters(
        ns.flow.FlowPerceiver(img_size=(64, 80), num_latents=256, num_self_attends_per_block=2).eval())
    a, b = torch.randn(1, 3, 64, 80), torch.randn(1, 3, 64, 80)
    rec, hooks = _hot_path_io(model.perceiver)
    with torch.inference_mode():
        model(a, b, test_mode=False)
    _check_wrapper_hot_path(model.perceiver, rec,
                            dict(num_blocks=1, num_self_attends_per_block=2, num_cross_attend_heads=1,
                                 num_self_attend_heads=16, use_query_residual=True),
                            dict(num_heads=1, use_query_residual=False, final_project=True))


def test_multimodal_wrapper_hot_path_small_video():
    """The fourth wrapper (multimodal_perceiver.py:146-161): three modalities padded to a common width, the label token
    replaced by the mask embedding (input_mask_probs label = 1.0, perceiver.py:481-493), subsampled output queries; the
    hooks keep the last of the chunk calls."""
    ns = ref_shim.load_wrappers()
    torch.manual_seed(0)
    model = ref_shim.perturb_parameters(
        ns.multimodal.MultiModalPerceiver(img_size=(16, 16), num_frames=2, num_classes=20, audio_samples_per_frame=64,
                                          num_self_attends_per_block=2, num_latents=48, num_latent_channels=512).eval())
    images, audio = torch.rand(1, 2, 3, 16, 16), 0.1 * torch.randn(1, 128, 1)
    rec, hooks = _hot_path_io(model.perceiver)
    with torch.inference_mode():
        out = model(images, audio, n_chunks=2)
    assert out["image"].shape == images.shape and out["label"].shape == (1, 20)
    (inputs, _), _, _ = rec["enc"]
    assert inputs.shape[1] == 2 * 4 * 4 + 128 // 16 + 1          # image patches + audio patches + the label token
    _check_wrapper_hot_path(model.perceiver, rec,
                            dict(num_blocks=1, num_self_attends_per_block=2, num_cross_attend_heads=1,
                                 num_self_attend_heads=8, use_query_residual=True),
                            dict(num_heads=1, use_query_residual=False, final_project=True))


@pytest.mark.parametrize("heads,qk,v", [(1, None, None), (4, 32, 80)])
def test_cross_attention_random(heads, qk, v):
    ns = ref_shim.load_reference()
    torch.manual_seed(1)
    m = ref_shim.perturb_parameters(ns.primitives.CrossAttention(q_in_channels=48, kv_in_channels=33, num_heads=heads,
                                                                 qk_channels=qk, v_channels=v).eval())
    q, kv = torch.randn(3, 17, 48), torch.randn(3, 129, 33)
    km = torch.rand(3, 129) > 0.3
    km[2] = False  # a sample with every key masked: the whole output row block is wiped
    mask = ns.primitives.make_cross_attention_mask(torch.ones(3, 17, dtype=torch.bool), km)
    with torch.inference_mode():
        want = m(q, kv, attention_mask=mask)
    got = O.cross_attention(dict(m.state_dict()), "", heads, True, q, kv, O.make_cross_attention_mask(torch.ones(3, 17, dtype=torch.bool), km))
    assert rel_err(got, want)[0] < 1e-5


def test_key_shard_combine_identity():
    """The (O, m, l) merge over key shards equals un-sharded attention, in fp64."""
    torch.manual_seed(2)
    q = torch.randn(2, 9, 2, 8, dtype=torch.float64) * 4
    k = torch.randn(2, 50, 2, 8, dtype=torch.float64)
    v = torch.randn(2, 50, 2, 5, dtype=torch.float64)
    km = torch.rand(2, 50) > 0.4
    km[:, 25:37] = False  # one shard holds only masked keys
    full = O.attend(q, k, v, O.make_cross_attention_mask(torch.ones(2, 9, dtype=torch.bool), km))
    parts = [O.attend_partial(q, k[:, s], v[:, s], km[:, s]) for s in (slice(0, 25), slice(25, 37), slice(37, 50))]
    merged = O.combine_partials(parts).permute(0, 2, 1, 3).reshape(2, 9, 10)
    assert float((merged - full).abs().max()) < 1e-12


def test_oracle_image_preprocessing_matches_reference():
    """The input-side glue of the pixels recipe restated in the oracle."""
    ref = ref_shim.load_wrappers()
    if ref is None:
        pytest.skip("reference tree not mounted")
    import importlib
    prep_mod = importlib.import_module("perceiver_io.io_processors.preprocessors")
    pe = importlib.import_module("perceiver_io.position_encoding")
    torch.manual_seed(0)
    for (h, w), sd in (((16, 12), 1), ((224, 224), 1), ((32, 32), 2)):
        prep = prep_mod.ImagePreprocessor(img_size=(h, w), input_channels=3, prep_type="pixels", spatial_downsample=sd,
                                          position_encoding_type=pe.PosEncodingType.FOURIER,
                                          fourier_position_encoding_kwargs=dict(concat_pos=True, max_resolution=(224, 224),
                                                                                num_bands=64, sine_only=False))
        img = torch.randn(2, 3, h, w)
        want, want_nopos = prep(img)
        got = O.image_inputs_pixels(img, 64, (224, 224), sd)
        assert got.shape == want.shape and torch.equal(got, want)
