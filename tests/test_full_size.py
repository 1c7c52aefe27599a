"""The four BASELINE.json configurations at FULL size and FULL depth (50,176 / 182,528 / 52,097 inputs, 48 / 24 / 8 / 26
layers): the CUDA path against

  * the committed subsample of the LIVE reference's output (tests/golden/full/*.npz, written by
    oracle/make_golden_full.py in the build container), and
  * the complete output of the fp32 CPU oracle, evaluated on the test machine's host cores (seconds per config).

Parameters and inputs are regenerated from seeds on both sides (oracle/full_configs.py).  Tolerance: BASELINE.json's
max|d| / max|ref| <= 1e-2.  Language, classification and multimodal meet it with bf16 operands.  Optical flow does not,
and cannot: a CPU emulation of the reference ALGORITHM with bf16-rounded operands already gives 1.86e-2 (SURVEY.md
section 0.4), the CUDA path measures 2.7e-2 on this fixture — the 322 -> 2 regression head amplifies operand rounding.
Flow therefore runs with fp16 operands (`precision="fp16"`: same kernels, same tensor-core rate, 11 instead of 8
mantissa bits — what `swap_hot_path` selects for such heads and what the reference's own mixed-precision mode uses,
flow_perceiver.py:129) and is held to 5e-3; every config is also checked in that mode.  The CPU-only tests pin the
oracle itself against the same fixtures.
"""
import os

import numpy as np
import pytest
import torch

from golden_util import rel_err
from oracle import full_configs as F

FULL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full")
# (config, arithmetic mode) -> bound on max|d| / max|ref| of the output
CASES = {("language", "bf16"): 1e-2, ("classification", "bf16"): 1e-2, ("multimodal", "bf16"): 1e-2,
         ("flow", "fp16"): 5e-3,
         ("language", "fp16"): 3e-3, ("classification", "fp16"): 3e-3, ("multimodal", "fp16"): 3e-3,
         # recorded, not a pass criterion of the model: bf16 operands on the flow head (see the module docstring)
         ("flow", "bf16"): 5e-2}


def _fixture(name):
    z = np.load(os.path.join(FULL_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def _drop_in(name, fx):
    import perceiverio_pytorch_b200 as pio
    cfg = F.FULL_CONFIGS[name]
    enc = F.seeded_fill(pio.PerceiverEncoder(**cfg["enc"]).eval(), int(fx["param_seeds"][0]))
    dec = F.seeded_fill(pio.PerceiverDecoder(**cfg["dec"]).eval(), int(fx["param_seeds"][1]))
    return enc, dec


def _sub_err(got, fx, key):
    want = torch.from_numpy(fx[key])
    sub = F.subsample(got.float().cpu(), int(fx[key + "_step"]))
    assert tuple(got.shape) == tuple(int(v) for v in fx[key + "_shape"])
    return float((sub.double() - want.double()).abs().max() / float(fx[key + "_absmax"]))


@pytest.mark.parametrize("name", ["language", "multimodal", "classification"])
def test_oracle_matches_live_reference_at_full_size(name):
    """CPU: the oracle reproduces the committed full-size reference vectors (flow, 12 s and several GB, is checked by
    oracle/make_golden_full.py when the fixtures are written and by the GPU test below)."""
    fx = _fixture(name)
    enc, dec = _drop_in(name, fx)      # the drop-in modules only hold the parameters here (no forward on CPU)
    data = F.hot_path_inputs(name, int(fx["input_seed"]))
    z, out = F.oracle_forward(name, dict(enc.state_dict()), dict(dec.state_dict()), data)
    assert _sub_err(z, fx, "latents") < 1e-5 and _sub_err(out, fx, "output") < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode", sorted(CASES))
def test_cuda_path_matches_reference_and_oracle_at_full_size(name, mode):
    from perceiverio_pytorch_b200 import engine
    tol = CASES[(name, mode)]
    fx = _fixture(name)
    enc, dec = _drop_in(name, fx)
    data = F.hot_path_inputs(name, int(fx["input_seed"]))
    z_ref, out_ref = F.oracle_forward(name, dict(enc.state_dict()), dict(dec.state_dict()), data)
    # the oracle itself against the live reference's vectors, on this machine
    assert _sub_err(z_ref, fx, "latents") < 1e-5 and _sub_err(out_ref, fx, "output") < 1e-5
    enc, dec = enc.cuda(), dec.cuda()
    enc.precision = dec.precision = mode
    assert engine.PRECISION == "bf16"            # the per-module attribute selects the mode, the global default stays
    cu = {k: (v.cuda() if isinstance(v, torch.Tensor) else None) for k, v in data.items()}
    with torch.inference_mode():
        z = enc(cu["inputs"], enc.latents(cu["inputs"]), input_mask=cu["input_mask"])
        out = dec(cu["query"], z, query_mask=cu["query_mask"])
        z2 = enc(cu["inputs"], enc.latents(cu["inputs"]), input_mask=cu["input_mask"])
        out2 = dec(cu["query"], z2, query_mask=cu["query_mask"])
    assert torch.isfinite(out).all()
    ez, eo = rel_err(z.cpu(), z_ref), rel_err(out.cpu(), out_ref)
    gz, go = _sub_err(z, fx, "latents"), _sub_err(out, fx, "output")
    print(f"\nFULL {name} [{mode}]: vs oracle latents max {ez[0]:.3e} l2 {ez[1]:.3e}, output max {eo[0]:.3e} l2 {eo[1]:.3e}; "
          f"vs reference subsample latents {gz:.3e}, output {go:.3e}")
    assert eo[0] <= tol and ez[0] <= 1e-2, (name, mode, ez, eo)
    assert go <= tol and gz <= 1e-2, (name, mode, gz, go)
    assert engine.PRECISION == "bf16"
    # twice in a row: bit-identical (no atomics anywhere on the path, fixed reduction orders)
    assert torch.equal(z, z2) and torch.equal(out, out2), name


@pytest.mark.gpu
def test_classification_full_size_from_images_matches_oracle():
    """configs[1] through the boundary bench.py times: images -> PositionedInput (pixels + Fourier table, fused into the
    encoder's LayerNorm) -> 48 layers -> decoder, B = 2; and the fused-LayerNorm tower (B * 512 >= 4096 rows) at B = 8
    against the dense-array path, twice, bit-identically."""
    import perceiverio_pytorch_b200 as pio
    from perceiverio_pytorch_b200 import engine
    name = "classification"
    fx = _fixture(name)
    enc, dec = _drop_in(name, fx)
    data = F.hot_path_inputs(name, int(fx["input_seed"]), batch=8)
    sub = {k: (v[:2] if isinstance(v, torch.Tensor) else None) for k, v in data.items()}
    z_ref, out_ref = F.oracle_forward(name, dict(enc.state_dict()), dict(dec.state_dict()), sub)
    enc, dec = enc.cuda(), dec.cuda()
    img = data["images"].cuda()
    table = pio.fourier_position_table((224, 224), 64, device="cuda")
    query = data["query"].cuda()
    with torch.inference_mode():
        pin = pio.PositionedInput(img.movedim(-3, -1).reshape(8, 224 * 224, 3), table)
        assert 8 * 512 >= engine.FUSE_LN_MIN_ROWS
        z = enc(pin, enc.latents(pin))
        out = dec(query, z)
        z_again = enc(pin, enc.latents(pin))
        dense = data["inputs"].cuda()
        z_dense = enc(dense, enc.latents(dense))
        # B = 2 with the fusion switched off -> exact two-pass LayerNorm kernels
        pin2 = pio.PositionedInput(img[:2].movedim(-3, -1).reshape(2, 224 * 224, 3), table)
        enc.fuse_layernorm = False
        out_small = dec(query[:2], enc(pin2, enc.latents(pin2)))
        enc.fuse_layernorm = None
    e8 = rel_err(out[:2].cpu(), out_ref)
    e2 = rel_err(out_small.cpu(), out_ref)
    ez = rel_err(z[:2].cpu(), z_ref)
    print(f"\nFULL classification from images: B=8 (fused LayerNorm tower) output max {e8[0]:.3e} l2 {e8[1]:.3e}, latents "
          f"{ez[0]:.3e}; B=2 (LayerNorm kernels) output max {e2[0]:.3e} l2 {e2[1]:.3e}; "
          f"PositionedInput vs dense latents {rel_err(z, z_dense)[0]:.3e}")
    assert e8[0] <= 1e-2 and e2[0] <= 1e-2 and ez[0] <= 1e-2
    assert torch.equal(z, z_again)
    assert rel_err(z, z_dense)[0] <= 8e-3


@pytest.mark.gpu
def test_flow_batch_two_takes_the_fused_layernorm_tower_within_tolerance():
    """Optical flow at B = 2 (2 x 2048 latent rows) reaches the fused-LayerNorm tower's row threshold; it must hold the
    same bound as B = 1 on the LayerNorm kernels (sample 0 against the B = 1 oracle result)."""
    name = "flow"
    fx = _fixture(name)
    enc, dec = _drop_in(name, fx)
    data = F.hot_path_inputs(name, int(fx["input_seed"]))
    z_ref, out_ref = F.oracle_forward(name, dict(enc.state_dict()), dict(dec.state_dict()), data)
    enc, dec = enc.cuda(), dec.cuda()
    enc.precision = dec.precision = "fp16"
    x = data["inputs"].cuda()
    x2 = torch.cat([x, torch.randn(1, x.shape[1], x.shape[2], device="cuda",
                                   generator=torch.Generator(device="cuda").manual_seed(5))], 0)
    with torch.inference_mode():
        z = enc(x2, enc.latents(x2))
        out = dec(x2, z)
    e = rel_err(out[:1].cpu(), out_ref)
    print(f"\nFULL flow B=2: output max {e[0]:.3e} l2 {e[1]:.3e}; latents {rel_err(z[:1].cpu(), z_ref)[0]:.3e}")
    assert e[0] <= 5e-3, e
