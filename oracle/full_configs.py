"""The four BASELINE.json model configurations at FULL size and depth, as seeded workloads.

TEST INFRASTRUCTURE ONLY (see oracle/perceiver_oracle.py): used by tests/, by oracle/make_golden_full.py (which runs
the LIVE reference on them in the build container) and by the checker legs of bench.py / tools.

Everything is derived from seeds, so a full-size case costs no fixture bytes for parameters or inputs: `seeded_fill`
writes every parameter of a module (reference or drop-in: both have the same parameter names, shapes and registration
order) from one generator, and `hot_path_inputs` draws the synthetic inputs of SURVEY.md section 8(d).  Only a
subsample of the reference's output is committed (tests/golden/full/*.npz).

Shapes follow the reference recipes: language_perceiver.py:24-70, classification_perceiver.py:76-125,
flow_perceiver.py:47-97, multimodal_perceiver.py:52-135 (hot-path shapes: SURVEY.md section 8, probed with forward hooks).
"""
from __future__ import annotations

import math

import torch

FULL_CONFIGS = {
    # BASELINE.json configs[0]
    "language": dict(
        enc=dict(num_input_channels=768, num_self_attends_per_block=26, num_blocks=1, num_latents=256,
                 num_latent_channels=1280, qk_channels=256, v_channels=1280, num_cross_attend_heads=8,
                 num_self_attend_heads=8),
        dec=dict(query_channels=768, final_project_out_channels=768, num_latent_channels=1280, qk_channels=256,
                 v_channels=768, num_heads=8, use_query_residual=False, final_project=False),
        B=1, Nk=2048, Nq=2048, masked=True),
    # BASELINE.json configs[1] (bench.py times it at B = 64; parity is checked on B = 2)
    "classification": dict(
        enc=dict(num_input_channels=261, num_self_attends_per_block=6, num_blocks=8, num_latents=512,
                 num_latent_channels=1024, num_cross_attend_heads=1, num_self_attend_heads=8),
        dec=dict(query_channels=1024, final_project_out_channels=1000, num_latent_channels=1024,
                 use_query_residual=True, num_heads=1, final_project=True),
        B=2, Nk=50176, Nq=1000, masked=False),
    # BASELINE.json configs[2]: the 182,528 preprocessed inputs are the decoder queries (output_queries.py:129-139)
    "flow": dict(
        enc=dict(num_input_channels=322, num_self_attends_per_block=24, num_blocks=1, num_latents=2048,
                 num_latent_channels=512, num_cross_attend_heads=1, num_self_attend_heads=16),
        dec=dict(query_channels=322, final_project_out_channels=2, num_latent_channels=512,
                 use_query_residual=False, num_heads=1, final_project=True),
        B=1, Nk=182528, Nq=182528, masked=False),
    # BASELINE.json configs[3]: one of the 128 chunk calls (6272 pixel + 15 audio + 1 label queries)
    "multimodal": dict(
        enc=dict(num_input_channels=704, num_self_attends_per_block=8, num_blocks=1, num_latents=784,
                 num_latent_channels=512, num_cross_attend_heads=1, num_self_attend_heads=8),
        dec=dict(query_channels=1026, final_project_out_channels=512, num_latent_channels=512,
                 use_query_residual=False, num_heads=1, final_project=True),
        B=1, Nk=52097, Nq=6288, masked=False),
}


def oracle_kwargs(name: str):
    """(encoder kwargs, decoder kwargs) in the oracle's terms."""
    c = FULL_CONFIGS[name]
    e, d = c["enc"], c["dec"]
    enc = dict(num_blocks=e["num_blocks"], num_self_attends_per_block=e["num_self_attends_per_block"],
               num_cross_attend_heads=e["num_cross_attend_heads"], num_self_attend_heads=e["num_self_attend_heads"],
               use_query_residual=True)
    dec = dict(num_heads=d["num_heads"], use_query_residual=d["use_query_residual"], final_project=d["final_project"])
    return enc, dec


def seeded_fill(module: torch.nn.Module, seed: int) -> torch.nn.Module:
    """Write every parameter from one seeded CPU generator, in registration order: matrices ~ N(0, 1/fan_in) (the
    variance of the reference's variance-scaling init; this also replaces FlowPerceiver's all-zero final_layer.weight,
    SURVEY.md section 0.3), LayerNorm weights 1 + 0.1 N and biases 0.1 N, other biases 0.02 N, the latent array 0.02 N
    (its init scale at perceiver.py:64-67)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in module.named_parameters():
            r = torch.randn(prm.shape, generator=g)
            if name.endswith("pos_embs"):
                prm.copy_(0.02 * r)
            elif "layer_norm" in name:
                prm.copy_(1.0 + 0.1 * r if name.endswith("weight") else 0.1 * r)
            elif name.endswith("bias"):
                prm.copy_(0.02 * r)
            else:
                prm.copy_(r / math.sqrt(prm.shape[-1]))
    return module


def hot_path_inputs(name: str, seed: int = 11, batch: int = None):
    """Seeded synthetic inputs of the hot path: dict(inputs [B, Nk, C], query [B, Nq, Cq], input_mask, query_mask,
    images (classification only: the [B, 3, 224, 224] batch the input array was built from))."""
    from . import perceiver_oracle as O
    c = FULL_CONFIGS[name]
    B = batch or c["B"]
    g = torch.Generator().manual_seed(seed)
    out = dict(input_mask=None, query_mask=None, images=None)
    if name == "classification":
        # the ImageNet-pixels recipe: 3 pixel channels + 258 Fourier channels per position (preprocessors.py:180-199)
        out["images"] = torch.randn(B, 3, 224, 224, generator=g)
        out["inputs"] = O.image_inputs_pixels(out["images"], 64, (224, 224), 1)
        out["query"] = (0.02 * torch.randn(1, c["Nq"], 1024, generator=g)).expand(B, -1, -1).contiguous()
    elif name == "flow":
        out["inputs"] = torch.randn(B, c["Nk"], 322, generator=g)
        out["query"] = out["inputs"]
    else:
        out["inputs"] = torch.randn(B, c["Nk"], c["enc"]["num_input_channels"], generator=g)
        out["query"] = torch.randn(B, c["Nq"], c["dec"]["query_channels"], generator=g)
    if c["masked"]:
        m = torch.zeros(B, c["Nk"], dtype=torch.bool)
        m[:, :1500] = True
        out["input_mask"] = m
        out["query_mask"] = m[:, :c["Nq"]].clone()
    return out


def oracle_forward(name: str, enc_state, dec_state, data):
    """fp32 CPU oracle forward of a full configuration -> (latents, output)."""
    from . import perceiver_oracle as O
    ek, dk = oracle_kwargs(name)
    with torch.inference_mode():
        z = O.encoder_forward(enc_state, "", inputs=data["inputs"], input_mask=data["input_mask"], **ek)
        out = O.decoder_forward(dec_state, "", query=data["query"], latents=z, query_mask=data["query_mask"], **dk)
    return z, out


def subsample(t: torch.Tensor, step: int) -> torch.Tensor:
    """Every `step`-th element of the flattened tensor (what the full-size fixtures store)."""
    return t.reshape(-1)[::step].contiguous()


# flat strides of the committed subsamples (latents, output), co-prime with the channel counts: fixtures stay < 0.7 MB
SUBSAMPLE = {"language": (5, 17), "classification": (17, 33), "flow": (17, 3), "multimodal": (9, 31)}
