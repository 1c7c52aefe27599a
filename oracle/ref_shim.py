"""Import the unmodified reference (JOBR0/PerceiverIO_Pytorch) in a container that lacks `timm` and
`matplotlib`.

TEST INFRASTRUCTURE ONLY (see oracle/perceiver_oracle.py).  The reference imports four init helpers from
`timm.models.layers` at module top (transformer_primitives.py:7, perceiver.py:6, position_encoding.py:10 ...)
and `matplotlib` through utils/utils.py:8-9.  Neither package is installed and there is no network, so this
module registers minimal stand-ins in `sys.modules` *before* the reference is imported.  The stand-ins only
affect random initialisation (irrelevant for parity: both sides always load the same state_dict) and
plotting (never called).

`load_reference()` returns a namespace with the reference modules, or ``None`` when the reference tree is
not mounted (e.g. on the GPU box) so callers can skip.
"""
from __future__ import annotations

import importlib
import math
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PIO_REFERENCE_ROOT", "/root/reference")


def _install_stand_ins():
    import torch

    if "timm.models.layers" not in sys.modules:
        def variance_scaling_(tensor, scale=1.0, mode="fan_in", distribution="truncated_normal"):
            fan_in = tensor.shape[1] if tensor.dim() > 1 else tensor.shape[0]
            fan_out = tensor.shape[0]
            denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
            std = math.sqrt(scale / denom) / .87962566103423978
            with torch.no_grad():
                return torch.nn.init.trunc_normal_(tensor, std=std, a=-2 * std, b=2 * std)

        def lecun_normal_(tensor):
            return variance_scaling_(tensor, 1.0, "fan_in", "truncated_normal")

        def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
            with torch.no_grad():
                return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.variance_scaling_ = variance_scaling_
        layers.lecun_normal_ = lecun_normal_
        layers.trunc_normal_ = trunc_normal_
        layers.to_2tuple = to_2tuple
        timm.models = models
        models.layers = layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})

    try:
        importlib.import_module("matplotlib")
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        anim = types.ModuleType("matplotlib.animation")
        anim.ArtistAnimation = object
        mpl.pyplot = plt
        mpl.animation = anim
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.animation": anim})


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "perceiver_io"))


def load_reference():
    """Return a namespace holding the reference's modules (or None if the tree is absent)."""
    if not reference_available():
        return None
    _install_stand_ins()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.primitives = importlib.import_module("perceiver_io.transformer_primitives")
    ns.perceiver = importlib.import_module("perceiver_io.perceiver")
    return ns


def load_wrappers():
    """The four task wrappers (heavier imports: cv2/einops)."""
    ns = load_reference()
    if ns is None:
        return None
    ns.classification = importlib.import_module("perceiver_io.classification_perceiver")
    ns.language = importlib.import_module("perceiver_io.language_perceiver")
    ns.flow = importlib.import_module("perceiver_io.flow_perceiver")
    ns.multimodal = importlib.import_module("perceiver_io.multimodal_perceiver")
    return ns


def perturb_parameters(module, seed: int = 1234):
    """Randomise biases and LayerNorm affines and re-initialise all-zero weights: at
    random init every bias is 0, every LN is (1, 0) and FlowPerceiver's final_layer.weight is all zeros,
    which would leave those code paths untested."""
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if name.endswith("bias"):
                if "layer_norm" in name or name.split(".")[-2].startswith("layer_norm"):
                    prm.copy_(0.1 * torch.randn(prm.shape, generator=g))
                else:
                    prm.copy_(0.02 * torch.randn(prm.shape, generator=g))
            elif "layer_norm" in name and name.endswith("weight"):
                prm.copy_(1.0 + 0.1 * torch.randn(prm.shape, generator=g))
            elif name.endswith("weight") and prm.dim() == 2 and float(prm.abs().max()) == 0.0:
                prm.copy_(torch.randn(prm.shape, generator=g) / math.sqrt(prm.shape[1]))
    return module
