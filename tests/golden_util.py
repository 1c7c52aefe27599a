"""Helpers shared by the CPU (oracle) and GPU (CUDA path) golden tests."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    params = {k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    inputs = {k[len("input::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("input::")}
    meta = {k[len("meta::"):]: z[k].item() for k in z.files if k.startswith("meta::")}
    return params, inputs, meta, torch.from_numpy(z["output"])


def load_golden_matrix(name):
    """The attention matrix of a return_matrix fixture ([B,H,Nq,Nk] probabilities), or None."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return torch.from_numpy(z["output::matrix"]) if "output::matrix" in z.files else None


def run_oracle(params, inputs, meta, dtype=torch.float32):
    """Evaluate a golden case with the CPU oracle.  Fixtures with meta return_matrix yield (matrix, output)."""
    from oracle import perceiver_oracle as O
    p = {k: v.to(dtype) if v.is_floating_point() else v for k, v in params.items()}
    x = {k: v.to(dtype) if v.is_floating_point() else v for k, v in inputs.items()}
    kind = meta["kind"]
    general = dict(attention_bias=x.get("bias"), return_matrix=bool(meta.get("return_matrix", 0)))
    if kind == "attention":
        return O.attention(p, "", meta["num_heads"], x["q"], x["kv"], x["kv"], x.get("dense_mask"), **general)
    if kind == "cross":
        mask = x.get("dense_mask")
        b, nq, nk = x["q"].shape[0], x["q"].shape[1], x["kv"].shape[1]
        if "key_mask" in x:
            mask = O.make_cross_attention_mask(torch.ones(b, nq, dtype=torch.bool), x["key_mask"])
        if "query_mask" in x:
            mask = O.make_cross_attention_mask(x["query_mask"], torch.ones(b, nk, dtype=torch.bool))
        return O.cross_attention(p, "", meta["num_heads"], bool(meta["use_query_residual"]), x["q"], x["kv"], mask,
                                 **general)
    if kind == "self":
        return O.self_attention(p, "", meta["num_heads"], x["x"], x.get("dense_mask"), **general)
    if kind == "encoder":
        return O.encoder_forward(p, "", num_blocks=meta["num_blocks"],
                                 num_self_attends_per_block=meta["num_self_attends_per_block"],
                                 num_cross_attend_heads=meta["num_cross_attend_heads"],
                                 num_self_attend_heads=meta["num_self_attend_heads"],
                                 use_query_residual=bool(meta["use_query_residual"]),
                                 inputs=x["inputs"], input_mask=x.get("input_mask"))
    if kind == "decoder":
        return O.decoder_forward(p, "", num_heads=meta["num_heads"],
                                 use_query_residual=bool(meta["use_query_residual"]),
                                 final_project=bool(meta["final_project"]),
                                 query=x["query"], latents=x["latents"], query_mask=x.get("query_mask"))
    raise ValueError(kind)


def rel_err(a, b):
    """max|a-b| / max|b| and relative L2 (element-wise relative error is meaningless)."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)), \
        float((a - b).norm() / b.norm().clamp_min(1e-30))
