"""Per-kernel device timeline of one BASELINE.json config (CUPTI activity records of replayed CUDA graphs):

    python tools/timeline.py flow [--subsystem decoder|encoder|tower|forward] [--precision fp16]

Prints the kernel time per replay by kernel name, the number of launches and the idle gaps between kernels."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import perceiverio_pytorch_b200 as pio  # noqa: E402
from perceiverio_pytorch_b200 import engine  # noqa: E402
from perceiverio_pytorch_b200.graph import GraphedForward  # noqa: E402
from bench_configs import CONFIGS, perturb  # noqa: E402
from bench import graph_gap_profile  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--subsystem", default="forward")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--replays", type=int, default=5)
    a = ap.parse_args()
    engine.set_precision(a.precision)
    cfg = CONFIGS[a.config]
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(**cfg["enc"]).eval()
    dec = pio.PerceiverDecoder(**cfg["dec"]).eval()
    perturb(enc, 1)
    perturb(dec, 2)
    enc, dec = enc.cuda(), dec.cuda()
    B, Nk, Nq = cfg["B"], cfg["Nk"], cfg["Nq"]
    x = torch.randn(B, Nk, cfg["enc"]["num_input_channels"], device="cuda")
    query = x if a.config == "flow" else torch.randn(B, Nq, cfg["dec"]["query_channels"], device="cuda")
    with torch.inference_mode():
        lat = enc.latents(x)
        z0 = enc.cross_attend._forward_factored(lat, x)[0]
        z1 = enc(x, lat)
    fns = {
        "encoder": (lambda xx: enc.cross_attend._forward_factored(enc.latents(xx), xx)[0], x),
        "tower": (lambda zz: _tower(enc, zz), z0),
        "decoder": (lambda zz: dec(query, zz), z1),
        "forward": (lambda xx: dec(xx if a.config == "flow" else query, enc(xx, enc.latents(xx))), x),
    }
    fn, inp = fns[a.subsystem]
    g = GraphedForward(fn, [inp], warmup=1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    r = graph_gap_profile(g, g.inputs[0], replays=a.replays, between=flush.zero_)
    print(json.dumps(dict(config=a.config, subsystem=a.subsystem, precision=a.precision, **r), indent=1))


def _tower(enc, z):
    for _ in range(enc._num_blocks):
        for sa in enc.self_attends:
            z = sa(z)
    return z


if __name__ == "__main__":
    main()
