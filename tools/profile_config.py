"""ncu target: two eager forwards of one BASELINE.json config at full size (launch list / per-kernel durations).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_config.py flow"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import perceiverio_pytorch_b200 as pio  # noqa: E402
from bench_configs import CONFIGS, perturb  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "flow"
cfg = CONFIGS[name]
torch.manual_seed(0)
enc = pio.PerceiverEncoder(**cfg["enc"]).eval()
dec = pio.PerceiverDecoder(**cfg["dec"]).eval()
perturb(enc, 1)
perturb(dec, 2)
enc, dec = enc.cuda(), dec.cuda()
B, Nk, Nq = cfg["B"], cfg["Nk"], cfg["Nq"]
x = torch.randn(B, Nk, cfg["enc"]["num_input_channels"], device="cuda")
query = torch.randn(B, Nq, cfg["dec"]["query_channels"], device="cuda")
imask = qmask = None
if cfg["masks"]:
    imask = torch.zeros(B, Nk, dtype=torch.bool, device="cuda")
    imask[:, :1500] = True
    qmask = imask[:, :Nq].clone()
with torch.inference_mode():
    for _ in range(2):
        out = dec(query, enc(x, enc.latents(x), input_mask=imask), query_mask=qmask)
torch.cuda.synchronize()
print("ok", name, float(out.abs().mean()))
