"""Does operand multicast (fewer L2 reads per FLOP) speed the big GEMMs up?  QKV shape of the bench tower on the
single-CTA kernel with 1 / 2 / 4-wide clusters (B multicast along M) against the CTA-pair kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from perceiverio_pytorch_b200 import ops  # noqa: E402
from chain_bench import chain_time  # noqa: E402

M, C = 32768, 1024
dev = "cuda"
with torch.inference_mode():
    a = [torch.randn(M, C, device=dev).to(torch.bfloat16) for _ in range(2)]
    w3 = [(0.02 * torch.randn(3 * C, C, device=dev)).to(torch.bfloat16) for _ in range(2)]
    b3 = torch.zeros(3 * C, device=dev)
    q16 = [torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    out = {}
    for name, kw in [("pair", dict()), ("single_cl1", dict(kernel=1, tile_n=256, cluster_m=1)),
                     ("single_cl2", dict(kernel=1, tile_n=256, cluster_m=2)),
                     ("single_cl4", dict(kernel=1, tile_n=256, cluster_m=4))]:
        t = chain_time(lambda i: ops.gemm(a[i & 1], w3[i & 1], M=M, N=3 * C, K=C, bias=b3, out_bf16=q16[i & 1],
                                          ldo16=3 * C, **kw), 20)
        out[name] = {"us": round(t, 1), "tflops": round(2 * M * C * 3 * C / t / 1e6)}
print(json.dumps(out))
