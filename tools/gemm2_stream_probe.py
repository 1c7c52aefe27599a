"""Developer aid: the (hi, lo) stream producer GEMM of the bench tower in a dependent chain, plus an elementwise kernel
with the same epilogue traffic (what HBM gives a mixed read / write stream of that size; pass any argument)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from perceiverio_pytorch_b200 import ops
from chain_bench import chain_time
M, C = 32768, 1024
dev = "cuda"
with torch.inference_mode():
    a = [torch.randn(M, C, device=dev).to(torch.bfloat16) for _ in range(2)]
    w = [(0.02 * torch.randn(C, C, device=dev)).to(torch.bfloat16) for _ in range(4)]
    b = torch.zeros(C, device=dev)
    hi = [torch.randn(M, C, device=dev).to(torch.bfloat16) for _ in range(2)]
    lo = [torch.randn(M, C, device=dev).to(torch.bfloat16) for _ in range(2)]
    st = ops.empty_row_stats(M, C, dev)
    out = {}
    out["stream"] = chain_time(lambda i: ops.gemm(a[i & 1], w[i % 4], M=M, N=C, K=C, bias=b, residual_hi16=hi[i & 1], residual_lo16=lo[i & 1], ldr16=C, out_bf16=hi[1 - (i & 1)], ldo16=C, out_lo16=lo[1 - (i & 1)], row_stats_out=st), 20)
    if len(sys.argv) > 1:
        out["elementwise_3r2w"] = chain_time(lambda i: (torch.add(a[i & 1], hi[i & 1], out=hi[1 - (i & 1)]), torch.add(a[i & 1], lo[i & 1], out=lo[1 - (i & 1)])), 20)
print(json.dumps({k: round(v, 1) for k, v in out.items()}))
