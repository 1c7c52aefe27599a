import torch, sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
from perceiverio_pytorch_b200 import ops
from chain_bench import chain_time
out = {}
with torch.inference_mode():
    for rows, C in [(4, 64), (256, 64), (4, 1280), (64, 1280), (256, 1280), (256, 1024), (256, 512), (512, 1280), (1024, 1280), (2048, 1280)]:
        x = torch.randn(rows, C, device='cuda'); g = torch.ones(C, device='cuda'); b = torch.zeros(C, device='cuda')
        y = torch.empty(rows, ops.pad8(C), device='cuda', dtype=torch.bfloat16)
        out[f"{rows}x{C}"] = round(chain_time(lambda i: ops.layernorm_bf16(x, g, b, out=y), 100), 2)
        out[f"{rows}x{C}_noaffine"] = round(chain_time(lambda i: ops.layernorm_bf16(x, None, None, out=y), 100), 2)
print(json.dumps(out))
