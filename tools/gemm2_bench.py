import torch, sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
from perceiverio_pytorch_b200 import ops
from chain_bench import chain_time
M, C = 32768, 1024
dev = 'cuda'
with torch.inference_mode():
    x = [torch.randn(M, C, device=dev) for _ in range(2)]
    a = [torch.randn(M, C, device=dev).to(torch.bfloat16) for _ in range(2)]
    w = [(0.02 * torch.randn(C, C, device=dev)).to(torch.bfloat16) for _ in range(4)]
    w3 = [(0.02 * torch.randn(3 * C, C, device=dev)).to(torch.bfloat16) for _ in range(2)]
    b = torch.zeros(C, device=dev); b3 = torch.zeros(3 * C, device=dev)
    y32 = [torch.empty(M, C, device=dev) for _ in range(2)]
    y16 = [torch.empty(M, C, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    q16 = [torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    st = ops.empty_row_stats(M, C, dev)
    out = {}
    out["producer_res_f32_raw_stats"] = chain_time(lambda i: ops.gemm(a[i & 1], w[i % 4], M=M, N=C, K=C, bias=b, residual=x[i & 1], ldr=C, out_f32=y32[i & 1], ldo32=C, out_bf16=y16[i & 1], ldo16=C, row_stats_out=st), 20)
    hi = [torch.empty(M, C, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    lo = [torch.empty(M, C, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    out["producer_pair_in_pair_out_stats"] = chain_time(lambda i: ops.gemm(a[i & 1], w[i % 4], M=M, N=C, K=C, bias=b, residual_hi16=hi[i & 1], residual_lo16=lo[i & 1], ldr16=C, out_bf16=hi[1 - (i & 1)], ldo16=C, out_lo16=lo[1 - (i & 1)], row_stats_out=st), 20)
    out["producer_res_f32"] = chain_time(lambda i: ops.gemm(a[i & 1], w[i % 4], M=M, N=C, K=C, bias=b, residual=x[i & 1], ldr=C, out_f32=y32[i & 1], ldo32=C), 20)
    out["producer_f32_only"] = chain_time(lambda i: ops.gemm(a[i & 1], w[i % 4], M=M, N=C, K=C, bias=b, out_f32=y32[i & 1], ldo32=C), 20)
    out["consumer_fc1_gelu"] = chain_time(lambda i: ops.gemm(a[i & 1], w[i % 4], M=M, N=C, K=C, bias=b, act=1, out_bf16=y16[i & 1], ldo16=C), 20)
    out["consumer_qkv"] = chain_time(lambda i: ops.gemm(a[i & 1], w3[i & 1], M=M, N=3 * C, K=C, bias=b3, out_bf16=q16[i & 1], ldo16=3 * C), 20)
    # correctness of the producer against torch
    ops.gemm(a[0], w[0], M=M, N=C, K=C, bias=b, residual=x[0], ldr=C, out_f32=y32[0], ldo32=C, out_bf16=y16[0], ldo16=C, row_stats_out=st)
    ref = a[0][:4096].float() @ w[0].float().t() + x[0][:4096]
    err = float((y32[0][:4096] - ref).abs().max() / ref.abs().max())
    err16 = float((y16[0][:4096].float() - ref).abs().max() / ref.abs().max())
    ssum = float((st[:4096].sum(1)[:, 0] - ref.sum(1)).abs().max() / ref.sum(1).abs().max())
print(json.dumps({"us": {k: round(v, 1) for k, v in out.items()}, "tflops": {k: round(2 * M * C * (3 * C if 'qkv' in k else C) / v / 1e6, 0) for k, v in out.items()}, "err": [err, err16, ssum]}))
