"""Cost of ONE kernel inside a dependent chain (what a small-latent tower layer is made of): each op is captured R times
back to back in a CUDA graph (programmatic dependent launch as in the real tower) and the replay time is divided by R.

    python tools/chain_bench.py [--rows 256] [--channels 1280] [--heads 8] [--qk 256] [--v 1280] [--reps 100]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from perceiverio_pytorch_b200 import engine, ops  # noqa: E402


def chain_time(fn, reps, iters=5):
    """fn(i) launches the op for chain position i.  Returns us per launch."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(3):
            fn(i)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(reps):
                fn(i)
        g.replay()
        s.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(iters):
            g.replay()
        b.record(s)
        s.synchronize()
    return a.elapsed_time(b) / (iters * reps) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=256)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--channels", type=int, default=1280)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--qk", type=int, default=256)
    ap.add_argument("--v", type=int, default=1280)
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--weights", type=int, default=48, help="distinct weight copies cycled through (HBM-resident weights)")
    a = ap.parse_args()
    dev = "cuda"
    M, C, H, QK, V = a.rows * a.batch, a.channels, a.heads, a.qk, a.v
    nqkv = 2 * QK + V
    torch.manual_seed(0)
    W = a.weights
    with torch.inference_mode():
        x = [torch.randn(M, C, device=dev) for _ in range(2)]
        xb = [t.to(ops.dtype16()) for t in x]
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        wqkv = [(0.02 * torch.randn(nqkv, C, device=dev)).to(ops.dtype16()) for _ in range(W)]
        wsq = [(0.02 * torch.randn(C, C, device=dev)).to(ops.dtype16()) for _ in range(W)]
        wo = [(0.02 * torch.randn(C, V, device=dev)).to(ops.dtype16()) for _ in range(W)]
        bias_qkv, bias_c = torch.zeros(nqkv, device=dev), torch.zeros(C, device=dev)
        qkv = torch.randn(M, ops.pad8(nqkv), device=dev).to(ops.dtype16())
        o16 = torch.randn(M, V, device=dev).to(ops.dtype16())
        y32 = [torch.empty(M, C, device=dev) for _ in range(2)]
        y16 = [torch.empty(M, ops.pad8(max(C, nqkv)), device=dev, dtype=ops.dtype16()) for _ in range(2)]
        out = {}
        out["layernorm"] = chain_time(lambda i: ops.layernorm_bf16(x[i & 1], g, b), a.reps)
        out["gemm_qkv"] = chain_time(lambda i: ops.gemm(xb[i & 1], wqkv[i % W], M=M, N=nqkv, K=C, bias=bias_qkv,
                                                        out_bf16=y16[i & 1], ldo16=y16[0].stride(0)), a.reps)
        out["attention"] = chain_time(lambda i: engine.attention(qkv, qkv.stride(0), 0, qkv, qkv.stride(0), QK, qkv,
                                                                 qkv.stride(0), 2 * QK, B=a.batch, H=H, Nq=a.rows,
                                                                 Nk=a.rows, dqk=QK // H, dv=V // H,
                                                                 scale=(QK // H) ** -0.5), a.reps)
        out["gemm_out_residual_f32"] = chain_time(lambda i: ops.gemm(o16, wo[i % W], M=M, N=C, K=V, bias=bias_c,
                                                                     residual=x[i & 1], ldr=C, out_f32=y32[i & 1],
                                                                     ldo32=C), a.reps)
        out["gemm_fc1_gelu"] = chain_time(lambda i: ops.gemm(xb[i & 1], wsq[i % W], M=M, N=C, K=C, bias=bias_c, act=1,
                                                             out_bf16=y16[i & 1], ldo16=y16[0].stride(0)), a.reps)
        out["gemm_fc1_gelu_same_weights"] = chain_time(lambda i: ops.gemm(xb[i & 1], wsq[0], M=M, N=C, K=C, bias=bias_c,
                                                                          act=1, out_bf16=y16[i & 1],
                                                                          ldo16=y16[0].stride(0)), a.reps)
        tiny = torch.randn(4, 64, device=dev)
        g64, b64 = torch.ones(64, device=dev), torch.zeros(64, device=dev)
        out["tiny_layernorm_4x64 (chain floor)"] = chain_time(lambda i: ops.layernorm_bf16(tiny, g64, b64), a.reps)
    print(json.dumps({"rows": M, "channels": C, "us_per_launch_in_chain": {k: round(v, 2) for k, v in out.items()},
                      "layer_sum_us": round(2 * out["layernorm"] + out["gemm_qkv"] + out["attention"] +
                                            out["gemm_out_residual_f32"] + 2 * out["gemm_fc1_gelu"], 1)}))


if __name__ == "__main__":
    main()
