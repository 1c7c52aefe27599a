"""Developer aid: build the library with -DPIO_GEMM2_TRACE, run one producer GEMM of the bench tower (pair stream in /
out, raw copy, row statistics) and print the epilogue timeline of warp 4 of CTA 0 (clock64 ticks): per 32-column chunk
c: 310+c accumulator in registers, 320+c bias / activation done, 330+c residual arrived, 340+c residual added + statistics,
350+c staging slot free, 360+c staging written, 370+c fence + warp sync, 380+c stores issued.
Usage (GPU box):  python tools/trace_gemm2.py"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    if os.environ.get("PIO_TRACE_CHILD") != "1":
        env = dict(os.environ, PIO_NVCC_EXTRA="-DPIO_GEMM2_TRACE " + os.environ.get("PIO_TRACE_EXTRA", ""), PIO_TRACE_CHILD="1")
        subprocess.run([sys.executable, "-m", "perceiverio_pytorch_b200.build", "--force"], check=True, env=env, cwd=ROOT)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, cwd=ROOT)
        subprocess.run([sys.executable, "-m", "perceiverio_pytorch_b200.build", "--force"], check=True, cwd=ROOT)
        sys.exit(r.returncode)
    import torch
    from perceiverio_pytorch_b200 import _lib, ops
    M, C = 32768, 1024
    dev = "cuda"
    mode = sys.argv[1] if len(sys.argv) > 1 else "pair"
    a = torch.randn(M, C, device=dev).to(torch.bfloat16)
    w = (0.02 * torch.randn(C, C, device=dev)).to(torch.bfloat16)
    b = torch.zeros(C, device=dev)
    x = torch.randn(M, C, device=dev)
    hi = [torch.randn(M, C, device=dev).to(torch.bfloat16) for _ in range(2)]
    lo = [(0.01 * torch.randn(M, C, device=dev)).to(torch.bfloat16) for _ in range(2)]
    y32 = torch.empty(M, C, device=dev)
    st = ops.empty_row_stats(M, C, dev).zero_()
    st[:, 0, 1] = C        # unit variance, zero mean for the consumer modes
    lib = _lib.load()
    lib.pio_debug_gemm2_trace.argtypes = [ctypes.c_void_p]
    buf = (ctypes.c_ulonglong * (1024 * 2))()

    def launch():
        if mode == "fc1":       # consumer: fused-LayerNorm statistics in, GELU, 16-bit out
            ops.gemm(hi[0], w, M=M, N=C, K=C, bias=b, act=1, out_bf16=hi[1], ldo16=C, row_stats_in=st, ln_colsum=b,
                     ln_channels=C, ln_eps=1e-5, reverse_tiles=True)
        elif mode == "qkv":
            ops.gemm(hi[0], w, M=M, N=C, K=C, bias=b, out_bf16=hi[1], ldo16=C, row_stats_in=st, ln_colsum=b,
                     ln_channels=C, ln_eps=1e-5)
        elif mode == "pair":
            ops.gemm(a, w, M=M, N=C, K=C, bias=b, residual_hi16=hi[0], residual_lo16=lo[0], ldr16=C, out_bf16=hi[1], ldo16=C,
                     out_lo16=lo[1], row_stats_out=st)
        else:
            ops.gemm(a, w, M=M, N=C, K=C, bias=b, residual=x, ldr=C, out_f32=y32, ldo32=C, out_bf16=hi[1], ldo16=C,
                     row_stats_out=st)
    with torch.inference_mode():
        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        lib.pio_debug_gemm2_trace(buf)
        launch()
        torch.cuda.synchronize()
        lib.pio_debug_gemm2_trace(buf)
    mma = [(buf[2 * i], buf[2 * i + 1]) for i in range(512, 1024) if buf[2 * i]]
    if mma:
        print("MMA issuer of CTA 0 (500 waits for a free accumulator, 501 has it, 502 tile issued): tag, clk, delta")
        prev = mma[0][1]
        for tag, clk in mma[:40]:
            print(f"  {tag:4d} {clk - mma[0][1]:9d} {clk - prev:7d}")
            prev = clk
    ev = [(buf[2 * i], buf[2 * i + 1]) for i in range(512) if buf[2 * i]]
    t0 = ev[0][1]
    print(f"producer GEMM {M}x{C}x{C} [{mode}]: epilogue of warp 4, CTA 0 (tag, clk since the first tag, delta)")
    prev = t0
    for tag, clk in ev[:120]:
        print(f"  {tag:4d} {clk - t0:9d} {clk - prev:7d}")
        prev = clk
