"""Developer aid: build the library with -DPIO_GEMM_TRACE, run a dependent chain of small GEMMs from a CUDA graph and
print CTA 0's timeline of the LAST launch (ns from kernel entry; clock64 ticks beside it).
Usage (GPU box):  python tools/trace_gemm.py [M,N,K[,bf16out|f32[,tile_n[,cluster_m]]]] ..."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    if os.environ.get("PIO_TRACE_CHILD") != "1":
        env = dict(os.environ, PIO_NVCC_EXTRA="-DPIO_GEMM_TRACE", PIO_TRACE_CHILD="1")
        subprocess.run([sys.executable, "-m", "perceiverio_pytorch_b200.build", "--force"], check=True, env=env, cwd=ROOT)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, cwd=ROOT)
        subprocess.run([sys.executable, "-m", "perceiverio_pytorch_b200.build", "--force"], check=True, cwd=ROOT)
        sys.exit(r.returncode)
    import torch
    from perceiverio_pytorch_b200 import _lib, ops

    def run(spec):
        f = spec.split(",")
        M, N, K = (int(v) for v in f[:3])
        mode = f[3] if len(f) > 3 else "bf16out"
        tile_n = int(f[4]) if len(f) > 4 else 0
        cluster_m = int(f[5]) if len(f) > 5 else None
        dev = "cuda"
        xb = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(2)]
        ws = [(0.02 * torch.randn(N, K, device=dev)).to(torch.bfloat16) for _ in range(48)]
        bias = torch.zeros(N, device=dev)
        res = torch.randn(M, N, device=dev)
        y16 = [torch.empty(M, ops.pad8(N), device=dev, dtype=torch.bfloat16) for _ in range(2)]
        y32 = [torch.empty(M, N, device=dev) for _ in range(2)]

        def launch(i):
            if mode == "bf16out":
                ops.gemm(xb[i & 1], ws[i % 48], M=M, N=N, K=K, bias=bias, act=1, out_bf16=y16[i & 1], ldo16=y16[0].stride(0),
                         tile_n=tile_n, cluster_m=cluster_m, kernel=1 if tile_n else None)
            else:
                ops.gemm(xb[i & 1], ws[i % 48], M=M, N=N, K=K, bias=bias, residual=res, ldr=N, out_f32=y32[i & 1], ldo32=N,
                         tile_n=tile_n, cluster_m=cluster_m, kernel=1 if tile_n else None)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s), torch.inference_mode():
            for i in range(3):
                launch(i)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for i in range(40):
                    launch(i)
            for _ in range(3):
                g.replay()
            s.synchronize()
        buf = (ctypes.c_ulonglong * (4 * 128 * 3))()
        lib = _lib.load()
        lib.pio_debug_gemm_trace.argtypes = [ctypes.c_void_p]
        lib.pio_debug_gemm_trace(buf)
        ev = []
        for slot in range(4):
            for i in range(128):
                tag, clk, gt = buf[(slot * 128 + i) * 3: (slot * 128 + i) * 3 + 3]
                if tag:
                    ev.append((gt, clk, slot, tag))
        ev.sort(key=lambda e: e[1])
        c0, g0 = ev[0][1], min(e[0] for e in ev if e[0])
        print(f"GEMM {M}x{N}x{K} [{mode}] tile_n={tile_n} cluster_m={cluster_m}, CTA 0 of the last launch of a 40-launch chain: tag, clock64 ticks, globaltimer ns")
        brief = os.environ.get("PIO_TRACE_BRIEF") == "1"
        for gt, clk, slot, tag in ev:
            if brief and not (tag < 10 or tag in (100, 200, 201, 205, 210, 219, 299, 400) or 300 <= tag < 400):
                continue
            print(f"  slot {slot} tag {tag:4d}  {clk - c0:8d} clk" + (f"  {gt - g0:8d} ns" if gt else ""))

    for spec in (sys.argv[1:] or ["256,1280,1280"]):
        run(spec)
