"""Developer aid: build the library with -DPIO_FLASHQT_TRACE, run the ImageNet-pixels encoder attention once and print the
pipeline timeline of CTA 0 (clock64 ticks, one time base for the MMA issuer and softmax warp 4).
Usage (GPU box):  python tools/trace_flash_qt.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    if os.environ.get("PIO_TRACE_CHILD") != "1":
        env = dict(os.environ, PIO_NVCC_EXTRA="-DPIO_FLASHQT_TRACE", PIO_TRACE_CHILD="1")
        subprocess.run([sys.executable, "-m", "perceiverio_pytorch_b200.build", "--force"], check=True, env=env, cwd=ROOT)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, cwd=ROOT)
        subprocess.run([sys.executable, "-m", "perceiverio_pytorch_b200.build", "--force"], check=True, cwd=ROOT)
        sys.exit(r.returncode)
    import torch
    from perceiverio_pytorch_b200 import ops
    B, nq, nk, d = 16, 512, 50176, 261
    q = torch.randn(nq, 264, device="cuda").to(torch.bfloat16)
    kv = torch.randn(B * nk, 264, device="cuda").to(torch.bfloat16)
    ops.attention_fwd(q, kv, kv, B=B, H=1, Nq=nq, Nk=nk, dqk=d, dv=d, strideQ=0, strideK=nk * 264, strideV=nk * 264,
                      ldq=264, ldk=264, ldv=264, num_splits=2)
    torch.cuda.synchronize()
