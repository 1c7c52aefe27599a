"""Developer aid: build the library with -DPIO_DECODE_TRACE, run the optical-flow decoder attention once and print the
pipeline timeline of the first CTA pair (clock64 ticks per SM).  Usage (GPU box):  python tools/trace_decode.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    if os.environ.get("PIO_TRACE_CHILD") != "1":
        env = dict(os.environ, PIO_NVCC_EXTRA="-DPIO_DECODE_TRACE", PIO_TRACE_CHILD="1")
        subprocess.run([sys.executable, "-m", "perceiverio_pytorch_b200.build", "--force"], check=True, env=env, cwd=ROOT)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, cwd=ROOT)
        subprocess.run([sys.executable, "-m", "perceiverio_pytorch_b200.build", "--force"], check=True, cwd=ROOT)
        sys.exit(r.returncode)
    import torch
    from perceiverio_pytorch_b200 import ops
    nq, nk, dqk, dv = 182528, 2048, 323, 322
    q = torch.randn(nq, 328, device="cuda").to(torch.bfloat16)
    kv = torch.randn(nk, 656, device="cuda").to(torch.bfloat16)
    ops.decoder_attention(q, kv, kv.view(-1)[328:], B=1, Nq=nq, Nk=nk, dqk=dqk, dv=dv, ldq=328, ldk=656, ldv=656,
                          strideQ=0, strideK=0, strideV=0, scale=dqk ** -0.5)
    torch.cuda.synchronize()
