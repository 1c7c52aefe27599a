"""Developer aid: build the library with -DPIO_FLASH2_TRACE, run the tower attention once and print CTA 0's pipeline
timeline (clock64 ticks).  Usage (GPU box):  python tools/trace_flash2.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    if os.environ.get("PIO_TRACE_CHILD") != "1":
        env = dict(os.environ, PIO_NVCC_EXTRA="-DPIO_FLASH2_TRACE", PIO_TRACE_CHILD="1")
        subprocess.run([sys.executable, "-m", "perceiverio_pytorch_b200.build", "--force"], check=True, env=env, cwd=ROOT)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env, cwd=ROOT)
        subprocess.run([sys.executable, "-m", "perceiverio_pytorch_b200.build", "--force"], check=True, cwd=ROOT)
        sys.exit(r.returncode)
    import torch
    from perceiverio_pytorch_b200 import ops
    if len(sys.argv) > 1 and sys.argv[1] == "flow":      # the optical-flow tower: 16 heads of 32 channels over 2048 latents
        qkv = torch.randn(2048, 1536, device="cuda").to(torch.bfloat16)
        qv, kv_, vv = qkv.view(-1), qkv.view(-1)[512:], qkv.view(-1)[1024:]
        ops.attention_fwd(qv, kv_, vv, B=1, H=16, Nq=2048, Nk=2048, dqk=32, dv=32, strideQ=0, strideK=0, strideV=0,
                          ldq=1536, ldk=1536, ldv=1536)
        torch.cuda.synchronize()
        sys.exit(0)
    qkv = torch.randn(64 * 512, 3072, device="cuda").to(torch.bfloat16)
    qv, kv_, vv = qkv.view(-1), qkv.view(-1)[1024:], qkv.view(-1)[2048:]
    ops.attention_fwd(qv, kv_, vv, B=64, H=8, Nq=512, Nk=512, dqk=128, dv=128, strideQ=512 * 3072, strideK=512 * 3072,
                      strideV=512 * 3072, ldq=3072, ldk=3072, ldv=3072)
    torch.cuda.synchronize()
