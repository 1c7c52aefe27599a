"""Determinism stress for the persistent two-tile attention kernel: the same launch repeated many times must give
bit-identical output (any difference is a pipeline race).  Usage (GPU box): python tools/stress_flash2.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from perceiverio_pytorch_b200 import ops  # noqa: E402

CASES = [  # B, H, Nq, Nk, dqk, dv
    (1, 16, 2048, 2048, 32, 32),      # flow tower
    (64, 8, 512, 512, 128, 128),      # classification tower
    (1, 8, 256, 256, 32, 160),        # language tower
    (1, 8, 784, 784, 64, 64),         # multimodal tower
    (3, 4, 300, 100, 64, 64),         # one key tile per item
    (2, 8, 130, 1000, 128, 128),      # ragged
    (1, 8, 2048, 256, 32, 96),        # language decoder
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    torch.manual_seed(0)
    bad = 0
    for (B, H, Nq, Nk, dqk, dv) in CASES:
        q = (torch.randn(B * Nq, H * dqk, device="cuda") * 2).to(torch.bfloat16)
        k = (torch.randn(B * Nk, H * dqk, device="cuda") * 2).to(torch.bfloat16)
        v = torch.randn(B * Nk, H * dv, device="cuda").to(torch.bfloat16)
        run = lambda: ops.attention_fwd(q.view(-1), k.view(-1), v.view(-1), B=B, H=H, Nq=Nq, Nk=Nk, dqk=dqk, dv=dv,
                                        strideQ=Nq * H * dqk, strideK=Nk * H * dqk, strideV=Nk * H * dv,
                                        ldq=H * dqk, ldk=H * dqk, ldv=H * dv)
        ref = run().clone()
        # fp32 reference
        qf = q.float().view(B, Nq, H, dqk).permute(0, 2, 1, 3)
        kf = k.float().view(B, Nk, H, dqk).permute(0, 2, 1, 3)
        vf = v.float().view(B, Nk, H, dv).permute(0, 2, 1, 3)
        o = torch.softmax(qf @ kf.transpose(-1, -2) / dqk ** 0.5, -1) @ vf
        o = o.permute(0, 2, 1, 3).reshape(B, Nq, H * dv)
        err = float((ref[:, :, :H * dv].float() - o).abs().max() / o.abs().max())
        mism = 0
        for i in range(reps):
            out = run()
            if not torch.equal(out, ref):
                mism += 1
                if mism == 1:
                    d = (out.float() - ref.float()).abs()
                    idx = torch.nonzero(d > 0)
                    print(f"  first mismatch at rep {i}: {idx.shape[0]} elements, max {float(d.max()):.3e}, "
                          f"first idx {idx[0].tolist()} last idx {idx[-1].tolist()}")
        torch.cuda.synchronize()
        print(f"B={B} H={H} Nq={Nq} Nk={Nk} dqk={dqk} dv={dv}: err vs fp32 {err:.2e}, {mism}/{reps} mismatching repeats")
        bad += mism
    print("STRESS", "FAIL" if bad else "OK")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
