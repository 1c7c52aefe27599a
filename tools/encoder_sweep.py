"""BASELINE.json configs[4]: synthetic scaling sweep of the encoder cross-attend (512 latents x 1024 channels attending
over Nk x 261 inputs, the ImageNet-pixels geometry), key axis sharded over the ranks with the LSE exchange of
perceiverio_pytorch_b200.parallel.KeyShard.

    python tools/encoder_sweep.py                       # 1 GPU
    torchrun --nproc-per-node N tools/encoder_sweep.py  # N GPUs, one rank per GPU (NCCL)

`--check` first verifies, on a reduced shape, that the key-sharded result equals the unsharded one computed on the same
rank (both through the CUDA kernels) and the fp32 oracle.  One JSON line per (Nk, B) point is printed by rank 0:
time is the max over ranks of the CUDA-event time of `steps` forwards."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import perceiverio_pytorch_b200 as pio  # noqa: E402
from perceiverio_pytorch_b200 import parallel  # noqa: E402


def xattn_flops(B, Nq, Nk, Cq, Ck, O):
    # reference-algorithm FLOPs of one CrossAttention block, qk = v = Ck
    return B * (2 * Nq * Cq * Ck + 2 * Nk * Ck * 2 * Ck + 2 * Nq * Nk * 2 * Ck + 2 * Nq * Ck * O + 4 * Nq * O * O)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--nk", type=int, nargs="*", default=[16384, 65536, 262144, 1048576])
    ap.add_argument("--batch", type=int, nargs="*", default=[1, 8])
    ap.add_argument("--channels", type=int, default=261)
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                    help="nccl: all_gather + combine; peer: combine kernel loads the peers' partials over NVLink")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    C = args.channels
    enc = pio.PerceiverEncoder(num_input_channels=C, num_self_attends_per_block=1, num_blocks=1, num_latents=512,
                               num_latent_channels=1024).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, prm in enc.named_parameters():
            if name.endswith("bias"):
                prm.copy_(0.05 * torch.randn(prm.shape, generator=g))
    enc = enc.to(dev)
    ca = enc.cross_attend

    if args.check:
        B, Nk = 2, 20000 + 37
        x = torch.randn(B, Nk, C, generator=torch.Generator().manual_seed(2)).to(dev)
        mask = torch.rand(B, Nk, generator=torch.Generator().manual_seed(3)) > 0.2
        mask[1, Nk // 2:] = False
        mask = mask.to(dev)
        with torch.inference_mode():
            lat = enc.latents(x)
            full, _ = ca._forward_factored(lat, x, key_mask=mask, row_keep=None)
            xs, ms, (b0, e0) = parallel.shard_keys(x, rank, world, mask, multiple=64)
            shard = parallel.KeyShard(local_splits=2, exchange=args.exchange)
            got, _ = ca._forward_factored(lat, xs.contiguous(), key_mask=ms.contiguous(), row_keep=None, shard=shard)
        err = float((got - full).abs().max() / full.abs().max())
        from oracle import perceiver_oracle as O
        p = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
        ref = O.cross_attention(p, "cross_attend.", 1, True, lat.cpu(), x.cpu(),
                                O.make_cross_attention_mask(torch.ones(B, 512, dtype=torch.bool), mask.cpu()))
        err_ref = float((got.cpu() - ref).abs().max() / ref.abs().max())
        t = torch.tensor([err, err_ref], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"check": "key-sharded encoder cross-attend", "world": world, "exchange": args.exchange,
                              "keys_rank0": [b0, e0],
                              "max_rel_err_vs_unsharded_cuda": float(t[0]), "max_rel_err_vs_fp32_oracle": float(t[1]),
                              "ok": bool(t[0] <= 5e-3 and t[1] <= 1e-2)}), flush=True)

    for Nk in args.nk:
        for B in args.batch:
            per = -(-Nk // world)
            per = -(-per // 64) * 64
            b0 = min(Nk, rank * per)
            e0 = min(Nk, b0 + per)
            if e0 - b0 <= 0 or B * (e0 - b0) * C * 4 > 40e9:
                continue
            xs = torch.randn(B, e0 - b0, C, device=dev)
            shard = parallel.KeyShard(exchange=args.exchange) if world > 1 else None
            with torch.inference_mode():
                lat = enc.latents(xs)
                for _ in range(2):
                    ca._forward_factored(lat, xs, key_mask=None, row_keep=None, shard=shard)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                    torch.cuda.synchronize()
                e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e_a.record()
                for _ in range(args.steps):
                    ca._forward_factored(lat, xs, key_mask=None, row_keep=None, shard=shard)
                e_b.record()
                torch.cuda.synchronize()
            t = torch.tensor([e_a.elapsed_time(e_b) / args.steps], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            if rank == 0:
                fl = xattn_flops(B, 512, Nk, 1024, C, 1024)
                print(json.dumps({"sweep": "encoder cross-attend, key axis sharded", "n_gpus": world,
                                  "exchange": args.exchange if world > 1 else "none", "Nk": Nk, "B": B,
                                  "channels": C, "ms": round(ms, 4), "samples_per_s": round(B / (ms * 1e-3), 2),
                                  "tflops_reference_algorithm": round(fl / (ms * 1e-3) / 1e12, 1),
                                  "input_gbs": round(B * Nk * C * 4 / (ms * 1e-3) / 1e9, 1)}), flush=True)
            del xs
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
