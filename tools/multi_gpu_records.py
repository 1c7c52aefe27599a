"""Multi-GPU records that bench.py appends to its JSON line (outside the timed headline region), one process per GPU:

* `key_shard` — BASELINE.json configs[4]: the encoder cross-attend (512 latents x 1024 channels over Nk x 261 inputs, the
  ImageNet-pixels geometry) with the KEY AXIS sharded over the ranks and the partial softmax statistics merged in one
  exchange step (perceiverio_pytorch_b200.parallel.KeyShard, peer-memory exchange when symmetric memory is available).
* `flow_b1` — the optical-flow config at batch 1, the case that cannot be batch-sharded: key-sharded encoder cross-attend
  -> latent tower replicated on every rank -> query-sharded decoder (no collective; an all_gather of the [Nq / W, 2]
  output rows is timed separately).

Every time is the max over ranks of a CUDA-event time.  Each record carries the time of the SAME work on one GPU
(measured on rank 0 alone in the same process, unsharded) so the scaling efficiency t1 / (W * tW) can be read off
one line; at W = 1 only the single-GPU numbers exist.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import perceiverio_pytorch_b200 as pio  # noqa: E402
from perceiverio_pytorch_b200 import parallel  # noqa: E402


def _time(fn, steps, world, dev):
    """Max over ranks of the mean CUDA-event time of fn() (ms)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _graphed(fn, example):
    """fn replayed from a CUDA graph (the batch-1 subsystems are bound by the host's launch rate when launched eagerly);
    falls back to the eager callable if the capture fails (e.g. a collective that cannot be captured)."""
    from perceiverio_pytorch_b200.graph import GraphedForward
    try:
        g = GraphedForward(fn, [example], warmup=1)
        return (lambda: g(g.inputs[0])), True
    except Exception:
        torch.cuda.synchronize()
        return (lambda: fn(example)), False


def _perturb(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if name.endswith("bias"):
                prm.copy_(0.05 * torch.randn(prm.shape, generator=g))


def _exchange_mode(world):
    if world == 1:
        return "none"
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
        return "peer"
    except Exception:
        return "nccl"


def key_shard_record(world, rank, dev, nks=(65536, 1048576), batches=(1, 8), steps=5):
    C = 261
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(num_input_channels=C, num_self_attends_per_block=1, num_blocks=1, num_latents=512,
                               num_latent_channels=1024).eval()
    _perturb(enc, 1)
    enc = enc.to(dev)
    ca = enc.cross_attend
    mode = _exchange_mode(world)
    points = []
    for Nk in nks:
        for B in batches:
            per = -(-(-(-Nk // world)) // 64) * 64
            b0 = min(Nk, rank * per)
            e0 = min(Nk, b0 + per)
            xs = torch.randn(B, e0 - b0, C, device=dev)
            shard = None
            used = "none"
            if world > 1:
                try:
                    shard = parallel.KeyShard(exchange=mode)
                    with torch.inference_mode():
                        ca._forward_factored(enc.latents(xs), xs, shard=shard)
                    used = mode
                except Exception:
                    shard = parallel.KeyShard(exchange="nccl")
                    used = "nccl (peer-memory exchange unavailable)"
            with torch.inference_mode():
                lat = enc.latents(xs)
                ms = _time(lambda: ca._forward_factored(lat, xs, shard=shard), steps, world, dev)
            one = None
            del xs
            if world > 1:
                # the same total work on ONE GPU (rank 0, unsharded), the other ranks wait
                if rank == 0 and B * Nk * C * 4 < 60e9:
                    xf = torch.randn(B, Nk, C, device=dev)
                    with torch.inference_mode():
                        latf = enc.latents(xf)
                        one = _time(lambda: ca._forward_factored(latf, xf), steps, 1, dev)
                    del xf
                dist.barrier()
            fl = B * (2 * 512 * 1024 * C + 2 * Nk * C * 2 * C + 2 * 512 * Nk * 2 * C + 2 * 512 * C * 1024 + 4 * 512 * 1024 * 1024)
            pt = {"Nk": Nk, "B": B, "ms": round(ms, 4), "exchange": used,
                  "tflops_reference_algorithm": round(fl / (ms * 1e-3) / 1e12, 1)}
            if one is not None:
                pt["ms_one_gpu"] = round(one, 4)
                pt["efficiency"] = round(one / (world * ms), 3)
            points.append(pt)
            torch.cuda.empty_cache()
    return {"what": "encoder cross-attend 512x1024 latents over Nk x 261 inputs, key axis sharded over the ranks, one "
                    "exchange step (packed (O, m, l) partials merged by the combine kernel)", "n_gpus": world,
            "points": points}


def flow_b1_record(world, rank, dev, steps=5, precision="fp16"):
    from perceiverio_pytorch_b200 import engine
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from bench_configs import CONFIGS, perturb
    cfg = CONFIGS["flow"]
    torch.manual_seed(0)
    enc = pio.PerceiverEncoder(**cfg["enc"]).eval()
    dec = pio.PerceiverDecoder(**cfg["dec"]).eval()
    perturb(enc, 1)
    perturb(dec, 2)
    enc, dec = enc.to(dev), dec.to(dev)
    enc.precision = dec.precision = precision
    Nk = cfg["Nk"]
    x = torch.randn(1, Nk, 322, device=dev, generator=torch.Generator(device=dev).manual_seed(7))   # same on every rank
    mode = _exchange_mode(world)
    engine.set_precision(precision)     # also inside the graph captures below
    try:
        return _flow_b1(world, rank, dev, steps, precision, enc, dec, x, Nk, mode)
    finally:
        engine.set_precision("bf16")


def _flow_b1(world, rank, dev, steps, precision, enc, dec, x, Nk, mode):
    with torch.inference_mode():
        lat = enc.latents(x)
        z0 = enc.cross_attend._forward_factored(lat, x)[0]

        def tower(z):
            for sa in enc.self_attends:
                z = sa(z)
            return z
        z1 = tower(z0)
    # ---- one GPU, unsharded, replayed from CUDA graphs (every rank computes it: the references for the subsystems) ----
    f_enc1, _ = _graphed(lambda xx: enc.cross_attend._forward_factored(enc.latents(xx), xx)[0], x)
    f_tower, _ = _graphed(tower, z0)
    f_dec1, _ = _graphed(lambda zz: dec(x, zz), z1)
    with torch.inference_mode():
        t_enc1, t_tower, t_dec1 = _time(f_enc1, steps, 1, dev), _time(f_tower, steps, 1, dev), _time(f_dec1, steps, 1, dev)
        rec = {"what": "FlowPerceiver hot path at batch 1 (182,528 inputs / queries): key-sharded encoder cross-attend -> "
                       "replicated latent tower -> query-sharded decoder", "n_gpus": world, "precision": precision,
               "ms_one_gpu": {"encoder_xattn": round(t_enc1, 4), "tower": round(t_tower, 4), "decoder": round(t_dec1, 4),
                              "sum": round(t_enc1 + t_tower + t_dec1, 4)}}
        if world == 1:
            return rec
        # ---- W GPUs ----
        xs, _, (kb, ke) = parallel.shard_keys(x, rank, world, None, multiple=64)
        xs = xs.contiguous()
        used = mode
        try:
            shard = parallel.KeyShard(exchange=mode)
            enc.cross_attend._forward_factored(lat, xs, shard=shard)
        except Exception:
            shard = parallel.KeyShard(exchange="nccl")
            used = "nccl (peer-memory exchange unavailable)"
        zs = enc.cross_attend._forward_factored(lat, xs, shard=shard)[0]
        err_enc = float((zs - z0).abs().max() / z0.abs().max())
        qs, _, (qb, qe) = parallel.shard_queries(x, rank, world)
        qs = qs.contiguous()
        out_s = dec(qs, z1).clone()
        full = dec(x, z1)
        err_dec = float((out_s - full[:, qb:qe]).abs().max() / full.abs().max())
    f_encw, enc_graphed = _graphed(lambda xx: enc.cross_attend._forward_factored(enc.latents(xx), xx, shard=shard)[0], xs)
    f_decw, _ = _graphed(lambda zz: dec(qs, zz), z1)
    with torch.inference_mode():
        per = -(-(-(-Nk // world)) // 128) * 128
        gathered = torch.empty((world, per, 2), dtype=torch.float32, device=dev)
        pad = torch.zeros((per, 2), dtype=torch.float32, device=dev)

        def f_gather():
            pad[: qe - qb] = out_s[0]
            dist.all_gather_into_tensor(gathered, pad)
        t_encw, t_decw, t_gather = _time(f_encw, steps, world, dev), _time(f_decw, steps, world, dev), _time(f_gather, steps, world, dev)

        f_all, all_graphed = _graphed(lambda xx: dec(qs, tower(enc.cross_attend._forward_factored(
            enc.latents(xx), xx, shard=shard)[0])), xs)
        t_all = _time(f_all, steps, world, dev)
    errs = torch.tensor([err_enc, err_dec], device=dev)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    rec.update({
        "exchange": used, "cuda_graph": {"sharded_encoder": enc_graphed, "forward": all_graphed},
        "ms": {"encoder_xattn": round(t_encw, 4), "tower_replicated": round(t_tower, 4), "decoder": round(t_decw, 4),
               "output_all_gather": round(t_gather, 4), "forward": round(t_all, 4)},
        "efficiency": {"encoder_xattn": round(t_enc1 / (world * t_encw), 3), "decoder": round(t_dec1 / (world * t_decw), 3),
                       "forward_speedup": round((t_enc1 + t_tower + t_dec1) / t_all, 3),
                       "forward_speedup_bound_amdahl": round((t_enc1 + t_tower + t_dec1) /
                                                             (t_enc1 / world + t_tower + t_dec1 / world), 3)},
        "sharded_vs_unsharded_max_rel_err": {"encoder_latents": float(errs[0]), "decoder_rows": float(errs[1])}})
    return rec
