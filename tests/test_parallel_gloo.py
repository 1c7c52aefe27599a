"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes exercise the key-axis shard plumbing of
perceiverio_pytorch_b200.parallel (slicing, packed all_gather, mask OR) with the oracle standing in for the CUDA
kernels that produce / merge the partial attention results."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from perceiverio_pytorch_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import perceiver_oracle as O
        torch.manual_seed(0)  # same data on every rank
        B, H, Nq, Nk, dqk, dv = 2, 2, 9, 300, 8, 5
        q = torch.randn(B, Nq, H, dqk, dtype=torch.float64) * 3
        k = torch.randn(B, Nk, H, dqk, dtype=torch.float64)
        v = torch.randn(B, Nk, H, dv, dtype=torch.float64)
        km = torch.rand(B, Nk) > 0.3
        km[0, 128:] = False          # rank 1 holds only masked keys for sample 0
        km[1, :] = False             # sample 1 has no valid key anywhere -> wiped rows
        full = O.attend(q, k, v, O.make_cross_attention_mask(torch.ones(B, Nq, dtype=torch.bool), km))

        shard = parallel.KeyShard()
        assert shard.world == world and shard.rank == rank
        kk, kmask, (b0, e0) = parallel.shard_keys(k, rank, world, km, multiple=128)
        vv = v[:, b0:e0]
        assert (b0, e0) == ((0, 256) if rank == 0 else (256, 300))
        any_key = shard.any_over_ranks(kmask.any(dim=1, keepdim=True))
        assert any_key.flatten().tolist() == [True, False]

        Op, m, l = O.attend_partial(q, kk, vv, kmask)              # [B,H,Nq,dv], [B,H,Nq]
        rows = B * H * Nq
        packed = parallel.pack_partial(Op.reshape(rows, dv), m, l)
        allp = shard.gather_packed(packed)
        assert allp.shape == (world, rows * (dv + 2))
        parts = []
        for w in range(world):
            o_w, m_w, l_w = parallel.packed_views(allp[w], rows, dv)
            parts.append((o_w.view(B, H, Nq, dv), m_w.view(B, H, Nq), l_w.view(B, H, Nq)))
        merged = O.combine_partials(parts).permute(0, 2, 1, 3).reshape(B, Nq, H * dv)
        merged = merged * any_key[:, :, None]                       # row_keep
        err = float((merged - full).abs().max())
        ret[rank] = err
    finally:
        dist.destroy_process_group()


def test_key_shard_plumbing_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] < 1e-12, dict(ret)


def test_shard_range_covers_axis():
    for n, world, mult in [(50176, 8, 64), (182528, 8, 64), (2048, 4, 128), (1000, 8, 128), (300, 2, 128)]:
        covered = []
        for r in range(world):
            b, e = parallel.shard_range(n, r, world, mult)
            assert b % mult == 0 or b == n
            covered.extend(range(b, e))
        assert covered == list(range(n))


def test_shard_queries_and_batch():
    q = torch.arange(2 * 1000 * 3).view(2, 1000, 3)
    parts = [parallel.shard_queries(q, r, 4)[0] for r in range(4)]
    assert torch.equal(torch.cat(parts, 1), q)
    x = torch.arange(64 * 2).view(64, 2)
    assert torch.equal(torch.cat([parallel.shard_batch(x, r, 8) for r in range(8)]), x)
