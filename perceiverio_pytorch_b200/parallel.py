"""Multi-GPU partitioning of the hot path on one NVSwitch box (one process per GPU, torch.distributed/NCCL).

SURVEY.md §8(e): the path shards naturally in three ways, and only one of them needs an exchange step:

* latent tower / whole model — batch axis, no collective (`shard_batch`);
* decoder cross-attend — query axis, no collective (`shard_queries`; an all_gather only if the caller wants the
  full output on every rank);
* encoder cross-attend — key axis: every rank streams its slice of the input array through the attention kernel
  and produces an un-normalised partial (O, m, l) per latent, packed as one fp32 buffer of B*H*Nq*(dv+2) floats
  (0.54 MB per sample for the ImageNet geometry — latency bound on NVSwitch).  The exchange step (`KeyShard`) is
  either
    - `exchange="peer"`: the packed partial lives in symmetric (peer-mapped) memory; after a symmetric-memory barrier
      ONE combine kernel on every rank loads all ranks' partials straight through NVLink and merges them (log-sum-exp)
      — the collective is fused into the consumer, nothing is gathered or copied; or
    - `exchange="nccl"`: one all_gather of the packed buffers followed by the same combine kernel on the gathered
      copy (the baseline, and what the gloo tests exercise on CPU).

The collective plumbing is backend-agnostic (it is exercised with gloo on CPU tensors in tests/); the kernels that
produce and merge the partials are the sm_100a ones.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int, multiple: int = 1) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of an axis of length n for `rank`, boundaries rounded to `multiple`
    (e.g. the key-tile width) so that every rank's slice starts on a tile boundary."""
    per = -(-n // world)
    per = -(-per // multiple) * multiple
    b = min(n, rank * per)
    return b, min(n, b + per)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    b, e = shard_range(x.shape[0], rank, world)
    return x[b:e]


def shard_queries(query: torch.Tensor, rank: int, world: int, query_mask: Optional[torch.Tensor] = None):
    """Query-axis slice for the decoder (no communication needed: every output row depends on its own query only)."""
    b, e = shard_range(query.shape[1], rank, world, 128)
    return query[:, b:e], (query_mask[:, b:e] if query_mask is not None else None), (b, e)


def shard_keys(inputs: torch.Tensor, rank: int, world: int, input_mask: Optional[torch.Tensor] = None,
               multiple: int = 128):
    """Key-axis slice of the encoder input array [B, Nk, C] (+ its mask)."""
    b, e = shard_range(inputs.shape[1], rank, world, multiple)
    return inputs[:, b:e], (input_mask[:, b:e] if input_mask is not None else None), (b, e)


def pack_partial(O: torch.Tensor, m: torch.Tensor, l: torch.Tensor) -> torch.Tensor:
    """(O [R, dv], m [R], l [R]) -> one flat fp32 buffer [R*dv | R | R] (R = B*H*Nq rows)."""
    return torch.cat([O.reshape(-1), m.reshape(-1), l.reshape(-1)])


def packed_views(buf: torch.Tensor, rows: int, dv: int):
    """Views (O [rows, dv], m [rows], l [rows]) into a packed partial buffer."""
    o = buf[: rows * dv].view(rows, dv)
    m = buf[rows * dv: rows * dv + rows]
    l = buf[rows * dv + rows: rows * dv + 2 * rows]
    return o, m, l


class KeyShard:
    """State of a key-axis shard of the encoder cross-attend across the ranks of `group`."""

    def __init__(self, group=None, local_splits: int = 0, exchange: str = "nccl"):
        if exchange not in ("nccl", "peer"):
            raise ValueError("exchange must be 'nccl' or 'peer'")
        self.group = group
        self.local_splits = local_splits   # key splits inside this rank's slice; 0 = pick to fill the SMs
        self.exchange = exchange
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._symm = {}                    # packed size -> [[(buffer, handle), (buffer, handle)], next index]

    # -- collectives ------------------------------------------------------------------------------------------
    def any_over_ranks(self, flag: torch.Tensor) -> torch.Tensor:
        """Logical OR of a bool tensor over the ranks ("does any rank hold a valid key for this sample")."""
        if self.world == 1:
            return flag
        t = flag.to(torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t.to(torch.bool)

    def gather_packed(self, packed: torch.Tensor) -> torch.Tensor:
        """all_gather of every rank's packed partial -> [world, n]."""
        if self.world == 1:
            return packed[None]
        out = torch.empty(self.world * packed.numel(), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(out, packed.contiguous(), group=self.group)
        return out.view(self.world, packed.numel())

    def _peer_buffer(self, n: int, device):
        """This rank's packed-partial buffer in symmetric memory (+ its rendezvous handle).  Two buffers are used
        alternately: a rank only overwrites a buffer two exchanges later, i.e. after a barrier that every peer can only
        have reached once it finished reading — so ONE barrier per exchange is enough."""
        import torch.distributed._symmetric_memory as symm_mem
        entry = self._symm.get(n)
        if entry is None:
            group = self.group if self.group is not None else dist.group.WORLD
            pairs = []
            for _ in range(2):
                t = symm_mem.empty(n, dtype=torch.float32, device=device)
                pairs.append((t, symm_mem.rendezvous(t, group.group_name)))
            entry = [pairs, 0]
            self._symm[n] = entry
        t, h = entry[0][entry[1]]
        entry[1] ^= 1
        return t, h

    # -- the exchange step ------------------------------------------------------------------------------------
    def combine(self, parts, row_keep=None, want_alive: bool = False):
        """parts = (O_part [S,B,H,Nq,dv], m [S,B,H,Nq], l [S,B,H,Nq]) from this rank's attention kernel.
        Returns the normalised attention output bf16 [B, Nq, pad8(H*dv)], identical on every rank — and, with
        want_alive, the u8 [B, Nq] flags of the rows that saw a valid key on SOME rank (what the reference's wipe of
        fully masked rows needs; derived from the merged sums, so no extra reduction over the ranks)."""
        from . import ops
        Op, mp, lp = parts
        S, B, H, Nq, dv = Op.shape
        rows = B * H * Nq
        n = rows * (dv + 2)
        alive = torch.empty((B, Nq), dtype=torch.uint8, device=Op.device) if want_alive else None
        if self.exchange == "peer" and self.world > 1:
            buf, hdl = self._peer_buffer(n, Op.device)
            o, m, l = packed_views(buf, rows, dv)
            # this rank's (merged) partial goes straight into its symmetric buffer
            ops.attention_combine(Op, mp, lp, normalised=False, merged_out=(o, m, l))
            hdl.barrier(channel=0)   # every rank's partial is complete and visible
            out = ops.attention_combine(None, None, None, row_keep=row_keep, shape=(self.world, B, H, Nq, dv),
                                        part_ptrs_dev=int(hdl.buffer_ptrs_dev), device=Op.device, row_alive=alive)
            if torch.cuda.is_current_stream_capturing():
                # a captured forward re-uses THIS buffer on every replay (the alternation above happens at capture time
                # only), so the peers must have finished reading it before the next replay overwrites it
                hdl.barrier(channel=1)
            return (out, alive) if want_alive else out
        if S == 1:
            packed = pack_partial(Op[0], mp[0], lp[0])
        else:
            packed = torch.empty(n, dtype=torch.float32, device=Op.device)
            o, m, l = packed_views(packed, rows, dv)
            ops.attention_combine(Op, mp, lp, normalised=False, merged_out=(o, m, l))
        allp = self.gather_packed(packed)                      # [world, n]
        out = ops.attention_combine(allp, allp.view(-1)[rows * dv:], allp.view(-1)[rows * dv + rows:],
                                    row_keep=row_keep, part_stride_O=n, part_stride_ml=n,
                                    shape=(self.world, B, H, Nq, dv), row_alive=alive)
        return (out, alive) if want_alive else out


def shard_encoder_keys(encoder, group=None, local_splits: int = 0, exchange: str = "nccl"):
    """Mark `encoder` (a PerceiverEncoder of this package) as key-sharded: its forward then expects this rank's slice
    of the input array (see `shard_keys`) and returns the full latents on every rank."""
    encoder.key_shard = KeyShard(group, local_splits, exchange)
    return encoder
