"""Multi-process sharding for one-process-per-GPU drivers (torch.distributed: NCCL on GPUs, gloo on CPU).

Pairs are independent: the path shards with NO data-path collective.  Rank 0 plans a
length-balanced assignment (LPT on DP cells over batches of length-binned pairs -- the same planner the
in-process batcher uses, gasalx_plan_batches), every rank aligns its own pairs, and the small result arrays
are gathered on rank 0 and scattered back to the caller's order.  The gather is the only communication;
it moves 12-20 bytes per pair.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np

from .aligner import plan_batches
from .synth import PairBatch


def shard_indices(q_lens: np.ndarray, t_lens: np.ndarray, world: int, max_batch_pairs: int = 4096, pieces_per_rank: int = 32):
    """Per-rank pair indices, balanced on sum(q*t) by LPT over length-binned batches.  The sharding batches are cut so that every
    rank receives about `pieces_per_rank` of them (LPT's imbalance is bounded by one piece)."""
    total = int((np.asarray(q_lens, dtype=np.int64) * np.asarray(t_lens, dtype=np.int64)).sum())
    piece = max(total // max(world * pieces_per_rank, 1), 1 << 20)
    p = plan_batches(q_lens, t_lens, n_gpus=world, max_batch_pairs=max_batch_pairs, max_batch_cells=piece)
    per_rank = [[] for _ in range(world)]
    st = p["batch_start"]
    for b, dev in enumerate(p["batch_gpu"]):
        per_rank[int(dev)].append(p["order"][st[b]:st[b + 1]])
    return [np.concatenate(x) if x else np.zeros(0, dtype=np.uint32) for x in per_rank]


def take_pairs(batch: PairBatch, idx: np.ndarray) -> PairBatch:
    """Sub-batch with the given pairs, re-laid out contiguously (gasal_host_batch_fill layout)."""
    from .synth import batch_from_sequences
    return batch_from_sequences([batch.query(int(i)) for i in idx], [batch.target(int(i)) for i in idx])


def align_sharded(batch: Optional[PairBatch], align_fn: Callable[[PairBatch], Dict[str, np.ndarray]], keys, dist=None) -> Optional[Dict[str, np.ndarray]]:
    """Rank 0 holds `batch`; returns the results in original order on rank 0 (None elsewhere).
    `align_fn` runs one rank's shard (on its GPU through the C ABI; a CPU stand-in in the gloo test)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        res = align_fn(batch)
        return {k: res[k] for k in keys}
    rank, world = dist.get_rank(), dist.get_world_size()
    shards = None
    if rank == 0:
        idx = shard_indices(batch.q_lens, batch.t_lens, world)
        shards = [(i, take_pairs(batch, i)) for i in idx]
    mine = [None]
    dist.scatter_object_list(mine, shards, src=0)
    my_idx, my_batch = mine[0]
    res = align_fn(my_batch) if my_batch.n else {k: np.zeros(0, dtype=np.int32) for k in keys}
    payload = (my_idx, {k: np.asarray(res[k]) for k in keys})
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return None
    out = {k: np.zeros(batch.n, dtype=np.int32) for k in keys}
    for idx, r in gathered:
        for k in keys:
            out[k][idx] = r[k]
    return out
