"""Worker for test_multiprocess_shard_and_gather_gloo (launched with torch.distributed.run, 2 ranks, gloo)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from genomics_gpu_b200 import sharding, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

dist.init_process_group("gloo")
rank = dist.get_rank()
orc = po.Oracle()
keys = ["aln_score", "query_batch_end", "target_batch_end", "query_batch_start", "target_batch_start"]


def cpu_stand_in(b):
    return orc.align(orc.pack(b.q_bytes), orc.pack(b.t_bytes), b.q_lens, b.t_lens, b.q_offs, b.t_offs, algo=po.LOCAL, start=po.WITH_START, n_threads=2)


batch = synth.random_ragged_pairs(400, 20, 300, 20, 500, seed=5) if rank == 0 else None
got = sharding.align_sharded(batch, cpu_stand_in, keys, dist)
if rank == 0:
    want = cpu_stand_in(batch)
    for k in keys:
        assert np.array_equal(got[k], want[k]), k
    print("GLOO_SHARD_GATHER_OK", flush=True)
dist.barrier()
dist.destroy_process_group()
