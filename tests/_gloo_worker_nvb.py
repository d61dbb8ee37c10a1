"""Worker for test_nvbio_oracle.py::test_multiprocess_job_sharding_gloo (torch.distributed.run, 2 ranks, gloo)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genomics_gpu_b200 import nvb, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

dist.init_process_group("gloo")
orc = po.NvbioOracle()
G = (2, -1, -2, -1)


def cpu_stand_in(b):
    return orc.batch(po.NVB_GOTOH, po.NVB_SEMI_GLOBAL, G, b.patterns, b.poffs, b.texts, b.toffs, b.pp, b.pt, b.n_pairs)[0]


for shared in (False, True):
    pats, poffs, texts, toffs = synth.nvb_jobs(2500, 5, 80, 30, 200, seed=8, shared_text=shared)
    batch = nvb.Batch(pats, poffs, texts, toffs)
    got = nvb.score_sharded(batch, cpu_stand_in, dist)
    if dist.get_rank() == 0:
        assert np.array_equal(got, cpu_stand_in(batch)), shared
if dist.get_rank() == 0:
    print("GLOO_NVB_SHARD_OK", flush=True)
dist.barrier()
dist.destroy_process_group()
