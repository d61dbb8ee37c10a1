"""One process per GPU (torchrun) for the nvbio-style scoring adapter: every rank scores its own 262144 reads (weak scaling, no data-path
collective -- jobs are independent); the timed region is bracketed by barrier + synchronize, device time is the max over ranks.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29551 tools/bench_nvb_multi.py [n_per_gpu] [steps]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomics_gpu_b200 import nvb, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pats, poffs, texts, toffs = synth.nvb_jobs(n, 150, 150, 600, 600, seed=31 + rank, sub=0.02, indel=0.0)
out = {}
with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
    cells = st.cells()
    for name, al in (("gotoh_semi_global", nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 2, -1, -2, -1)), ("gotoh_local", nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1))):
        for _ in range(3):
            st.enact(al)
        st.last_ms()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = 0.0
        for _ in range(steps):
            st.enact(al); ms += st.last_ms()
        torch.cuda.synchronize()
        t = torch.tensor([ms / steps], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = {"ms_per_step_max_over_ranks": float(t.item()), "gcups_all_gpus": world * cells / float(t.item()) / 1e6, "kernel": st.kernel()}
if rank == 0:
    print(json.dumps({"workload": f"{n} reads of 150 bp x 600 bp per GPU, Gotoh 2/-1/-2/-1, scores only", "n_gpus": world, "scaling": "weak", "results": out}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
