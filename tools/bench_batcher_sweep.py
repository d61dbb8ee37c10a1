"""Experiment: whole-job wall clock of gasalx_align_all on 1 M pairs 150 x 600 (local WITH_START) against streams / host threads / batch size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import align_all
job = synth.config2(1_000_000, seed=2)
print("cores", os.cpu_count(), flush=True)
out = None
for ns, nt, mb in ((4, 4, 131072), (4, 4, 65536), (8, 8, 65536), (8, 8, 32768), (12, 12, 32768), (16, 16, 32768), (16, 8, 32768), (8, 8, 131072), (16, 16, 65536), (24, 20, 32768)):
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        out = align_all(job, algo=capi.LOCAL, start=capi.WITH_START, n_gpus=1, gpu_ids=[0], streams_per_gpu=ns, host_threads_per_gpu=nt, max_batch_pairs=mb, out=out)
        best = min(best, time.perf_counter() - t0)
    print(f"streams={ns:3d} threads={nt:3d} batch={mb:7d}: {best*1e3:7.2f} ms  {job.cells/best/1e9:8.1f} GCUPS  gpu_ms={out['_gpu_ms'][0]:.1f}", flush=True)
