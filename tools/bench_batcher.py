"""Whole-job throughput through the library-side batcher (gasalx_align_all): length binning, LPT sharding over the GPUs of ONE
process (one worker thread per GPU), pinned streaming, host-side gather.  BASELINE.json configs[4] (mixed 100 bp - 10 kb, local
WITH_START) and configs[1] (150 x 600) -- wall-clock from host buffers to results in the caller's order."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import align_all

max_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n5 = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
jobs = {"C5_mixed_100bp_10kb": (synth.config5(n5, seed=5), dict(algo=capi.LOCAL, start=capi.WITH_START)),
        "C2_150x600": (synth.config2(1_000_000 if len(sys.argv) <= 3 else int(sys.argv[3]), seed=2), dict(algo=capi.LOCAL, start=capi.WITH_START))}      # scheduling: the library's automatic choices
out = {}
for name, (job, kw) in jobs.items():
    ref = None
    for g in [x for x in (1, 2, 4, 8) if x <= max_gpus]:
        best = None
        res = None
        for rep in range(3):
            t0 = time.perf_counter()
            res = align_all(job, n_gpus=g, gpu_ids=list(range(g)), out=res, **kw)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        chk = int(res["aln_score"].astype(np.int64).sum())
        if ref is None: ref = chk
        assert chk == ref, "results differ between GPU counts"
        out[f"{name}/gpus{g}"] = {"pairs": job.n, "wall_s": best, "gcups_wall": job.cells / best / 1e9, "gpu_ms": [round(float(x), 2) for x in res["_gpu_ms"]], "score_checksum": chk}
        print(name, g, out[f"{name}/gpus{g}"], flush=True)
json.dump(out, open("gpurun_out/bench_batcher.json", "w"), indent=1)
