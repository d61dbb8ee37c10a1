"""Bench of the nvbio-style batched scoring adapter on the sw-benchmark shape (NvB/sw-benchmark/sw-benchmark.cu:560-640: 256 k reads per
batch, Gotoh 2/-1/-2/-1 global / semi-global / local + edit-distance semi-global, GCUPS = sum(M * N) / t).  Prints one JSON object.
    python tools/bench_nvb.py [n_reads] [read_len] [text_len] [reps] [cpu_jobs]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from genomics_gpu_b200 import nvb, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 150
    tl = int(sys.argv[3]) if len(sys.argv) > 3 else 600
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    cpu_jobs = int(sys.argv[5]) if len(sys.argv) > 5 else 4000
    pats, poffs, texts, toffs = synth.nvb_jobs(n, m, m, tl, tl, seed=31, sub=0.02, indel=0.0)
    batch = nvb.Batch(pats, poffs, texts, toffs)
    out = {"workload": f"{n} reads of {m} bp against {tl} bp text windows (codes 0..3), Gotoh 2/-1/-2/-1 + edit distance, scores only", "results": {}}
    with nvb.Stream(batch) as st:
        cells = st.cells()
        cases = [("gotoh_global", nvb.aligner(nvb.GOTOH, nvb.GLOBAL, 2, -1, -2, -1)), ("gotoh_semi_global", nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 2, -1, -2, -1)),
                 ("gotoh_local", nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1)), ("ed_semi_global", nvb.aligner(nvb.EDIT_DISTANCE, nvb.SEMI_GLOBAL))]
        for policy, tag in ((0, "auto"), (2, "band_blocked"), (1, "int32")):
            nvb.set_policy(policy)
            for name, al in cases + [(c[0] + "_sinks", c[1]) for c in cases[1:3]]:
                sinks = name.endswith("_sinks")
                best = 1e30
                for _ in range(reps + 1):
                    st.enact(al, sinks); best = min(best, st.last_ms())
                sc, _, _ = st.results()
                out["results"][f"{name}_{tag}"] = {"kernel": st.kernel(), "ms": best, "gcups": cells / best / 1e6, "score_sum": int(sc.astype(np.int64).sum())}
        nvb.set_policy(0)
    if cpu_jobs:
        from oracle import pyoracle as po
        orc = po.NvbioOracle()
        k = min(cpu_jobs, n)
        t0 = time.perf_counter()
        orc.batch(po.NVB_GOTOH, po.NVB_LOCAL, (2, -1, -2, -1), pats[:poffs[k]], poffs[:k + 1], texts[:toffs[k]], toffs[:k + 1])
        dt = time.perf_counter() - t0
        out["cpu_restatement_gcups"] = {"jobs": k, "gcups": float(np.sum(np.diff(poffs[:k + 1]).astype(np.int64) * np.diff(toffs[:k + 1]).astype(np.int64))) / dt / 1e9,
                                        "threads": __import__("os").cpu_count()}
    out["launches"] = nvb.kernel_launches()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
