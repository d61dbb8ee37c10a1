"""Quick device-side timing of one batch per mode (development aid; bench.py is the judged benchmark)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
batch = synth.config2(n, seed=2)
cells = batch.cells
lib = capi.load()
for algo, name in ((capi.LOCAL, "local"), (capi.SEMI_GLOBAL, "semi_global"), (capi.GLOBAL, "global")):
    for start in (capi.WITHOUT_START, capi.WITH_START):
        if algo == capi.GLOBAL and start == capi.WITH_START:
            continue
        for policy in [int(x) for x in os.environ.get('QUICK_POLICIES', '0,1').split(',')]:
            lib.gasalx_set_kernel_policy(policy)
            with Aligner(algo=algo, start=start, max_query_len=157, max_target_len=607, max_n_alns=n, n_streams=1) as al:
                best = None
                for it in range(3):
                    al.align(batch)
                    ms = al.last_ms(0)
                    best = ms if best is None or ms[2] < best[2] else best
                print(f"{name:12s} start={start} policy={policy} kernel={al.kernel_name(0):6s} n={n} total={best[0]:8.2f} ms h2d+pack={best[1]:7.2f} ms dp={best[2]:8.2f} ms  "
                      f"GCUPS(dp)={cells / best[2] / 1e6:8.1f}  GCUPS(e2e)={cells / best[0] / 1e6:8.1f}", flush=True)
lib.gasalx_set_kernel_policy(0)
