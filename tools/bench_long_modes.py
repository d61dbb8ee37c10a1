"""Device-timed DP stage of the three wavefront modes on ONE long-read batch (reads of 2-9 kb in flanked windows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
lib = capi.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
b = synth.long_read_windows(n, 2000, 9000, 1500, seed=5)
mq = int(b.q_lens.max()) + 7; mt = int(b.t_lens.max()) + 7
print(f"{b.n} pairs, {b.cells/1e9:.1f} Gcells", flush=True)
for algo, nm in ((capi.LOCAL, "local"), (capi.SEMI_GLOBAL, "semi"), (capi.GLOBAL, "global")):
    for start in ((capi.WITHOUT_START, capi.WITH_START) if algo != capi.GLOBAL else (capi.WITHOUT_START,)):
        for policy in (0, 3):
            lib.gasalx_set_kernel_policy(policy)
            with Aligner(algo=algo, start=start, max_query_len=mq, max_target_len=mt, max_n_alns=b.n, n_streams=1) as al:
                best = 1e9
                for _ in range(3):
                    al.align(b); best = min(best, al.last_ms(0)[2])
                print(f"{nm:6s} start={start} policy={policy} kernel={al.kernel_name(0):7s} dp={best:8.2f} ms GCUPS={b.cells/best/1e6:8.1f}", flush=True)
lib.gasalx_set_kernel_policy(0)
