"""Device-side timing of the long-pair configurations (development aid):
C3 = global 16k x 16k (BASELINE.json configs[2]), C5 = local WITH_START on mixed 100 bp - 10 kb pairs (configs[4])."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner

lib = capi.load()
what = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else (128 if what == "c3" else 20000)
policies = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
t0 = time.time()
if what == "c3":
    batch = synth.config3(n, 16384, seed=3); algo, start = capi.GLOBAL, capi.WITHOUT_START
elif what == "l3":   # the C3 sequences under LOCAL scoring (scores fit 16 bits: exercises the s16x2 wavefront kernel)
    batch = synth.config3(n, 16384, seed=3); algo, start = capi.LOCAL, capi.WITHOUT_START
else:
    batch = synth.config5(n, seed=5); algo, start = capi.LOCAL, capi.WITH_START
print(f"generated {batch.n} pairs, {batch.cells/1e9:.2f} Gcells in {time.time()-t0:.1f}s", flush=True)
mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
for policy in policies:
    lib.gasalx_set_kernel_policy(policy)
    with Aligner(algo=algo, start=start, max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1) as al:
        best = None
        for it in range(3):
            res = al.align(batch)
            ms = al.last_ms(0)
            best = ms if best is None or ms[2] < best[2] else best
        print(f"{what} policy={policy} kernel={al.kernel_name(0)} n={batch.n} dp={best[2]:.2f} ms GCUPS(dp)={batch.cells/best[2]/1e6:.1f} checksum={int(res['aln_score'].astype(np.int64).sum())}", flush=True)
lib.gasalx_set_kernel_policy(0)
