"""Device-side timing of config 4: banded (k_band=64) local alignment + CIGAR on 250 bp read / ~298 bp window pairs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
batch = synth.config4(n, seed=4)
mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
for name, kw in (("banded k=64 score/end", dict(algo=capi.BANDED, k_band=64)),
                 ("banded k=64 + CIGAR", dict(algo=capi.BANDED, k_band=64, start=capi.WITH_TB)),
                 ("local + CIGAR (unbanded)", dict(algo=capi.LOCAL, start=capi.WITH_TB)),
                 ("local with start (fast path)", dict(algo=capi.LOCAL, start=capi.WITH_START))):
    with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1, **kw) as al:
        best = None
        for it in range(3):
            res = al.align(batch)
            ms = al.last_ms(0)
            best = ms if best is None or ms[2] < best[2] else best
        print(f"{name:30s} kernel={al.kernel_name(0):6s} n={n} total={best[0]:8.2f} ms dp+tb={best[2]:8.2f} ms GCUPS(dp)={batch.cells/best[2]/1e6:8.1f} pairs/s={n/best[2]*1e3:.3g}", flush=True)
