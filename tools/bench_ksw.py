"""Device-side timing of the KSW seed-extension kernel (algo = KSW): 150 bp reads extending along 200 bp windows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
b = synth.make_extension_pairs(n, 150, 200)          # read = a mutated copy of the window's first 150 bases: the extension runs the whole read
for seed_score in (20, 100):
    with Aligner(max_query_len=157, max_target_len=207, max_n_alns=b.n, n_streams=1, algo=capi.KSW) as al:
        seeds = np.full(b.n, seed_score, dtype=np.uint32)
        best = None
        for _ in range(3):
            r = al.align(b, seeds=seeds)
            ms = al.last_ms(0); best = ms if best is None or ms[2] < best[2] else best
        print(f"KSW seed={seed_score} kernel={al.kernel_name(0)} n={n} dp={best[2]:.3f} ms pairs/s={n / best[2] * 1e3:.3g} mean score={r['aln_score'].mean():.1f} mean q_end={r["query_batch_end"].mean():.1f} gcups(full rectangle)={b.cells / best[2] / 1e6:.0f}", flush=True)
