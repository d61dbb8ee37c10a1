"""ncu target: the headline step's kernels at the end of round 2 (C2 shape, 65536 pairs): LOCAL WITH_START (fused kernel) and SEMI_GLOBAL
WITH_START (score-only kernel + start-pass kernel START_REV), plus the score-only LOCAL / GLOBAL kernels.
    python tools/profile_c2_r02c.py          (plain run first; then under ncu --set full -k regex:'k_local_s16|k_semiglobal_s16|k_global_s16')"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
b = synth.config2(65536, seed=2)
for name, kw in (("local WITH_START", dict(algo=capi.LOCAL, start=capi.WITH_START)), ("semi-global WITH_START", dict(algo=capi.SEMI_GLOBAL, start=capi.WITH_START)),
                 ("local", dict(algo=capi.LOCAL)), ("global", dict(algo=capi.GLOBAL)),
                 ("semi-global head NONE tail BOTH WITH_START", dict(algo=capi.SEMI_GLOBAL, head=capi.NONE, tail=capi.BOTH, start=capi.WITH_START))):
    with Aligner(max_query_len=157, max_target_len=607, max_n_alns=b.n, n_streams=1, **kw) as al:
        al.align(b)
        print(f"{name:45s} kernel={al.kernel_name(0)} dp_ms={al.last_ms(0)[2]:.3f} launches={al.launches(0)}", flush=True)
