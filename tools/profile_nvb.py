"""ncu target for the nvbio-style scoring kernels: one enact of each Gotoh type on the sw-benchmark shape (tools/bench_nvb.py).
    ncu --set full --import-source on --clock-control none -k regex:k_nvb_s16 -c 3 -o gpurun_out/prof_nvb python tools/profile_nvb.py [n] [policy]"""
import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from genomics_gpu_b200 import nvb, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
policy = int(sys.argv[2]) if len(sys.argv) > 2 else 0
pats, poffs, texts, toffs = synth.nvb_jobs(n, 150, 150, 600, 600, seed=31, sub=0.02, indel=0.0)
with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
    for pol in ((0, 1) if policy == 2 else (policy,)):          # policy 2: packed path first, then the int32 path
        nvb.set_policy(pol)
        for type_ in (nvb.GLOBAL, nvb.SEMI_GLOBAL, nvb.LOCAL):
            st.enact(nvb.aligner(nvb.GOTOH, type_, 2, -1, -2, -1))
            sc, _, _ = st.results()
            print(type_, st.kernel(), st.last_ms(), int(sc.sum()))
nvb.set_policy(0)
