"""ncu target for the kernels added late in round 2: the score-only BANDED sweep on the s16x2 lanes (k_banded_tb_s16<NMASK, TB = false>)
and the nvbio quad-strip scoring kernel (k_nvb_strip<TYPE, SINK>), next to the band-blocked packed kernel they replace.
    python tools/profile_new_r02b.py [jobs]          (plain run first; then the same command under ncu --set full -k regex:...)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomics_gpu_b200 import capi, synth, nvb
from genomics_gpu_b200.aligner import Aligner

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
b4 = synth.config4(N, seed=4)
mq = int(b4.q_lens.max()) + 7; mt = int(b4.t_lens.max()) + 7
with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=b4.n, n_streams=1, algo=capi.BANDED, k_band=64) as al:
    al.align(b4)
    print(f"C4 banded k=64 score/end kernel={al.kernel_name(0)} ms={al.last_ms(0)[2]:.3f}", flush=True)
pats, poffs, texts, toffs = synth.nvb_jobs(N, 150, 150, 600, 600, seed=31, sub=0.02, indel=0.0)
with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
    for policy in (0, 2):
        nvb.set_policy(policy)
        for t, name in ((nvb.GLOBAL, "global"), (nvb.SEMI_GLOBAL, "semi-global"), (nvb.LOCAL, "local")):
            st.enact(nvb.aligner(nvb.GOTOH, t, 2, -1, -2, -1)); print(f"nvbio Gotoh {name:12s} kernel={st.kernel():12s} ms={st.last_ms():.3f}", flush=True)
    nvb.set_policy(0)
    st.enact(nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 2, -1, -2, -1), True); print(f"nvbio Gotoh semi-global + sinks kernel={st.kernel()} ms={st.last_ms():.3f}", flush=True)
