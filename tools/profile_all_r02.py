"""ncu target of round 2: ONE launch set of every kernel family of the engine at a representative (bounded) size, in a fixed order.
    python tools/profile_all_r02.py                      (plain run first; then the same command under ncu --set full -k regex:k_)
Sizes: C2 shape 131072 pairs; C4 shape 131072 pairs; long pairs 32 x 8 kb (wave16) and 24 x 16 kb global; nvbio shape 131072 jobs;
NW matrix N = 16000.  Prints one line per stage with the engine's own event time."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_gpu_b200 import capi, synth, nvb
from genomics_gpu_b200.aligner import Aligner

lib = capi.load()
N2 = int(sys.argv[1]) if len(sys.argv) > 1 else 131072


def run(name, batch, policy=0, **kw):
    lib.gasalx_set_kernel_policy(policy)
    mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
    extra = {k: kw.pop(k) for k in ("q_op", "t_op", "seeds", "cigar_text_style") if k in kw}
    with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1, **kw) as al:
        al.align(batch, **extra)
        print(f"{name:42s} kernel={al.kernel_name(0):20s} ms={al.last_ms(0)[2]:9.3f} launches={al.launches(0)}", flush=True)
    lib.gasalx_set_kernel_policy(0)


b2 = synth.config2(N2, seed=2)
lib.gasalx_set_host_packing(0)                                                        # ASCII pages once: k_pack is on the path
run("C2 local WITH_START (ASCII pages: k_pack)", b2, algo=capi.LOCAL, start=capi.WITH_START)
lib.gasalx_set_host_packing(1)
run("C2 semi-global WITH_START", b2, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)
run("C2 local score/end", b2, algo=capi.LOCAL)
run("C2 global", b2, algo=capi.GLOBAL)
b2n = synth.with_n_bases(b2.slice(0, N2 // 2), 0.05, seed=9)
run("C2 local WITH_START, 5 % pairs with N (k_classify, N pass)", b2n, algo=capi.LOCAL, start=capi.WITH_START)
ops = (np.arange(b2.n) & 3).astype(np.uint8)
run("C2 local + reverse/complement ops (k_revcomp)", b2.slice(0, N2 // 2), algo=capi.LOCAL, q_op=ops[:N2 // 2], t_op=ops[:N2 // 2][::-1].copy())
small = b2.slice(0, N2 // 4)
run("C2 local int32 general kernel (policy 1)", small, policy=1, algo=capi.LOCAL, start=capi.WITH_START)
run("C2 semi-global int32 general kernel", small, policy=1, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)
run("C2 global int32 general kernel", small, policy=1, algo=capi.GLOBAL)
run("C2 local second best (int32)", small, algo=capi.LOCAL, second_best=True)
run("KSW extension (k_ksw_generic)", small, algo=capi.KSW, seeds=np.full(small.n, 20, dtype=np.uint32))
b4 = synth.config4(N2, seed=4)
run("C4 banded k=64 score/end (int32)", b4, algo=capi.BANDED, k_band=64)
run("C4 banded k=64 + CIGAR (s16x2) + CIGAR text", b4, algo=capi.BANDED, k_band=64, start=capi.WITH_TB, cigar_text_style=1)
run("C4 local + CIGAR unbanded (s16x2)", b4.slice(0, N2 // 2), algo=capi.LOCAL, start=capi.WITH_TB)
b5 = synth.config5(1500, seed=5)
run("C5 mixed 100 bp - 10 kb local WITH_START", b5, algo=capi.LOCAL, start=capi.WITH_START)
long_l = synth.config3(32, 8192, seed=3)
run("long local WITH_START 32 x 8 kb (wave16)", long_l, algo=capi.LOCAL, start=capi.WITH_START)
run("long local WITH_TB 32 x 8 kb (wave16 + tb)", long_l.slice(0, 16), algo=capi.LOCAL, start=capi.WITH_TB)
run("long semi-global 32 x 8 kb (wave int32)", long_l, algo=capi.SEMI_GLOBAL, start=capi.WITH_START)
g3 = synth.config3(48, 16384, seed=3)
run("C3 global 48 x 16 kb (wave16)", g3, algo=capi.GLOBAL)
run("C3 global 16 x 16 kb (wave int32)", g3.slice(0, 16), algo=capi.GLOBAL)

# nvbio-style adapter
pats, poffs, texts, toffs = synth.nvb_jobs(N2, 150, 150, 600, 600, seed=31, sub=0.02, indel=0.0)
with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
    for t in (nvb.GLOBAL, nvb.SEMI_GLOBAL, nvb.LOCAL):
        st.enact(nvb.aligner(nvb.GOTOH, t, 2, -1, -2, -1)); print("nvb score type", t, st.kernel(), st.last_ms(), flush=True)
    st.enact(nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1), True); print("nvb local + sinks", st.kernel(), st.last_ms(), flush=True)
    nvb.set_policy(1)
    st.enact(nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 2, -1, -2, -1)); print("nvb int32", st.kernel(), st.last_ms(), flush=True)
    nvb.set_policy(0)
    st.traceback(nvb.aligner(nvb.GOTOH, nvb.LOCAL, 2, -1, -2, -1)); st.alignments(False); print("nvb traceback", st.kernel(), st.last_ms(), flush=True)
pats, poffs, texts, toffs = synth.nvb_banded_jobs(N2, 150, 150, 31, seed=32, sub=0.02, indel=0.002, slack=3)
with nvb.Stream(nvb.Batch(pats, poffs, texts, toffs)) as st:
    al = nvb.aligner(nvb.GOTOH, nvb.SEMI_GLOBAL, 2, -1, -2, -1)
    st.enact_banded(al, 31); print("nvb banded31", st.kernel(), st.last_ms(), flush=True)
    st.traceback(al, 31); st.alignments(False); print("nvb banded31 traceback", st.kernel(), st.last_ms(), flush=True)

# NW full matrix
from genomics_gpu_b200 import nw
r, q = synth.nw_benchmark_pair(16000, seed=7)
_, ms, _ = nw.nw_matrix(r, q, want_matrix=False)
print("nw matrix N=16000", ms, flush=True)
