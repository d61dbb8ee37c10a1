"""Experiment: the s16x2 kernels at 96 bp x 600 bp (66.5 KB of shared memory per CTA: three CTAs per SM) with and without extra
shared memory that caps the residency at two / one CTA per SM (GASALX_FAST_SMEM_PAD_KB): does a third CTA per SM buy throughput?"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from genomics_gpu_b200 import capi, synth
    from genomics_gpu_b200.aligner import Aligner
    qlen = int(sys.argv[2])
    b = synth.make_read_window_pairs(262144, qlen, 600, sub=0.02, indel=0.003, seed=2)
    for algo, nm, st in ((capi.LOCAL, "local", capi.WITHOUT_START), (capi.LOCAL, "local", capi.WITH_START), (capi.SEMI_GLOBAL, "semi", capi.WITHOUT_START), (capi.GLOBAL, "global", capi.WITHOUT_START)):
        with Aligner(algo=algo, start=st, max_query_len=qlen + 7, max_target_len=607, max_n_alns=b.n, n_streams=1) as al:
            best = 1e9
            for _ in range(4):
                al.align(b); best = min(best, al.last_ms(0)[2])
            print(f"   q={qlen} {nm:6s} start={st} dp={best:7.3f} ms GCUPS={b.cells/best/1e6:8.1f}", flush=True)
else:
    for qlen, pads in ((96, (0, 30, 100)), (64, (0, 10, 30, 60, 150)), (150, (0, 100))):
        for pad in pads:
            env = dict(os.environ, GASALX_FAST_SMEM_PAD_KB=str(pad))
            smem = (((qlen + 7) // 8 * 8) + 8) * 5 * 128 / 1024 + pad
            print(f"== q={qlen} pad={pad} KB -> {smem:.1f} KB per CTA -> {int(227 // (smem + 1))} CTAs per SM", flush=True)
            subprocess.run([sys.executable, __file__, "child", str(qlen)], env=env)
