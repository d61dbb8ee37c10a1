set -x
python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -40 > gpurun_out/r2_tests4.log; tail -25 gpurun_out/r2_tests4.log
export QUICK_POLICIES=0
python tools/quick_bench.py 262144 > gpurun_out/r2_quick4.log 2>&1; cat gpurun_out/r2_quick4.log
python tools/bench_c4.py 262144 > gpurun_out/r2_c4.log 2>&1; cat gpurun_out/r2_c4.log
python - > gpurun_out/r2_nfrac.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '.')
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
b2 = synth.config2(262144, seed=2)
for f in (0.0, 0.01, 0.05, 0.20):
    b = synth.with_n_bases(b2, f, seed=17) if f else b2
    for algo, nm in ((capi.LOCAL, 'local'), (capi.SEMI_GLOBAL, 'semi')):
        with Aligner(algo=algo, start=capi.WITH_START, max_query_len=157, max_target_len=607, max_n_alns=b.n, n_streams=1) as al:
            best = 1e9
            for _ in range(3):
                al.align(b); best = min(best, al.last_ms(0)[2])
            print(f"N-frac {f:4.2f} {nm:6s} dp={best:7.3f} ms GCUPS={b.cells/best/1e6:8.1f}", flush=True)
PY
cat gpurun_out/r2_nfrac.log
