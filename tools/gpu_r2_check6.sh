set -x
python -m pytest tests -m gpu -q --maxfail=20 -k "parity_holes or parity_gpu" 2>&1 | tail -30 > gpurun_out/r2_tests6.log; tail -8 gpurun_out/r2_tests6.log
export QUICK_POLICIES=0
for v in main noprune; do
  if [ $v = main ]; then unset GASAL_B200_LIB; else export GASAL_B200_LIB=$PWD/build/variants/$v/libgasal_b200.so; fi
  echo "== variant $v"; python tools/quick_bench.py 262144 2>&1 | grep "local"
done > gpurun_out/r2_quick6.log 2>&1
cat gpurun_out/r2_quick6.log
unset GASAL_B200_LIB
python - <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
# real reads (ragged) and unrelated pairs (no alignment: the pruning never triggers)
g = np.load('tests/golden/real4000_inputs.npz')
for name, b in (("unrelated 150x600", synth.random_ragged_pairs(100000, 150, 150, 600, 600, seed=3, related=0.0, repeat_frac=0.0)),):
    for start in (capi.WITHOUT_START, capi.WITH_START):
        with Aligner(algo=capi.LOCAL, start=start, max_query_len=157, max_target_len=607, max_n_alns=b.n, n_streams=1) as al:
            best = 1e9
            for _ in range(3):
                al.align(b); best = min(best, al.last_ms(0)[2])
            print(f"{name} start={start} dp={best:7.3f} ms GCUPS={b.cells/best/1e6:8.1f}", flush=True)
PY
