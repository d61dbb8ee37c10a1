timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "wave or mixed or fullsize or parity_gpu or batcher" 2>&1 | tail -6
python - <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
lib = capi.load()
rng = np.random.default_rng(5)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
qs, ts = [], []
for _ in range(1200):
    t, q = synth.nw_style_pair(rng, int(rng.integers(2000, 9000)))
    qs.append(q); ts.append(bytes(rng.choice(acgt, size=int(rng.integers(0, 1500)))) + t + bytes(rng.choice(acgt, size=int(rng.integers(0, 1500)))))
b = synth.batch_from_sequences(qs, ts)
mq = int(b.q_lens.max()) + 7; mt = int(b.t_lens.max()) + 7
for start, nm in ((capi.WITHOUT_START, "score/end"), (capi.WITH_START, "with start")):
    for policy in (3, 0):
        lib.gasalx_set_kernel_policy(policy)
        with Aligner(algo=capi.SEMI_GLOBAL, start=start, max_query_len=mq, max_target_len=mt, max_n_alns=b.n, n_streams=1) as al:
            best = 1e9
            for _ in range(3):
                res = al.align(b); best = min(best, al.last_ms(0)[2])
            print(f"long semi-global {nm:10s} policy={policy} kernel={al.kernel_name(0):7s} dp={best:8.2f} ms GCUPS={b.cells/best/1e6:8.1f} checksum={int(res['aln_score'].astype(np.int64).sum())}", flush=True)
lib.gasalx_set_kernel_policy(0)
PY
