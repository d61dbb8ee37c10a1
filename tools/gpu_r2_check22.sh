python -m pytest tests/test_nvbio_gpu.py -m gpu -q --maxfail=5 -x -k "quad_strip or ragged or golden or range_proof or tie_order" 2>&1 | tail -8
python -m pytest tests -m gpu -q --maxfail=5 -x -k "banded" 2>&1 | tail -4
python tools/bench_nvb.py 262144 150 600 3 0 2>&1 | tail -3 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    for k, v in d['results'].items():
        if 'int32' not in k: print(f\"{k:40s} {v['kernel']:12s} {v['ms']:8.3f} ms {v['gcups']:8.1f} GCUPS\")
"
python tools/bench_c4.py 2>&1 | head -2
