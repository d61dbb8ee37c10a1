python -m pytest tests/test_nvbio_gpu.py -m gpu -q --maxfail=5 -x 2>&1 | tail -15
python tools/bench_nvb.py 262144 150 600 3 0 2>&1 | tail -3 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    for k, v in d['results'].items(): print(f\"{k:40s} {v['kernel']:12s} {v['ms']:8.3f} ms {v['gcups']:8.1f} GCUPS\")
"
