python -m pytest tests -m gpu -q --maxfail=5 -k "ksw" 2>&1 | tail -12 | cut -c1-300
python tools/bench_ksw.py 2>&1 | tail -3
