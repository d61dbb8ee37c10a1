python -m pytest tests -m gpu -q --maxfail=10 -k "banded" 2>&1 | tail -8
python tools/bench_c4.py 2>&1 | tail -5
