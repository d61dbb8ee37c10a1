python -m pytest tests -m gpu -q --maxfail=8 -k "batcher or parity_gpu or parity_holes or mixed or fullsize or banded" 2>&1 | tail -12
python tools/bench_c4.py 2>&1 | head -2
