QUICK_POLICIES=0 python tools/quick_bench.py 2>&1 | cut -c1-150
python -m pytest tests -m gpu -q --maxfail=8 -k "parity_gpu or parity_holes or mixed" 2>&1 | tail -3 | cut -c1-300
