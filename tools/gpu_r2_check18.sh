python -m pytest tests -m gpu -q --maxfail=20 -k "parity_holes or parity_gpu or fullsize or host_packing or mixed or ref_gpu or dropin" 2>&1 | tail -6
python tools/bench_real.py 2>&1 | tail -4
