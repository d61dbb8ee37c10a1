python -m pytest tests -m gpu -q --maxfail=20 -k "banded or parity_holes or cigar or parity_gpu or mixed or dropin or ref_gpu" 2>&1 | tail -5
python tools/bench_c4.py 262144 2>&1 | tail -4
