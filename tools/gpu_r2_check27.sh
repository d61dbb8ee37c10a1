QUICK_POLICIES=0 python tools/quick_bench.py 2>&1 | cut -c1-150
python -m pytest tests -m gpu -q --maxfail=8 -k "all_head_tail or mixed or semiglobal or parity_holes" 2>&1 | tail -3 | cut -c1-300
