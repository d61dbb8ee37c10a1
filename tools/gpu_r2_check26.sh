python -m pytest tests -m gpu -q --maxfail=8 -k "all_head_tail or mixed or semiglobal" 2>&1 | tail -30 | cut -c1-330
