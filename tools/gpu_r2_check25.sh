python tools/diag_headtail.py 2>&1 | grep -v "^   pair" | awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' | sort | uniq -c | sort -rn | head -20
python -m pytest tests -m gpu -q --maxfail=6 -k "all_head_tail or mixed or semiglobal or nvbio_gpu" 2>&1 | tail -6
