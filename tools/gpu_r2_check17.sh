python -m pytest tests -m gpu -q --maxfail=10 -k "batcher or sharding or mixed" 2>&1 | tail -4
python tools/bench_batcher.py 2 20000 2>&1 | tail -6
