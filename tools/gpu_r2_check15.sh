timeout 900 python -m pytest tests/test_wave_gpu.py -m gpu -q --maxfail=10 2>&1 | tail -30
