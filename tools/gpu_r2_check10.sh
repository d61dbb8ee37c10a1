set -x
timeout 900 python -m pytest tests/test_nvbio_tb_gpu.py tests/test_nvbio_gpu.py -m gpu -q --maxfail=10 2>&1 | tail -40 > gpurun_out/r2_tests10.log; tail -30 gpurun_out/r2_tests10.log
