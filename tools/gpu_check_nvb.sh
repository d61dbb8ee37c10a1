# one gpurun call: parity of the nvbio adapter, its bench, and compute-sanitizer memcheck passes (new module + a slice of the GASAL path)
set -x
python -m pytest tests/test_nvbio_gpu.py -x -q 2>&1 | tail -15
python tools/bench_nvb.py > gpurun_out/bench_nvb.json 2> gpurun_out/bench_nvb.err; echo bench rc=$?; tail -c 1500 gpurun_out/bench_nvb.json
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_nvbio_gpu.py -x -q -k "golden or errors or symbols" > gpurun_out/sanitizer_nvb.log 2>&1; echo sanitizer_nvb rc=$?; tail -5 gpurun_out/sanitizer_nvb.log
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -k "local_with_start or semiglobal_default_with_start or global_score or edge_shapes" > gpurun_out/sanitizer_gasal.log 2>&1; echo sanitizer_gasal rc=$?; tail -5 gpurun_out/sanitizer_gasal.log
