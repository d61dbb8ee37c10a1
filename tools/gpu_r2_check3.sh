set -x
timeout 1200 python bench.py --steps 10 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo bench rc=$?; tail -5 gpurun_out/r2_bench3.err
