set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_nvb.py > gpurun_out/bench_nvb_s4c.json 2> gpurun_out/bench_nvb_s4c.err; echo nvbbench rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py > gpurun_out/bench_s4_final3.json 2> gpurun_out/bench_s4_final3.err; echo bench rc=$?
