set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_s4_final2.json 2> gpurun_out/bench_s4_final2.err; echo bench rc=$?
python tools/bench_nvb.py > gpurun_out/bench_nvb_s4b.json 2> gpurun_out/bench_nvb_s4b.err; echo nvbbench rc=$?
ncu --set full --import-source on --clock-control none -k regex:'k_nvb_s16|k_nvb_i32w' -c 12 -o gpurun_out/prof_nvb_s4b python tools/profile_nvb.py 262144 2 > gpurun_out/ncu_nvb_s4b.log 2>&1; echo nvb rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
