set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_s4_final.json 2> gpurun_out/bench_s4_final.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_s4_final.json 2> gpurun_out/bench_ref_s4_final.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s4.csv python bench.py --steps 2 --warmup 1 --extras 0 --refgpu-pairs 0 --cpu-pairs 1000 > gpurun_out/ncu_launch_s4.log 2>&1; echo launchlist rc=$?
ncu --set full --import-source on --clock-control none -k regex:'k_pack|k_local_s16|k_semiglobal_s16|k_global_s16' -c 10 -o gpurun_out/prof_fast_s4 python tools/profile_target.py 262144 1 > gpurun_out/ncu_fast_s4.log 2>&1; echo fast rc=$?
ncu --set full --import-source on --clock-control none -k regex:'k_nvb_s16|k_nvb_i32w' -c 6 -o gpurun_out/prof_nvb_s4 python tools/profile_nvb.py 262144 2 > gpurun_out/ncu_nvb_s4.log 2>&1; echo nvb rc=$?
python tools/bench_nvb.py > gpurun_out/bench_nvb_s4.json 2> gpurun_out/bench_nvb_s4.err; echo nvbbench rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
