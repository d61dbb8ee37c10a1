set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_s3_final.json 2> gpurun_out/bench_s3_final.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_s3_final.json 2> gpurun_out/bench_ref_s3_final.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_s3.csv python bench.py --steps 2 --warmup 1 --extras 0 --refgpu-pairs 0 --cpu-pairs 1000 > gpurun_out/ncu_launch_s3.log 2>&1; echo launchlist rc=$?
ncu --set full --import-source on --clock-control none -k regex:'k_pack|k_local_s16|k_semiglobal_s16|k_global_s16' -c 10 -o gpurun_out/prof_fast_s3 python tools/profile_target.py 262144 1 > gpurun_out/ncu_fast_s3.log 2>&1; echo fast rc=$?
NW_REF=0 ncu --set full --import-source on --clock-control none -k regex:k_nw_matrix_ws -c 1 -o gpurun_out/prof_nw_s3 python tools/bench_nw.py 32000 1 > gpurun_out/ncu_nw_s3.log 2>&1; echo nw rc=$?
ncu --set full --import-source on --clock-control none -k regex:'k_local_s16<2>|k_local_s16<\(int\)2>|k_get_tb|k_banded' -c 4 -o gpurun_out/prof_tb_s3 python tools/bench_c4.py 65536 > gpurun_out/ncu_tb_s3.log 2>&1; echo tb rc=$?
python tools/bench_nw.py 32000 5 > gpurun_out/bench_nw_s3.json 2>&1
tail -c 400 gpurun_out/bench_nw_s3.json
