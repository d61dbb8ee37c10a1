python -m pytest tests/test_nvbio_gpu.py -m gpu -q --maxfail=5 -x -k "min_score or cpp_front or one_call" 2>&1 | tail -8
python tools/profile_new_r02b.py 65536 > gpurun_out/prof_r02b_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'k_banded_tb_s16|k_nvb_strip|k_nvb_s16' -c 40 -o gpurun_out/prof_r02b python tools/profile_new_r02b.py 65536 > gpurun_out/prof_r02b_ncu.log 2>&1
echo rc=$?
ncu -i gpurun_out/prof_r02b.ncu-rep --page raw --csv > gpurun_out/prof_r02b_raw.csv 2> gpurun_out/prof_r02b_raw.err
sz=$(stat -c %s gpurun_out/prof_r02b.ncu-rep); if [ "$sz" -gt 45000000 ]; then rm gpurun_out/prof_r02b.ncu-rep; echo "report removed ($sz bytes), csv kept"; fi
cat gpurun_out/prof_r02b_plain.log
