python tools/profile_c2_r02c.py > gpurun_out/prof_r02c_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'k_local_s16|k_semiglobal_s16|k_global_s16' -c 12 -o gpurun_out/prof_r02c python tools/profile_c2_r02c.py > gpurun_out/prof_r02c_ncu.log 2>&1
echo rc=$?
ncu -i gpurun_out/prof_r02c.ncu-rep --page raw --csv > gpurun_out/prof_r02c_raw.csv 2> gpurun_out/prof_r02c_raw.err
sz=$(stat -c %s gpurun_out/prof_r02c.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm gpurun_out/prof_r02c.ncu-rep; fi
cat gpurun_out/prof_r02c_plain.log
