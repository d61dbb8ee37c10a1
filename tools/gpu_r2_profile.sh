set -x
python tools/profile_all_r02.py 65536 > gpurun_out/prof_r02_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'^k_|k_nvb|k_nw' -c 120 -o gpurun_out/prof_all_r02 python tools/profile_all_r02.py 65536 > gpurun_out/prof_r02_ncu.log 2>&1
echo rc=$?
ncu -i gpurun_out/prof_all_r02.ncu-rep --page raw --csv > gpurun_out/prof_all_r02_raw.csv 2> gpurun_out/prof_all_r02_raw.err
ls -la gpurun_out/
sz=$(stat -c %s gpurun_out/prof_all_r02.ncu-rep); if [ "$sz" -gt 45000000 ]; then rm gpurun_out/prof_all_r02.ncu-rep; echo "report removed ($sz bytes), csv kept"; fi
tail -5 gpurun_out/prof_r02_ncu.log
