set -x
python tools/profile_c4.py > gpurun_out/prof_c4_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:'k_banded_tb_s16' -s 2 -c 2 -o gpurun_out/prof_c4_r02 python tools/profile_c4.py > gpurun_out/prof_c4_ncu.log 2>&1
echo rc=$?; cat gpurun_out/prof_c4_plain.log; ls -la gpurun_out/*.ncu-rep
