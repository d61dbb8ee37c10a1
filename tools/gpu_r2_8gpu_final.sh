set -x
nvidia-smi --query-gpu=index,name --format=csv | head -10
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --extras 0 --refgpu-pairs 0 --cpu-pairs 20000 > gpurun_out/r2_bench_8gpu_final3.json 2> gpurun_out/r2_bench_8gpu_final3.err; echo rc=$?; tail -3 gpurun_out/r2_bench_8gpu_final3.err
