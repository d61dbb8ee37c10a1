# one gpurun call: parity of the nvbio adapter and its bench
set -x
python -m pytest tests/test_nvbio_gpu.py -x -q 2>&1 | tail -15
python tools/bench_nvb.py > gpurun_out/bench_nvb.json 2> gpurun_out/bench_nvb.err; echo bench rc=$?; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_nvb.json'))
for k,v in d['results'].items(): print(k, v['kernel'], round(v['ms'],3), round(v['gcups'],1))
print(d.get('cpu_restatement_gcups'))
P
