"""Time the reference's own CUDA build (sm_100) on the same GPU, on a sample of the C2 workload."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomics_gpu_b200 import synth
from oracle import pyoracle as po

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
batch = synth.config2(n, seed=2)
out = {}
for name, algo, mq in (("local_with_start", po.LOCAL, 160), ("semi_global_with_start", po.SEMI_GLOBAL, 608), ("local_score_only", po.LOCAL, 160)):
    start = po.WITHOUT_START if name.endswith("only") else po.WITH_START
    for bs in (5000, n // 2):
        r = po.RefGpu(mq).run(batch, algo=algo, start=start, batch_size=bs, reps=3)
        out[f"{name}/batch{bs}"] = r["_timing"]
        print(name, bs, r["_timing"], flush=True)
json.dump(out, open(os.path.join("gpurun_out", "ref_gpu_baseline.json"), "w"), indent=1)
