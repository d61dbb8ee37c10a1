"""Device-side timing of the full-matrix NW kernel at the reference benchmark's default N = 32000 (NW/main.cu:67) against the
HBM write roofline (4 B per cell), plus -- when oracle/_ref/nw_ref_gpu exists -- the reference's own nw_gpu3 on the same GPU."""
import json, os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_gpu_b200 import nw as nwmod, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nwmod.set_variant(int(os.environ.get('NW_VARIANT', '0')))
ref, qry = synth.nw_benchmark_pair(n, seed=7)
best = None
for it in range(reps + 1):
    _, ms, h = nwmod.nw_matrix(ref, qry, want_matrix=False, want_checksum=(it == 0))
    if it and (best is None or ms[1] < best):
        best = ms[1]
try:
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    hbm = 6650.0
cells = float(n) * n
out = {"N": n, "kernel_ms": best, "gcups": cells / best / 1e6, "write_gbs": 4 * cells / best / 1e6, "hbm_peak_gbs": hbm,
       "frac_of_hbm": 4 * cells / best / 1e6 / hbm, "bytes_def": "4 B written per cell (the int32 matrix); inputs are 2N bytes"}
exe = os.path.join(ROOT, "oracle", "_ref", "nw_ref_gpu")
if os.path.exists(exe) and os.environ.get("NW_REF", "1") == "1":
    with tempfile.TemporaryDirectory() as td:      # the reference appends runtimes_*.txt files to its cwd
        p = subprocess.run([exe, "-N", str(n), "-3"], cwd=td, capture_output=True, text=True, timeout=600)
    txt = re.sub(r"\x1b\[[0-9;]*m", "", p.stdout)
    m = re.search(r"GPU kernel time \(version 3\): ([0-9.]+) ms", txt)
    c = re.search(r"CPU time: ([0-9.]+) ms", txt)
    out["reference_nw_gpu3_ms"] = float(m.group(1)) if m else None
    out["reference_nw_cpu_ms"] = float(c.group(1)) if c else None
    out["reference_verified"] = "Verification succeeded" in txt
    if m:
        out["speedup_vs_reference_gpu3"] = float(m.group(1)) / best
print(json.dumps(out))
