"""Opcode histogram of the hottest (largest innermost) loop of each DP kernel in libgasal_b200.so -> profiles/sass_*_r01.txt.
Evidence for DESIGN.md's per-cell instruction counts (cuobjdump -sass; no GPU needed)."""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "genomics_gpu_b200", "lib", "libgasal_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", sass)[1:]
want = {"k_local_s16ILi0": "local_s16_score_end", "k_semiglobal_s16ILi0": "semiglobal_s16_score_end", "k_global_s16": "global_s16",
        "k_waveILi1ELi2": "wave_global_sr2", "k_waveILi3ELi2": "wave_local_sr2"}
out = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    key = next((v for k, v in want.items() if k in name), None)
    if not key:
        continue
    ins = []
    for ln in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for addr, txt in ins:
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    lo, hi = max(inner, key=lambda l: l[1] - l[0])
    hist = collections.Counter()
    for addr, txt in ins:
        if lo <= addr <= hi:
            t = txt.split()
            hist[t[1] if t[0].startswith("@") else t[0]] += 1
    out.append(f"== {key} ({name}): hottest innermost loop 0x{lo:x}..0x{hi:x}, {sum(hist.values())} instructions")
    for op, c in hist.most_common():
        out.append(f"   {c:5d}  {op}")
open(os.path.join(ROOT, "profiles", "sass_loops_r01.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
