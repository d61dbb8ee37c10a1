"""Opcode histogram of the largest innermost loop of every kernel whose mangled name contains argv[2] in object / library argv[1]
(cuobjdump -sass; no GPU needed).  Evidence for DESIGN.md's per-cell instruction counts."""
import collections, re, subprocess, sys
sass = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", sass)[1:]:
    name = f.split("\n", 1)[0].strip()
    if sys.argv[2] not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in (re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln) for ln in f.split("\n")) if m]
    loops = []
    for addr, txt in ins:
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    if not inner:
        continue
    lo, hi = max(inner, key=lambda l: l[1] - l[0])
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", txt).split()[0] for addr, txt in ins if lo <= addr <= hi)
    print(f"{name}: loop {lo:#x}-{hi:#x}, {sum(hist.values())} instructions")
    print("   " + ", ".join(f"{k} {v}" for k, v in hist.most_common()))
