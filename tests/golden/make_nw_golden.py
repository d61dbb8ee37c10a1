"""Generates tests/golden/nw_golden.npz from the REFERENCE's own nw_cpu and generateQuery (Non-CDP/NW/main.cu compiled for
the host: `make -C oracle ref_nw_host`).  Run in the build container, where /root/reference exists; the vectors travel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

ref_host = po.NwRefHost()
out = {"sizes": np.array([1, 7, 64, 100, 130, 333], dtype=np.int64)}
for n in out["sizes"]:
    ref, qry = ref_host.generate(int(n), seed=1000 + int(n))
    mat = ref_host.matrix(ref, qry)
    out[f"ref_{n}"] = ref; out[f"qry_{n}"] = qry; out[f"mat_{n}"] = mat
    out[f"sum_{n}"] = np.uint64(po.nw_checksum_of(mat))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nw_golden.npz"), **out)
print("wrote nw_golden.npz", {int(n): int(out[f"mat_{n}"][-1, -1]) for n in out["sizes"]})
