"""Device-timed DP stage on the reference's real read/window pairs (tests/golden/real4000_inputs.npz replicated) and on C2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "real4000_inputs.npz"))
rep = 66
n0 = len(g["q_lens"])

def unpack(words, total_bytes):
    out = np.empty(total_bytes, np.uint8)
    w = words.astype(np.uint32)
    for k in range(8):
        out[k::8] = ((w >> (28 - 4 * k)) & 15).astype(np.uint8)
    lut = np.full(16, ord('N'), np.uint8); lut[1] = ord('A'); lut[3] = ord('C'); lut[7] = ord('G'); lut[4] = ord('T')
    return lut[out]
q_ascii = unpack(g["packed_q"], g["packed_q"].size * 8); t_ascii = unpack(g["packed_t"], g["packed_t"].size * 8)
qo = g["q_offs"].astype(np.uint64); to = g["t_offs"].astype(np.uint64)
Q = np.tile(q_ascii, rep); T = np.tile(t_ascii, rep)
q_offs = np.concatenate([qo + r * q_ascii.size for r in range(rep)]).astype(np.uint32)
t_offs = np.concatenate([to + r * t_ascii.size for r in range(rep)]).astype(np.uint32)
b = synth.PairBatch(Q, T, q_offs, t_offs, np.tile(g["q_lens"], rep).astype(np.uint32), np.tile(g["t_lens"], rep).astype(np.uint32))
c2 = synth.config2(262144, seed=2)
for name, batch in (("real x66", b), ("C2", c2)):
    mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
    for algo, nm in ((capi.SEMI_GLOBAL, "semi"), (capi.LOCAL, "local")):
        with Aligner(algo=algo, start=capi.WITH_START, max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1) as al:
            best = 1e9
            for _ in range(4):
                al.align(batch); best = min(best, al.last_ms(0)[2])
            print(f"{name:9s} {nm:6s} WITH_START dp={best:7.3f} ms GCUPS={batch.cells/best/1e6:8.1f}", flush=True)
