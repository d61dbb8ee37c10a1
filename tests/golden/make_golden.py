"""Generate the golden vectors under tests/golden/ from the REFERENCE's own kernels.

Run in the build container (needs /root/reference and `make -C oracle ref_host`):
    python tests/golden/make_golden.py
The outputs come from oracle/_ref/libgasal_ref_host_mq1024.so, i.e. the unmodified kernel headers of
Non-CDP/GASAL2/src/kernels/ compiled for the host (oracle/ref_shim/).  Inputs are regenerated from
seeds by genomics_gpu_b200.synth, so only the reference's outputs are stored.
  golden_synth.npz : every mode on a seeded ragged batch (inputs stored too: tiny)
  golden_real20k.npz : local WITH_START + semi-global WITH_START on the 20,000 real read/window pairs the
                       reference ships (test_prog/query_batch.fasta.gz, target_batch.fasta.gz); outputs only.
"""
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from genomics_gpu_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FIXTURE = "/root/reference/Non-CDP/GASAL2/test_prog"
K3 = ["aln_score", "query_batch_end", "target_batch_end"]
K5 = K3 + ["query_batch_start", "target_batch_start"]
K2 = ["aln_score2", "query_batch_end2", "target_batch_end2"]


def synth_batch():
    return synth.random_ragged_pairs(256, 9, 200, 9, 300, seed=20261018, n_frac=0.01)


def modes():
    """(name, kwargs, keys) for every mode of the path"""
    out = [("local", dict(algo=po.LOCAL), K3), ("local_start", dict(algo=po.LOCAL, start=po.WITH_START), K5),
           ("local_second", dict(algo=po.LOCAL, second_best=True), K3 + K2),
           ("local_tb", dict(algo=po.LOCAL, start=po.WITH_TB), K5 + ["n_cigar_ops", "cigar"]),
           ("local_s2353", dict(algo=po.LOCAL, start=po.WITH_START, scoring=(2, 3, 5, 3)), K5),
           ("global", dict(algo=po.GLOBAL), ["aln_score"]), ("global_tb", dict(algo=po.GLOBAL, start=po.WITH_TB), ["aln_score", "n_cigar_ops", "cigar"]),
           ("ksw", dict(algo=po.KSW), K3)]
    for h in range(4):
        for t in range(4):
            out.append((f"semi_h{h}_t{t}_start", dict(algo=po.SEMI_GLOBAL, head=h, tail=t, start=po.WITH_START), K5))
    out.append(("semi_second", dict(algo=po.SEMI_GLOBAL, second_best=True), K3 + K2))
    for kb in (8, 24, 64, 4096):
        out.append((f"banded_k{kb}", dict(algo=po.BANDED, k_band=kb), K3))
    return out


def read_fasta_gz(path):
    seqs, cur = [], []
    with gzip.open(path, "rt") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line[0] in "><+/":
                if cur:
                    seqs.append("".join(cur).encode()); cur = []
            else:
                cur.append(line)
    if cur:
        seqs.append("".join(cur).encode())
    return seqs


def real_batch():
    q = read_fasta_gz(os.path.join(REF_FIXTURE, "query_batch.fasta.gz"))
    t = read_fasta_gz(os.path.join(REF_FIXTURE, "target_batch.fasta.gz"))
    return synth.batch_from_sequences(q, t)


def main():
    ref = po.RefHost(1024)
    orc = po.Oracle()
    b = synth_batch()
    pq, pt = ref.pack(b.q_bytes, b.t_bytes, b.n)
    store = {"q_bytes": b.q_bytes, "t_bytes": b.t_bytes, "q_offs": b.q_offs, "t_offs": b.t_offs, "q_lens": b.q_lens, "t_lens": b.t_lens,
             "packed_q": pq, "packed_t": pt, "seed_scores": np.random.default_rng(1).integers(1, 60, size=b.n).astype(np.uint32)}
    for name, kw, keys in modes():
        if kw["algo"] == po.KSW:
            kw = dict(kw, seed_scores=store["seed_scores"])
        r = ref.align(pq, pt, b.q_lens, b.t_lens, b.q_offs, b.t_offs, **kw)
        for k in keys:
            store[f"{name}/{k}"] = r[k]
    np.savez_compressed(os.path.join(HERE, "golden_synth.npz"), **store)
    print("golden_synth.npz:", len(store), "arrays")

    rb = real_batch()
    rpq, rpt = orc.pack(rb.q_bytes), orc.pack(rb.t_bytes)
    rstore = {"n": np.array([rb.n]), "q_lens": rb.q_lens, "t_lens": rb.t_lens}
    for name, kw in (("local_start", dict(algo=po.LOCAL, start=po.WITH_START)), ("semi_start", dict(algo=po.SEMI_GLOBAL, start=po.WITH_START)),
                     ("global", dict(algo=po.GLOBAL)), ("banded_k64", dict(algo=po.BANDED, k_band=64))):
        r = ref.align(rpq, rpt, rb.q_lens, rb.t_lens, rb.q_offs, rb.t_offs, **kw)
        for k in (K5 if "start" in name else (["aln_score"] if name == "global" else K3)):
            rstore[f"{name}/{k}"] = r[k].astype(np.int16) if np.abs(r[k]).max() < 32000 else r[k]
    np.savez_compressed(os.path.join(HERE, "golden_real20k.npz"), **rstore)
    print("golden_real20k.npz:", rb.n, "pairs")

    # the first 4000 real pairs as 4-bit packed INPUTS, so the GPU box (which has no /root/reference) can run them too
    sub = rb.slice(0, 4000)
    np.savez_compressed(os.path.join(HERE, "real4000_inputs.npz"), packed_q=orc.pack(sub.q_bytes), packed_t=orc.pack(sub.t_bytes),
                        q_offs=sub.q_offs, t_offs=sub.t_offs, q_lens=sub.q_lens, t_lens=sub.t_lens)
    print("real4000_inputs.npz written")


if __name__ == "__main__":
    main()
