"""Drop-in proof: the reference's own example program (test_prog/test_prog.cpp), compiled UNCHANGED against our
include/ headers and linked with libgasal_b200.so (oracle/Makefile `dropin`), must print the reference's results."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "test_prog_b200")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/test_prog_b200 not built")]

CODE2CHAR = np.zeros(16, dtype=np.uint8)
for ch in b"ACGTN":
    CODE2CHAR[ch & 15] = ch


def unpack(words, off, length):
    w = words[off // 8:(off + length + 7) // 8]
    nibs = np.stack([(w >> (28 - 4 * k)) & 15 for k in range(8)], axis=1).reshape(-1)[:length]
    return CODE2CHAR[nibs].tobytes().decode()


def write_fasta(path, words, offs, lens, n):
    with open(path, "w") as f:
        for i in range(n):
            f.write(f">s{i}\n{unpack(words, int(offs[i]), int(lens[i]))}\n")


def parse(stdout):
    rows = []
    for line in stdout.splitlines():
        if not line.startswith("query_name="):
            continue
        d = dict(kv.split("=", 1) for kv in line.split("\t"))
        rows.append(d)
    return rows


@pytest.mark.parametrize("mode,args,keys", [
    ("local_start", ["-y", "local", "-s"], {"score": "aln_score", "query_batch_start": "query_batch_start", "target_batch_start": "target_batch_start",
                                             "query_batch_end": "query_batch_end", "target_batch_end": "target_batch_end"}),
    ("semi_start", ["-y", "semi_global", "-s"], {"score": "aln_score", "target_batch_start": "target_batch_start", "target_batch_end": "target_batch_end"}),
    ("global", ["-y", "global"], {"score": "aln_score"}),
])
def test_reference_test_prog_runs_on_our_engine(tmp_path, mode, args, keys):
    inp = np.load(os.path.join(HERE, "golden", "real4000_inputs.npz"))
    gold = np.load(os.path.join(HERE, "golden", "golden_real20k.npz"))
    n = 1200
    qf, tf = str(tmp_path / "q.fasta"), str(tmp_path / "t.fasta")
    write_fasta(qf, inp["packed_q"], inp["q_offs"], inp["q_lens"], n)
    write_fasta(tf, inp["packed_t"], inp["t_offs"], inp["t_lens"], n)
    p = subprocess.run([BIN, "-p", "-n", "1"] + args + [qf, tf], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    rows = parse(p.stdout)
    assert len(rows) == n
    for out_key, gold_key in keys.items():
        got = np.array([int(r[out_key]) for r in rows])
        want = gold[f"{mode}/{gold_key}"][:n].astype(np.int64)
        assert np.array_equal(got, want), f"{mode}: {out_key}"
