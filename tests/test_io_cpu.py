"""CPU tests of the caller-side I/O helpers: native FASTA pair reader, host-side 4-bit packing,
CIGAR string formatting.  No GPU, no /root/reference at run time."""
import gzip

import numpy as np
import pytest

from genomics_gpu_b200 import synth
from genomics_gpu_b200.aligner import cigar_string, pack_host, read_fasta_pairs, sam_record


def write_fasta(path, seqs, marks=None, width=None, gz=False, crlf=False):
    nl = "\r\n" if crlf else "\n"
    out = []
    for i, s in enumerate(seqs):
        out.append((marks[i] if marks else ">") + f"seq{i} some description")
        s = s.decode()
        out.extend([s[j:j + width] for j in range(0, len(s), width)] if width else [s])
    data = (nl.join(out) + nl).encode()
    with (gzip.open if gz else open)(path, "wb") as f:
        f.write(data)


@pytest.mark.parametrize("gz,width,crlf", [(False, None, False), (True, 60, False), (False, 7, True)])
def test_fasta_pairs_layout_matches_host_batch_fill(tmp_path, gz, width, crlf):
    b = synth.random_ragged_pairs(200, 1, 300, 1, 400, seed=3, n_frac=0.02)
    qs = [b.query(i) for i in range(b.n)]; ts = [b.target(i) for i in range(b.n)]
    marks_q = [">", "<", "/", "+"] * 50; marks_t = [">"] * 199 + ["+"]
    qp, tp = str(tmp_path / ("q.fa" + (".gz" if gz else ""))), str(tmp_path / ("t.fa" + (".gz" if gz else "")))
    write_fasta(qp, qs, marks_q, width, gz, crlf); write_fasta(tp, ts, marks_t, width, gz, crlf)
    got, q_op, t_op, qn, tn = read_fasta_pairs(qp, tp)
    assert got.n == b.n
    assert np.array_equal(got.q_bytes, b.q_bytes) and np.array_equal(got.t_bytes, b.t_bytes)          # same padding with 'N' to x8
    assert np.array_equal(got.q_offs.astype(np.uint64), b.q_offs.astype(np.uint64)) and np.array_equal(got.t_offs.astype(np.uint64), b.t_offs.astype(np.uint64))
    assert np.array_equal(got.q_lens, b.q_lens) and np.array_equal(got.t_lens, b.t_lens)
    assert list(q_op) == [0, 1, 2, 3] * 50 and list(t_op) == [0] * 199 + [3]                           # test_prog.cpp:80-89
    assert qn[5] == "seq5 some description" and tn[-1] == "seq199 some description"


def test_fasta_pairs_errors(tmp_path):
    write_fasta(str(tmp_path / "a.fa"), [b"ACGT", b"AC"]); write_fasta(str(tmp_path / "b.fa"), [b"ACGT"])
    with pytest.raises(RuntimeError):
        read_fasta_pairs(str(tmp_path / "a.fa"), str(tmp_path / "b.fa"))       # different record counts (test_prog.cpp:131)
    with pytest.raises(RuntimeError):
        read_fasta_pairs(str(tmp_path / "a.fa"), str(tmp_path / "missing.fa"))


def test_pack_host_matches_device_layout(oracle):
    assert pack_host(np.frombuffer(b"ACGTNacg", dtype=np.uint8))[0] == 0x1374E137                      # SURVEY 8(a9) known answer
    b = synth.random_ragged_pairs(100, 1, 200, 1, 200, seed=4, n_frac=0.05)
    assert np.array_equal(pack_host(b.q_bytes), oracle.pack(b.q_bytes))


def test_cigar_string():
    enc = lambda ops: np.array([(c << 2) | "MXDI".index(o) for c, o in reversed(ops)], dtype=np.uint8)   # stored in reverse order
    assert cigar_string(enc([(12, "M"), (1, "X"), (3, "D"), (20, "M")]), 4) == "12M1X3D20M"
    assert cigar_string(enc([(63, "M"), (63, "M"), (10, "M"), (2, "I"), (5, "M")]), 5) == "136M2I5M"   # runs capped at 63 are merged back
    assert cigar_string(enc([(7, "M")]), 1) == "7M"
    assert cigar_string(np.zeros(0, dtype=np.uint8), 0) == ""


def test_sam_record_fields():
    line = sam_record("r1", 16, "chr2", 1234, 37, "3S97M", b"ACGT" * 25, b"I" * 100, 97, 40)
    f = line.split("\t")
    assert f[:9] == ["r1", "16", "chr2", "1234", "37", "3S97M", "*", "0", "0"]
    assert f[9] == "ACGT" * 25 and f[10] == "I" * 100 and f[11:] == ["AS:i:97", "XS:i:40"]
    assert sam_record("", 4, "", 0, 0, "", None, None, 0).split("\t") == ["*", "4", "*", "0", "0", "*", "*", "0", "0", "*", "*", "AS:i:0"]


def test_fastq_reads_against_fasta_windows(tmp_path):
    """queries as FASTQ (gzip), targets as FASTA: the format is sniffed per file; qualities are ignored"""
    b = synth.random_ragged_pairs(50, 10, 150, 10, 300, seed=8)
    with gzip.open(tmp_path / "q.fq.gz", "wb") as f:
        for i in range(b.n):
            s = b.query(i)
            f.write(b"@read%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    write_fasta(str(tmp_path / "t.fa"), [b.target(i) for i in range(b.n)])
    got, q_op, t_op, qn, tn = read_fasta_pairs(str(tmp_path / "q.fq.gz"), str(tmp_path / "t.fa"))
    assert got.n == b.n and qn[7] == "read7"
    assert np.array_equal(got.q_bytes, b.q_bytes) and np.array_equal(got.t_bytes, b.t_bytes)
    assert np.array_equal(got.q_lens, b.q_lens) and not q_op.any()
