"""CIGAR text formatted on the device (the reference prints it on the host, test_prog.cpp:382-428) and the
SAM line built from it: the device text must equal the host decoder of the same bytes, and the bytes equal the oracle's."""
import re

import numpy as np
import pytest

from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner, cigar_string, sam_record
from oracle import pyoracle as po

from helpers import K5, assert_same, oracle_align, cigar_slices

pytestmark = pytest.mark.gpu


def run_text(batch, style, stride=64, **kw):
    mq = int(batch.q_lens.max()) + 7; mt = int(batch.t_lens.max()) + 7
    with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=batch.n, n_streams=1, **kw) as al:
        return al.align(batch, cigar_text_style=style, cigar_stride=stride)


def sam_style(ref_text, q_start, q_end, q_len):
    """reference letters -> SAM: X counts fold into M, clipped query ends become S"""
    ops = [(int(c), o) for c, o in re.findall(r"(\d+)([MXDI])", ref_text)]
    merged = []
    for c, o in ops:
        o = "M" if o == "X" else o
        if merged and merged[-1][1] == o:
            merged[-1][0] += c
        else:
            merged.append([c, o])
    s = (f"{q_start}S" if q_start > 0 else "") + "".join(f"{c}{o}" for c, o in merged)
    tail = q_len - 1 - q_end
    return s + (f"{tail}S" if tail > 0 else "")


@pytest.mark.parametrize("policy", [0, 1])
def test_device_text_matches_host_decoder_and_oracle(oracle, policy):
    b = synth.random_ragged_pairs(500, 30, 220, 30, 330, seed=31, n_frac=0.0)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_TB)
    capi.load().gasalx_set_kernel_policy(policy)
    try:
        got = run_text(b, 0, stride=256, algo=capi.LOCAL, start=capi.WITH_TB)
    finally:
        capi.load().gasalx_set_kernel_policy(0)
    assert_same(got, want, K5 + ["n_cigar_ops"], "local tb")
    assert cigar_slices(got, b) == cigar_slices(want, b)
    for i in range(b.n):
        o, n = int(b.q_offs[i]), int(want["n_cigar_ops"][i])
        assert got["cigar_text"][i] == cigar_string(want["cigar"][o:o + n], n), i
        assert int(got["cigar_text_lens"][i]) == len(got["cigar_text"][i])


def test_sam_style_clips_and_sam_line(oracle):
    b = synth.make_read_window_pairs(300, 100, 260, sub=0.03, indel=0.01, seed=32)
    want = oracle_align(oracle, b, algo=po.LOCAL, start=po.WITH_TB)
    got = run_text(b, 1, stride=128, algo=capi.LOCAL, start=capi.WITH_TB)
    for i in range(b.n):
        o, n = int(b.q_offs[i]), int(want["n_cigar_ops"][i])
        ref_text = cigar_string(want["cigar"][o:o + n], n)
        exp = sam_style(ref_text, int(want["query_batch_start"][i]), int(want["query_batch_end"][i]), int(b.q_lens[i]))
        assert got["cigar_text"][i] == exp, (i, got["cigar_text"][i], exp)
        # a SAM CIGAR consumes exactly the query
        qcons = sum(int(c) for c, op in re.findall(r"(\d+)([MIDS])", exp) if op in "MIS")
        assert qcons == int(b.q_lens[i])
    i = 5
    line = sam_record(f"read{i}", 0, "win", int(got["target_batch_start"][i]) + 1, 60, got["cigar_text"][i], b.query(i), None, int(got["aln_score"][i]))
    f = line.split("\t")
    assert f[0] == "read5" and f[3] == str(int(want["target_batch_start"][i]) + 1) and f[5] == got["cigar_text"][i]
    assert f[9] == b.query(i).decode() and f[11] == f"AS:i:{int(want['aln_score'][i])}"


def test_truncation_is_flagged():
    b = synth.make_read_window_pairs(64, 150, 300, sub=0.05, indel=0.03, seed=33)
    got = run_text(b, 0, stride=8, algo=capi.LOCAL, start=capi.WITH_TB)
    full = run_text(b, 0, stride=512, algo=capi.LOCAL, start=capi.WITH_TB)
    for i in range(b.n):
        t = full["cigar_text"][i]
        if len(t) > 7:
            assert int(got["cigar_text_lens"][i]) == (0x80000000 | len(t))
            assert got["cigar_text"][i] == t[:7]
        else:
            assert got["cigar_text"][i] == t


def test_no_traceback_results_is_an_error():
    b = synth.make_read_window_pairs(32, 100, 200, sub=0.02, indel=0.0, seed=34)
    with pytest.raises(RuntimeError):
        run_text(b, 0, algo=capi.LOCAL, start=capi.WITH_START)


@pytest.mark.parametrize("kw", [dict(algo=capi.GLOBAL, start=capi.WITH_TB), dict(algo=capi.BANDED, k_band=64, start=capi.WITH_TB)])
@pytest.mark.parametrize("style", [0, 1])
def test_text_for_global_and_banded_traceback(kw, style):
    """the device formatter against the host decoder of the SAME result bytes (GLOBAL has no clipped ends: no soft clips)"""
    b = synth.make_read_window_pairs(200, 101, 131, sub=0.03, indel=0.01, seed=35)
    got = run_text(b, style, stride=160, **kw)
    for i in range(b.n):
        o, n = int(b.q_offs[i]), int(got["n_cigar_ops"][i])
        ref_text = cigar_string(got["cigar"][o:o + n], n)
        if style == 0:
            exp = ref_text
        elif kw["algo"] == capi.GLOBAL:
            exp = sam_style(ref_text, 0, int(b.q_lens[i]) - 1, int(b.q_lens[i]))
        else:
            exp = sam_style(ref_text, int(got["query_batch_start"][i]), int(got["query_batch_end"][i]), int(b.q_lens[i]))
        assert got["cigar_text"][i] == exp, (i, got["cigar_text"][i], exp)
