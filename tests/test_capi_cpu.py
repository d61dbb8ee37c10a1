"""CPU tests of the drop-in boundary: the shared library loads without a GPU and exports every symbol that
include/gasal_b200.h declares (no compute calls here); the ctypes structs mirror the header."""
import ctypes as C
import os
import re
import subprocess

from genomics_gpu_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gasal_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_\s\*]*?\b(gasalx?_[A-Za-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_library_loads_and_exports_every_declared_symbol():
    lib = capi.load()
    decl = declared_functions()
    assert len(decl) >= 30
    missing = [n for n in decl if not hasattr(lib, n)]
    assert not missing, f"declared in gasal_b200.h but not exported: {missing}"
    assert sorted(capi.EXPORTED_SYMBOLS) == decl, "capi.EXPORTED_SYMBOLS is out of sync with include/gasal_b200.h"
    assert b"sm_100a" in lib.gasalx_version()


def test_symbols_have_c_linkage():
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    for n in declared_functions():
        assert n in exported, f"{n} is not an unmangled (extern \"C\") export"
    # the C++ front end for unchanged GASAL2 callers lives in the same library
    assert any("Parameters" in s and "parse" in s for s in exported)


def test_struct_layout_matches_header():
    """compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror"""
    import tempfile
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "gasal_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(gasal_gpu_storage_t), offsetof(gasal_gpu_storage_t, host_res), offsetof(gasal_gpu_storage_t, current_n_alns),
         offsetof(gasal_gpu_storage_t, str), offsetof(gasal_gpu_storage_t, b200_priv), sizeof(gasal_params_t));
  printf("%zu %zu %zu\n", sizeof(gasal_res_t), sizeof(host_batch_t), sizeof(gasal_subst_scores));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", exe, src], check=True)   # plain C, no CUDA headers
        a, b = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")[:2]
    sz, o_res, o_cur, o_str, o_priv, sz_par = map(int, a.split())
    G = capi.GpuStorage
    assert (sz, o_res, o_cur, o_str, o_priv) == (C.sizeof(G), G.host_res.offset, G.current_n_alns.offset, G.str.offset, G.b200_priv.offset)
    assert sz_par == C.sizeof(capi.Params)
    assert list(map(int, b.split())) == [C.sizeof(capi.GasalRes), C.sizeof(capi.HostBatch), C.sizeof(capi.SubstScores)]


def test_params_default_matches_reference_defaults():
    lib = capi.load()
    p = capi.Params()
    lib.gasal_params_default(C.byref(p))     # pure host function
    # reference src/args_parser.cpp:12-29
    assert (p.algo, p.start_pos, p.secondBest, p.isPacked, p.isReverseComplement) == (capi.UNKNOWN, capi.WITHOUT_START, 0, 0, 0)
    assert (p.semiglobal_skipping_head, p.semiglobal_skipping_tail, p.k_band) == (capi.TARGET, capi.TARGET, 0)


def test_non_fatal_error_mode_returns_a_status_without_a_gpu():
    """gasalx_try_init_streams validates before it touches CUDA: an impossible batch returns GASALX_E_ARG instead of exit()"""
    import ctypes as C
    lib = capi.load()
    p = capi.Params(); lib.gasal_params_default(C.byref(p)); p.algo = capi.LOCAL
    v = lib.gasal_init_gpu_storage_v(1)
    rc = lib.gasalx_try_init_streams(C.byref(v), 100_000, 100_000, 1_000_000, C.byref(p))      # 10^11 bytes: beyond 32-bit offsets
    assert rc == 1                                                                             # GASALX_E_ARG
    assert b"32-bit byte offsets" in lib.gasalx_last_error()
    lib.gasal_destroy_gpu_storage_v(C.byref(v))
