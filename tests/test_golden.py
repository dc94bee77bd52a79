"""Oracle vs the committed golden fixtures (outputs of the compiled reference, tests/golden/
make_golden.py), plus ABI surface checks.  Runs without a GPU and without /root/reference."""
import ctypes as C
import json
import os
import re

import numpy as np

import helpers as H

GOLD = os.path.join(H.ROOT, "tests", "golden")


def _codes(s):
    return np.frombuffer(s.encode().translate(bytes.maketrans(b"ACGT", b"\0\1\2\3")), dtype=np.uint8).copy()


def test_oracle_reproduces_ksw_known_answers():
    kat = json.load(open(os.path.join(GOLD, "ksw_kat.json")))
    assert len(kat) >= 40
    for k in kat:
        cig, score = H.oracle_ksw(_codes(k["q"]), _codes(k["t"]), k["w"])
        assert H.cigar_str(cig) == k["cigar"] and score == k["score"]


def test_oracle_reproduces_reference_sv_signatures():
    """discover_func restated (oracle/svsig_oracle.c) on the golden SvReads == qx2svsig's file."""
    d = os.path.join(GOLD, "svsig_small")
    q, s = H.LesvDb(os.path.join(d, "db"), "query"), H.LesvDb(os.path.join(d, "db"), "subject")
    reads = H.load_sv_reads(os.path.join(d, "svr", "subject_00000000.sv_read"))
    contig = s.codes(0)
    got = []
    for r in reads:
        n, sigs = H.oracle_discover(r, q.codes(r[0]), contig, 40)
        got += [t + (q.name(t[0]),) for t in sigs]
    ins, dele = H.load_sv_signatures(os.path.join(d, "svsig", "subject_00000000.sv_signature"))
    assert sorted(x for x in got if x[12] == 6) == ins
    assert sorted(x for x in got if x[12] == 0) == dele
    assert len(ins) + len(dele) >= 5


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads without a GPU and exports what include/lesvgpu.h declares."""
    from lesv_b200 import capi
    lib = capi.load()
    hdr = open(os.path.join(H.ROOT, "include", "lesvgpu.h")).read()
    names = set(re.findall(r"\b(lesvgpu_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), n
    assert lib.lesvgpu_device_count() >= 0


def test_no_cpu_fallback_without_gpu():
    from lesv_b200 import capi
    if capi.load().lesvgpu_device_count() > 0:
        return
    h = C.c_void_p()
    assert capi.load().lesvgpu_ctx_create(0, 0, C.byref(h)) == capi.E_NOGPU


def test_library_cell_count_equals_oracle_loop():
    """lesvgpu_ksw_cells (closed form when the band reaches the last cell) == the oracle's row-by-row count
    (ksw2_extd2_sse.c:124-138); host arithmetic only, no GPU."""
    from lesv_b200 import capi
    rng = np.random.default_rng(12)
    n_ok = 0
    for it in range(4000):
        ql, tl = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
        w = abs(ql - tl) + int(rng.integers(0, 4000)) if it % 3 else int(rng.integers(1, 3000))
        if it % 7 == 0:
            w = max(ql, tl)
        if it % 11 == 0:
            ql = 1 + it % 5
        w = max(w, 1)
        got = capi.ksw_cells(ql, tl, w)
        exp = H.oracle_cells(ql, tl, w)
        assert got == exp, (ql, tl, w, got, exp)
        n_ok += exp[1]
    assert n_ok > 2000
