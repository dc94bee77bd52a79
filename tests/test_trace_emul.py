"""trace_core.cuh (per-candidate alignment code of the mapper kernels, built for the host by tests/emul)
against the compiled reference: edlib_nw (src/algo/edlib_wrapper.c:32), the MEM extraction of
InitHitFindData_SetupMapAlignInfoFromHit (src/algo/init_hit_finder.c:469) and hbn_traceback
(src/algo/hbn_traceback.c:313, overhang step off).  Alignment strings must be identical."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from lesv_b200.synth import mutate, revcomp_codes

pytestmark = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built")
_e = {}


def emul():
    if "l" not in _e:
        _e["l"] = C.CDLL(os.path.join(H.ROOT, "tests", "emul", "libtrace_emul.so"))
        _e["l"].lesv_emul_traceback.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                                C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    return _e["l"]


class KString(C.Structure):
    _fields_ = [("l", C.c_size_t), ("m", C.c_size_t), ("s", C.c_void_p)]


class KVec(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p)]


class ChainSeed(C.Structure):
    _fields_ = [("length", C.c_int), ("qoff", C.c_int64), ("soff", C.c_int64), ("sdir", C.c_int), ("hash", C.c_uint64)]


def ops_to_strings(ops, q, t):
    qa, ta, i, j = [], [], 0, 0
    for op in ops:
        if op == 0:
            qa.append("ACGT"[q[i]]); ta.append("ACGT"[t[j]]); i += 1; j += 1
        elif op == 1:
            qa.append("ACGT"[q[i]]); ta.append("-"); i += 1
        else:
            qa.append("-"); ta.append("ACGT"[t[j]]); j += 1
    return "".join(qa), "".join(ta), i, j


def ref_nw(q, t):
    lib, _ = H.ref()
    lib.EdlibAlignDataNew.restype = C.c_void_p
    lib.edlib_nw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    if "ed" not in _e:
        _e["ed"] = lib.EdlibAlignDataNew()
    qa, ta = KString(), KString()
    lib.edlib_nw(_e["ed"], H._ptr(q), len(q), H._ptr(t), len(t), -1, C.byref(qa), C.byref(ta))
    a = C.string_at(qa.s, qa.l).decode() if qa.l else ""
    b = C.string_at(ta.s, ta.l).decode() if ta.l else ""
    H._libc.free(qa.s); H._libc.free(ta.s)
    return a, b


def test_nw_path_equals_edlib():
    rng = np.random.default_rng(5)
    lib = emul()
    for it in range(600):
        mode = it % 4
        n = int(rng.integers(1, 400 if mode else 1500))
        q = rng.integers(0, 4 if mode != 3 else 2, n, dtype=np.uint8)       # mode 3: two-letter, many ties
        if mode == 1:
            t = rng.integers(0, 4, int(rng.integers(1, 300)), dtype=np.uint8)
        else:
            t = np.ascontiguousarray(mutate(q, float(rng.uniform(0.02, 0.35)), rng))
            if len(t) == 0:
                t = q[:1].copy()
        if mode == 2 and n > 60:                                            # a long gap inside
            k = int(rng.integers(10, n - 40)); t = np.ascontiguousarray(np.concatenate([t[:k], t[k + 30:]]))
        ops = np.zeros(len(q) + len(t) + 8, dtype=np.uint8)
        cnt = lib.lesv_emul_nw(H._ptr(q), len(q), H._ptr(t), len(t), H._ptr(ops), len(ops))
        assert cnt > 0
        a, b, i, j = ops_to_strings(ops[:cnt], q, t)
        assert (i, j) == (len(q), len(t))
        ra, rb = ref_nw(q, t)
        assert (a, b) == (ra, rb), (it, len(q), len(t))


def test_nw_path_hirschberg_equals_edlib():
    """Problems past edlib's 1 MB traceback limit go through its Hirschberg recursion."""
    rng = np.random.default_rng(9)
    lib = emul()
    for it in range(14):
        n = int(rng.integers(1900, 6000))
        q = rng.integers(0, 4, n, dtype=np.uint8)
        t = np.ascontiguousarray(mutate(q, float(rng.uniform(0.05, 0.3)), rng))
        if it % 3 == 1:       # unrelated middle part
            k = n // 3
            t = np.ascontiguousarray(np.concatenate([t[:k], rng.integers(0, 4, 2500, dtype=np.uint8), t[k + 2000:]]))
        if it % 3 == 2:       # low complexity: many co-optimal splits
            q = rng.integers(0, 2, n, dtype=np.uint8)
            t = np.ascontiguousarray(mutate(q, 0.2, rng) & 1)
        assert 20 * ((len(q) + 63) // 64) * len(t) + 8 * len(t) >= 1 << 20
        ops = np.zeros(len(q) + len(t) + 8, dtype=np.uint8)
        cnt = lib.lesv_emul_nw(H._ptr(q), len(q), H._ptr(t), len(t), H._ptr(ops), len(ops))
        assert cnt > 0, cnt
        a, b, i, j = ops_to_strings(ops[:cnt], q, t)
        ra, rb = ref_nw(q, t)
        assert (a, b) == (ra, rb), (it, len(q), len(t))


class RefMapper:
    """InitHitFindData + HbnTracebackData of the reference, driven like s_hbn_extend_subject_subseq_hit_list."""

    def __init__(self):
        lib, _ = H.ref()
        self.lib = lib
        lib.InitHitFindDataNew.restype = C.c_void_p
        lib.HbnTracebackDataNew.restype = C.c_void_p
        lib.InitHitFindData_Init.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.InitHitFindData_AddQuery.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.InitHitFindData_FindHits.argtypes = [C.c_void_p, C.c_int]
        lib.InitHitFindData_SetupMapAlignInfoFromHit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.hbn_traceback.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.c_double, C.c_int]
        self.ihf = lib.InitHitFindDataNew(10, 10, 30)
        self.tb = lib.HbnTracebackDataNew()

    def hits(self, fwd, rev, window, strand):
        """[(score, chain seeds [(w, r)], mems [(q, s, len)])] for the top hits of one window."""
        lib = self.lib
        lib.InitHitFindData_Init(self.ihf, 0, None, H._ptr(fwd), H._ptr(rev), len(fwd))
        lib.InitHitFindData_AddQuery(self.ihf, 0, None, H._ptr(window), None, len(window))
        lib.InitHitFindData_FindHits(self.ihf, strand)
        hl = KVec.from_address(self.ihf + 184)   # offsetof(InitHitFindData, hit_list)
        out = []
        from test_chain_emul import InitHit
        arr = C.cast(hl.a, C.POINTER(InitHit))
        for k in range(min(hl.n, 3)):
            h = arr[k]
            cs = C.cast(h.chain_seed_array, C.POINTER(ChainSeed))
            seeds = [(int(cs[x].qoff), int(cs[x].soff)) for x in range(h.chain_seed_count)]
            sl = KVec()
            lib.InitHitFindData_SetupMapAlignInfoFromHit(self.ihf, C.byref(h), C.byref(sl))
            ms = C.cast(sl.a, C.POINTER(ChainSeed))
            mems = [(int(ms[x].qoff), int(ms[x].soff), int(ms[x].length)) for x in range(sl.n)]
            out.append((h.score, seeds, mems, sl))
        return out

    def traceback(self, query, window, sl, over_hang=0):
        r = self.lib.hbn_traceback(self.tb, H._ptr(query), len(query), H._ptr(window), len(window), sl.a, int(sl.n), 100, 0.0,
                                   over_hang)
        f = (C.c_int * 8).from_address(self.tb)
        qas = C.c_void_p.from_address(self.tb + 184).value
        qae = C.c_void_p.from_address(self.tb + 192).value
        sas = C.c_void_p.from_address(self.tb + 200).value
        n = qae - qas
        return r, (f[0], f[1], f[3], f[4]), C.string_at(qas, n).decode() if n > 0 else "", C.string_at(sas, n).decode() if n > 0 else ""


def emul_traceback(query, window, mems, over_hang=0):
    lib = emul()
    m3 = np.ascontiguousarray(np.array(mems, dtype=np.int32).reshape(-1))
    res = np.zeros(8, dtype=np.int32)
    ops = np.zeros(3 * len(query) + 2 * len(window) + 4096, dtype=np.uint8)
    lib.lesv_emul_traceback(H._ptr(query), len(query), H._ptr(window), len(window), H._ptr(m3), len(mems), 100, 0.0,
                            over_hang, H._ptr(res), H._ptr(ops))
    return res, ops[:res[5]]


def emul_mems(query, window, seeds, cap):
    lib = emul()
    w = np.ascontiguousarray([a for a, _ in seeds], dtype=np.int32)
    r = np.ascontiguousarray([b for _, b in seeds], dtype=np.int32)
    out = np.zeros(3 * cap, dtype=np.int32)
    n = lib.lesv_emul_mems(H._ptr(query), H._ptr(window), H._ptr(w), H._ptr(r), len(seeds), 10, H._ptr(out), cap)
    return [tuple(int(x) for x in out[3 * i:3 * i + 3]) for i in range(n)]


@pytest.mark.parametrize("seed,over_hang", [(0, 0), (1, 0), (2, 1), (3, 1)])
def test_mems_and_traceback_equal_reference(seed, over_hang):
    rng = np.random.default_rng(40 + seed)
    R = RefMapper()
    n_ok = n_flag = 0
    for it in range(60):
        L = int(rng.integers(1500, 12000))
        ref = rng.integers(0, 4, L + 6000, dtype=np.uint8)
        a = int(rng.integers(0, 3000))
        sample = ref[a:a + L].copy()
        kind = it % 5
        if kind == 1:     # deletion in the read
            k = L // 2; sample = np.concatenate([sample[:k], sample[k + int(rng.integers(40, 900)):]])
        elif kind == 2:   # insertion in the read
            k = L // 3; sample = np.concatenate([sample[:k], rng.integers(0, 4, int(rng.integers(40, 900)), dtype=np.uint8), sample[k:]])
        elif kind == 3:   # read hangs over the window start
            sample = np.concatenate([rng.integers(0, 4, int(rng.integers(20, 400)), dtype=np.uint8), sample])
        read = np.ascontiguousarray(mutate(sample, float(rng.uniform(0.05, 0.2)), rng))
        strand = int(rng.integers(0, 2))
        fwd = np.ascontiguousarray(revcomp_codes(read)) if strand else read      # stored read; `read` is the aligned strand
        rev = np.ascontiguousarray(revcomp_codes(fwd))
        query = rev if strand else fwd
        lo = max(0, a - int(rng.integers(0, 2500)))
        window = np.ascontiguousarray(ref[lo:min(len(ref), a + L + int(rng.integers(0, 2500)))])
        for score, seeds, mems, sl in R.hits(fwd, rev, window, strand):
            assert emul_mems(query, window, seeds, len(mems) + 8) == mems, (it, "mems")
            if not mems:
                continue
            r, coords, qas, sas = R.traceback(query, window, sl, over_hang)
            res, ops = emul_traceback(query, window, mems, over_hang)
            assert res[0] in (0, 1), (it, res)
            assert int(res[0]) == int(r), (it, res, r)
            if r:
                assert tuple(int(x) for x in res[1:5]) == coords, (it, res, coords)
                qa, sa, _, _ = ops_to_strings(ops, query[res[1]:], window[res[3]:])
                assert (qa, sa) == (qas, sas), (it, len(qa), len(qas))
                n_ok += 1
                n_flag += int(res[7] != 0)
    assert n_ok >= 50
    assert not over_hang or n_flag >= 20       # the overhang step really ran
