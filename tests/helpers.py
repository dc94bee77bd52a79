"""Shared test plumbing: ctypes views of the oracle, the compiled reference and the emulator."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from lesv_b200.synth import mutate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liblesv_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libhbn_ref.so")
EMUL_SO = os.path.join(ROOT, "tests", "emul", "libksw_emul.so")

_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


class OraclePrm(C.Structure):
    _fields_ = [(n, C.c_int8) for n in ("match", "mismatch", "ambi", "q", "e", "q2", "e2")]


class Ez(C.Structure):  # ksw_extz_t, reference src/algo/ksw2.h:24-33
    _fields_ = [("max_zd", C.c_uint32), ("max_q", C.c_int), ("max_t", C.c_int), ("mqe", C.c_int),
                ("mqe_t", C.c_int), ("mte", C.c_int), ("mte_q", C.c_int), ("score", C.c_int),
                ("m_cigar", C.c_int), ("n_cigar", C.c_int), ("reach_end", C.c_int),
                ("cigar", C.POINTER(C.c_uint32))]


_cache = {}


def oracle():
    if "o" not in _cache:
        lib = C.CDLL(ORACLE_SO)
        lib.lesv_oracle_ksw_cells.restype = C.c_int64
        lib.lesv_oracle_calc_ident_perc.restype = C.c_double
        lib.lesv_oracle_calc_effective_ident_perc.restype = C.c_double
        prm = OraclePrm()
        lib.lesv_oracle_ksw_default_params(C.byref(prm))
        _cache["o"] = (lib, prm)
    return _cache["o"]


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref():
    if "r" not in _cache:
        lib = C.CDLL(REF_SO)
        mat = (C.c_int8 * 25)()
        lib.ksw_gen_simple_mat(5, mat, C.c_int8(2), C.c_int8(5), C.c_int8(1))
        _cache["r"] = (lib, mat)
    return _cache["r"]


def emul():
    if "e" not in _cache:
        _cache["e"] = C.CDLL(EMUL_SO)
    return _cache["e"]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def oracle_ksw(q, t, w):
    """(cigar list, score) from the oracle restatement."""
    lib, prm = oracle()
    cig = C.POINTER(C.c_uint32)()
    sc = C.c_int()
    n = lib.lesv_oracle_ksw_extd2(len(q), _ptr(q), len(t), _ptr(t), C.byref(prm), w, C.byref(cig), C.byref(sc))
    out = [cig[i] for i in range(n)]
    if n:
        _libc.free(cig)
    return out, sc.value


def oracle_cells(ql, tl, w):
    lib, _ = oracle()
    ok = C.c_int()
    c = lib.lesv_oracle_ksw_cells(ql, tl, w, C.byref(ok))
    return int(c), bool(ok.value)


def ref_ksw(q, t, w):
    """(cigar list, score) from the compiled reference's ksw_extd2_sse (flag 0, zdrop -1)."""
    lib, mat = ref()
    ez = Ez()
    lib.ksw_extd2_sse(None, len(q), _ptr(q), len(t), _ptr(t), C.c_int8(5), mat, C.c_int8(5), C.c_int8(4),
                      C.c_int8(56), C.c_int8(1), w, -1, -1, 0, C.byref(ez))
    out = [ez.cigar[i] for i in range(ez.n_cigar)]
    if ez.n_cigar:
        _libc.free(ez.cigar)
    return out, ez.score


def emul_ksw(q, t, w, nt):
    lib = emul()
    cig = C.POINTER(C.c_uint32)()
    n = lib.lesv_emul_ksw_extd2(len(q), _ptr(q), len(t), _ptr(t), w, nt, C.byref(cig))
    if n < 0:
        return n
    out = [cig[i] for i in range(n)]
    _libc.free(cig)
    return out


def _tb(lib, fn, q, t, w, nt=None):
    cig = C.POINTER(C.c_uint32)()
    tb = C.POINTER(C.c_uint8)()
    rb = C.c_int()
    if nt is None:
        _, prm = oracle()
        n = getattr(lib, fn)(len(q), _ptr(q), len(t), _ptr(t), C.byref(prm), w, C.byref(cig), C.byref(tb), C.byref(rb))
    else:
        n = getattr(lib, fn)(len(q), _ptr(q), len(t), _ptr(t), w, nt, C.byref(cig), C.byref(tb), C.byref(rb))
    nrow = len(q) + len(t) - 1
    arr = np.ctypeslib.as_array(tb, shape=(nrow * rb.value,)).reshape(nrow, rb.value).copy()
    _libc.free(tb)
    if n > 0:
        _libc.free(cig)
    return arr


def oracle_tb(q, t, w):
    return _tb(oracle()[0], "lesv_oracle_ksw_extd2_tb", q, t, w)


def emul_tb(q, t, w, nt):
    return _tb(emul(), "lesv_emul_ksw_extd2_tb", q, t, w, nt)


def min_threads(ql, tl, w):
    ww = w if w >= 0 else max(ql, tl)
    return (min(ww + 1, tl) + 15) // 16 + 3


def random_pair(rng, it, max_len=700, small=80):
    """Random (query, target, band) in the shapes that stress the band edges."""
    mode = it % 5
    if mode == 0:
        q = rng.integers(0, 4, int(rng.integers(1, small)), dtype=np.uint8)
        t = rng.integers(0, 4, int(rng.integers(1, small)), dtype=np.uint8)
    elif mode == 4:  # low complexity: long exact runs, many score ties
        k = int(rng.integers(1, 4))
        q = (rng.integers(0, 2, int(rng.integers(1, max_len)), dtype=np.uint8) * k).astype(np.uint8)
        t = mutate(q, float(rng.uniform(0, 0.2)), rng)
        if len(t) == 0:
            t = q.copy()
    else:
        q = rng.integers(0, 4, int(rng.integers(1, max_len)), dtype=np.uint8)
        t = mutate(q, float(rng.uniform(0, 0.3)), rng)
        if rng.random() < 0.3:
            k = int(rng.integers(0, len(t) + 1))
            t = np.concatenate([t[:k], rng.integers(0, 4, int(rng.integers(1, 100)), dtype=np.uint8), t[k:]])
        if len(t) == 0:
            t = q.copy()
        if rng.random() < 0.5:
            q, t = t, q
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    big = max(len(q), len(t))
    w = [-1, big, int(rng.integers(0, 40)), int(rng.integers(0, big + 20)),
         abs(len(q) - len(t)) + int(rng.integers(0, 30))][int(rng.integers(0, 5))]
    return q, t, w


def cigar_str(c):
    return "".join(f"{x >> 4}{'MID'[x & 15]}" for x in c)
