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


def emul_ksw(q, t, w, nt, eager=False):
    lib = emul()
    lib.lesv_emul_set_eager(1 if eager else 0)
    cig = C.POINTER(C.c_uint32)()
    n = lib.lesv_emul_ksw_extd2(len(q), _ptr(q), len(t), _ptr(t), w, nt, C.byref(cig))
    if n < 0:
        return n
    out = [cig[i] for i in range(n)]
    _libc.free(cig)
    return out


def emul_ksw_skew(q, t, w, nt, k, seed):
    lib = emul()
    cig = C.POINTER(C.c_uint32)()
    n = lib.lesv_emul_ksw_extd2_skew(len(q), _ptr(q), len(t), _ptr(t), w, nt, k, C.c_uint(seed), C.byref(cig))
    if n < 0:
        return n
    out = [cig[i] for i in range(n)]
    _libc.free(cig)
    return out


def emul_ksw_skew_group(q, t, w, nt, group, k, seed):
    lib = emul()
    cig = C.POINTER(C.c_uint32)()
    n = lib.lesv_emul_ksw_extd2_skew_group(len(q), _ptr(q), len(t), _ptr(t), w, nt, group, k, C.c_uint(seed), C.byref(cig))
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


# ---- lesv packed database -------------------------------------------------
class LesvDb:
    """numpy view of <db_dir>/<title>.{pac,seq_info,hdr,volume_info.bin} (corelib/seqdb.c)."""

    def __init__(self, db_dir, title):
        base = os.path.join(db_dir, title + ".")
        vi = np.fromfile(base + "volume_info.bin", dtype=np.uint8)[:64]
        self.num_seqs = int(np.frombuffer(vi[4:8].tobytes(), dtype=np.int32)[0])
        f = np.frombuffer(vi[8:64].tobytes(), dtype=np.uint64)
        self.seq_offset_to = int(f[2])
        si = np.fromfile(base + "seq_info", dtype=np.uint64).reshape(-1, 6)[:self.num_seqs]
        self.seq_offset = si[:, 0].astype(np.int64)
        self.seq_size = si[:, 1].astype(np.int64)
        self.hdr_offset = si[:, 2].astype(np.int64)
        self.hdr = open(base + "hdr", "rb").read()
        self.pac = np.fromfile(base + "pac", dtype=np.uint8)
        self.read_offset = (self.seq_offset + 1).astype(np.uint64)   # raw_read_offset_array

    def name(self, i):
        o = int(self.hdr_offset[i])
        return self.hdr[o:self.hdr.index(b"\0", o)].decode()

    def codes(self, i):
        """forward residues of sequence i as uint8 codes (_get_pac, hbn_aux.h:126)."""
        off, n = int(self.seq_offset[i]), int(self.seq_size[i])
        l = np.arange(off, off + n, dtype=np.int64)
        return ((self.pac[l >> 2] >> ((~l & 3) << 1)) & 3).astype(np.uint8)


def load_sv_reads(path):
    out = []
    with open(path) as f:
        for line in f:
            p = line.rstrip("\n").split("\t")
            v = [int(x) for x in p[:9]]
            # file order: qid qdir qoff qend qsize sid soff send dist
            out.append((v[0], v[5], v[1], v[2], v[3], v[4], v[6], v[7], v[8]))
    return out  # (query_id, subject_id, qdir, qoff, qend, qsize, soff, send, dist)


def load_sv_signatures(path):
    """-> (ins multiset, del multiset) as sorted lists of 13-int tuples + read name."""
    ins, dele = [], []
    with open(path) as f:
        for line in f:
            p = line.rstrip("\n").split("\t")
            rec = tuple(int(x) for x in p[:13]) + (p[13],)
            (ins if rec[12] == 6 else dele).append(rec)
    return sorted(ins), sorted(dele)


class OracleSvRead(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("qid", "sid", "qdir", "qoff", "qend", "qsize", "soff", "send", "dist")]


class OracleSvSig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("qid", "qdir", "qfrom", "qto", "fqfrom", "fqto", "qsize", "sfrom", "sto",
                                       "fsfrom", "fsto", "ssize", "type")]


def oracle_discover(svr, read_fwd, contig, min_indel):
    """lesv_oracle_discover_one -> (n or -1, list of 13-int tuples)."""
    lib, _ = oracle()
    s = OracleSvRead(*[int(x) for x in (svr[0], svr[1], svr[2], svr[3], svr[4], svr[5], svr[6], svr[7], svr[8])])
    out = (OracleSvSig * 64)()
    n = lib.lesv_oracle_discover_one(C.byref(s), _ptr(read_fwd), len(read_fwd), _ptr(contig), len(contig), min_indel, out, 64)
    return n, [tuple(getattr(out[i], f) for f, _ in OracleSvSig._fields_) for i in range(max(n, 0))]


def trace_emul():
    """Host build of trace_core.cuh (tests/emul/trace_emul.cpp): the sequential restatement of edlib_nw etc."""
    if "t" not in _cache:
        so = os.path.join(ROOT, "tests", "emul", "libtrace_emul.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul"), "libtrace_emul.so"], check=True)
        _cache["t"] = C.CDLL(so)
    return _cache["t"]
