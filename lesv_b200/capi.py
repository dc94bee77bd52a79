"""ctypes binding of liblesvgpu.so (include/lesvgpu.h) -- the same C ABI a lesv driver links.

No torch types cross this boundary: numpy arrays are passed as plain pointers.  Importing this
module never falls back to a CPU path: if the shared library is missing, `load()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblesvgpu.so")

OK, E_CUDA, E_ARG, E_NOMEM, E_NOGPU = 0, 1, 2, 3, 4


class KswPair(C.Structure):
    _fields_ = [("query", C.c_void_p), ("qlen", C.c_int), ("target", C.c_void_p), ("tlen", C.c_int),
                ("max_distance", C.c_int)]


class KswResult(C.Structure):
    _fields_ = [("ok", C.c_int), ("n_cigar", C.c_int), ("cigar_off", C.c_size_t),
                ("aln_size", C.c_int), ("n_match", C.c_int)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("pack_ms", C.c_float), ("dp_ms", C.c_float),
                ("scan_ms", C.c_float), ("d2h_ms", C.c_float), ("dp_launches", C.c_int),
                ("kernel_launches", C.c_int), ("h2d_bytes", C.c_size_t), ("d2h_bytes", C.c_size_t),
                ("dp_cells", C.c_uint64)]


class ClassTiming(C.Structure):
    _fields_ = [("ctas_per_alignment", C.c_int), ("threads", C.c_int), ("n_alignments", C.c_int), ("n_ctas", C.c_int),
                ("cells", C.c_uint64), ("done_ms", C.c_float)]


_lib = None


def load() -> C.CDLL:
    """dlopen liblesvgpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    lib.lesvgpu_device_count.restype = C.c_int
    lib.lesvgpu_ctx_create.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.lesvgpu_ctx_destroy.argtypes = [C.c_void_p]
    lib.lesvgpu_last_error.argtypes = [C.c_void_p]
    lib.lesvgpu_last_error.restype = C.c_char_p
    lib.lesvgpu_last_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    lib.lesvgpu_ksw_batch.argtypes = [C.c_void_p, C.POINTER(KswPair), C.c_int, C.POINTER(KswResult),
                                      C.POINTER(C.POINTER(C.c_uint32))]
    lib.lesvgpu_ksw_cells.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.lesvgpu_ksw_cells.restype = C.c_int64
    _lib = lib
    return lib


class Context:
    """One GPU, one feeder thread (lesvgpu_ctx)."""

    def __init__(self, device: int = 0, arena_bytes: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.lesvgpu_ctx_create(device, arena_bytes, C.byref(h))
        if rc != OK:
            raise RuntimeError(f"lesvgpu_ctx_create(device={device}) failed with status {rc} "
                               "(no B200 / sm_100 GPU visible?)")
        self.h = h

    def class_timing(self):
        """Launch classes of the last K-ksw batch: list of dicts (shape, alignments, cells, done_ms)."""
        self.lib.lesvgpu_last_class_timing.argtypes = [C.c_void_p, C.POINTER(ClassTiming), C.c_int]
        arr = (ClassTiming * 32)()
        n = min(32, self.lib.lesvgpu_last_class_timing(self.h, arr, 32))
        return [{f: getattr(arr[i], f) for f, _ in ClassTiming._fields_} for i in range(n)]

    def close(self):
        if self.h:
            self.lib.lesvgpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def error(self) -> str:
        return self.lib.lesvgpu_last_error(self.h).decode()

    def timing(self) -> Timing:
        t = Timing()
        self.lib.lesvgpu_last_timing(self.h, C.byref(t))
        return t

    def ksw_batch(self, queries, targets, max_distances, want_cigars: bool = True):
        """Batch form of nw_ksw2_extd2. queries/targets: lists of uint8 code arrays.

        Returns (results, cigars): results is a ctypes array of KswResult, cigars a list of numpy
        uint32 arrays (BAM-coded, first op first) or None per pair.
        """
        n = len(queries)
        pairs = (KswPair * n)()
        keep = []
        for i in range(n):
            q = np.ascontiguousarray(queries[i], dtype=np.uint8)
            t = np.ascontiguousarray(targets[i], dtype=np.uint8)
            keep.append((q, t))
            pairs[i].query = q.ctypes.data
            pairs[i].qlen = len(q)
            pairs[i].target = t.ctypes.data
            pairs[i].tlen = len(t)
            pairs[i].max_distance = int(max_distances[i])
        res = (KswResult * n)()
        pool = C.POINTER(C.c_uint32)()
        rc = self.lib.lesvgpu_ksw_batch(self.h, pairs, n, res, C.byref(pool))
        if rc != OK:
            raise RuntimeError(f"lesvgpu_ksw_batch failed ({rc}): {self.error()}")
        cigars = None
        if want_cigars:
            cigars = []
            for i in range(n):
                if res[i].ok:
                    a = np.ctypeslib.as_array(pool, shape=(res[i].cigar_off + res[i].n_cigar,))
                    cigars.append(a[res[i].cigar_off:].copy())
                else:
                    cigars.append(None)
        return res, cigars


def ksw_cells(qlen: int, tlen: int, max_distance: int):
    ok = C.c_int()
    c = load().lesvgpu_ksw_cells(qlen, tlen, max_distance, C.byref(ok))
    return int(c), bool(ok.value)


class SvRead(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("query_id", "subject_id", "qdir", "qoff", "qend", "qsize", "soff", "send", "dist")]


class SvSig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("qid", "qdir", "qfrom", "qto", "fqfrom", "fqto", "qsize", "sfrom", "sto",
                                       "fsfrom", "fsto", "ssize", "type")]

    def astuple(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class SvSigRead(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("status", "n_sig", "qb", "qe", "sb", "se", "eff_mat", "eff_len", "band")]


def _bind_svsig(lib):
    lib.lesvgpu_subject_load.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.lesvgpu_svsig_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                        C.POINTER(SvRead), C.c_int, C.c_int, C.c_void_p,
                                        C.POINTER(SvSigRead), C.POINTER(C.POINTER(SvSig)), C.POINTER(C.c_int)]


def subject_load(ctx: "Context", pac: np.ndarray):
    _bind_svsig(ctx.lib)
    pac = np.ascontiguousarray(pac, dtype=np.uint8)
    rc = ctx.lib.lesvgpu_subject_load(ctx.h, pac.ctypes.data, pac.size)
    if rc != OK:
        raise RuntimeError(f"lesvgpu_subject_load failed ({rc}): {ctx.error()}")


def svsig_batch(ctx: "Context", query_pac: np.ndarray, query_read_offset: np.ndarray, subject_seq_offset: int,
                subject_size: int, reads, min_indel: int, trf_mask=None):
    """discover_func over a list of SvRead tuples
    (query_id, subject_id, qdir, qoff, qend, qsize, soff, send, dist).
    Returns (per_read ctypes array, list of signature tuples)."""
    _bind_svsig(ctx.lib)
    n = len(reads)
    arr = (SvRead * n)()
    for i, r in enumerate(reads):
        for (name, _), v in zip(SvRead._fields_, r):
            setattr(arr[i], name, int(v))
    per = (SvSigRead * max(1, n))()
    sigs = C.POINTER(SvSig)()
    n_sigs = C.c_int()
    query_pac = np.ascontiguousarray(query_pac, dtype=np.uint8)
    offs = np.ascontiguousarray(query_read_offset, dtype=np.uint64)
    trf = None if trf_mask is None else np.ascontiguousarray(trf_mask, dtype=np.uint8)
    rc = ctx.lib.lesvgpu_svsig_batch(ctx.h, query_pac.ctypes.data, offs.ctypes.data, subject_seq_offset, subject_size,
                                     arr, n, min_indel, None if trf is None else trf.ctypes.data,
                                     per, C.byref(sigs), C.byref(n_sigs))
    if rc != OK:
        raise RuntimeError(f"lesvgpu_svsig_batch failed ({rc}): {ctx.error()}")
    return per, [sigs[i].astuple() for i in range(n_sigs.value)]


# ---- qx2svr (include/lesvgpu.h: lesvgpu_svr_batch) ------------------------------------------------------
M4_DTYPE = np.dtype([(n, np.int32) for n in ("qid", "qdir", "qoff", "qend", "qsize", "sid", "sdir", "soff", "send", "ssize")] +
                    [("ident_perc", np.float64), ("score", np.int32), ("pad_", np.int32)])          # M4Record, 56 bytes
SVREAD_DTYPE = np.dtype([(n, np.int32) for n in ("query_id", "subject_id", "qdir", "qoff", "qend", "qsize", "soff", "send", "dist")])
HSP_DTYPE = np.dtype([(n, np.int32) for n in ("qid", "qdir", "qoff", "qend", "qsize", "sid", "soff", "send")] +
                     [("ssize", np.int64)] + [(n, np.int32) for n in ("aln_len", "n_match", "chain_score", "ddf_score")] +
                     [("ops_off", np.uint64), ("flags", np.int32), ("pad_", np.int32)])            # lesvgpu_hsp, 72 bytes


class SvrOpts(C.Structure):
    _fields_ = [("min_seq_size", C.c_int), ("min_ident_perc", C.c_double), ("max_overhang", C.c_int)]


class SvrStats(C.Structure):
    _fields_ = [("n_m4", C.c_size_t), ("n_refine_jobs", C.c_size_t), ("n_chain_jobs", C.c_size_t)]


def svr_batch(ctx: "Context", query_pac, query_read_offset, s_offsets, s_sizes, m4: np.ndarray, min_seq_size=3000,
              min_ident=70.0, max_overhang=300, trf_mask=None):
    """process_one_m4_list over the M4Record array (M4_DTYPE) of one query volume -> (SvRead array (SVREAD_DTYPE), stats)."""
    lib = ctx.lib
    lib.lesvgpu_svr_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Seq), C.c_int, C.c_void_p, C.c_size_t,
                                      C.POINTER(SvrOpts), C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(SvrStats)]
    assert m4.dtype == M4_DTYPE and M4_DTYPE.itemsize == 56
    m4 = np.ascontiguousarray(m4)
    query_pac = np.ascontiguousarray(query_pac, dtype=np.uint8)
    offs = np.ascontiguousarray(query_read_offset, dtype=np.uint64)
    trf = None if trf_mask is None else np.ascontiguousarray(trf_mask, dtype=np.uint8)
    opts = SvrOpts(min_seq_size, min_ident, max_overhang)
    out, n, st = C.c_void_p(), C.c_int(), SvrStats()
    rc = lib.lesvgpu_svr_batch(ctx.h, query_pac.ctypes.data, offs.ctypes.data, _seq_array(s_offsets, s_sizes), len(s_offsets),
                               m4.ctypes.data, len(m4), C.byref(opts), None if trf is None else trf.ctypes.data,
                               C.byref(out), C.byref(n), C.byref(st))
    if rc != OK:
        raise RuntimeError(f"lesvgpu_svr_batch failed ({rc}): {ctx.error()}")
    if n.value == 0:
        return np.zeros(0, dtype=SVREAD_DTYPE), st
    buf = (C.c_char * (n.value * SVREAD_DTYPE.itemsize)).from_address(out.value)
    return np.frombuffer(buf, dtype=SVREAD_DTYPE).copy(), st


def svsig_batch_raw(ctx: "Context", query_pac, query_read_offset, subject_seq_offset, subject_size, reads: np.ndarray,
                    min_indel: int):
    """lesvgpu_svsig_batch over a SVREAD_DTYPE array; returns (status array, n_signatures) without per-record Python work."""
    _bind_svsig(ctx.lib)
    reads = np.ascontiguousarray(reads)
    assert reads.dtype == SVREAD_DTYPE
    n = len(reads)
    per = (SvSigRead * max(1, n))()
    sigs, n_sigs = C.POINTER(SvSig)(), C.c_int()
    rc = ctx.lib.lesvgpu_svsig_batch(ctx.h, query_pac.ctypes.data, query_read_offset.ctypes.data, subject_seq_offset, subject_size,
                                     C.cast(reads.ctypes.data, C.POINTER(SvRead)), n, min_indel, None, per, C.byref(sigs), C.byref(n_sigs))
    if rc != OK:
        raise RuntimeError(f"lesvgpu_svsig_batch failed ({rc}): {ctx.error()}")
    status = np.frombuffer(per, dtype=np.int32).reshape(-1, 9)[:n, 0].copy()
    return status, n_sigs.value


SVSIG_DTYPE = np.dtype([(n, np.int32) for n in ("qid", "qdir", "qfrom", "qto", "fqfrom", "fqto", "qsize", "sfrom", "sto",
                                               "fsfrom", "fsto", "ssize", "type")])           # lesvgpu_svsig, 52 bytes


def svsig_contigs_raw(ctx: "Context", query_pac, query_read_offset, subject_seqs, n_subject, reads: np.ndarray, min_indel: int,
                      want_sigs: bool = False):
    """lesvgpu_svsig_batch_contigs over a SVREAD_DTYPE array whose subject_id indexes subject_seqs (a _seq_array()).
    Returns (status array, n_signatures[, signature array (SVSIG_DTYPE), contig of every signature])."""
    lib = ctx.lib
    lib.lesvgpu_svsig_batch_contigs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Seq), C.c_int, C.POINTER(SvRead), C.c_int,
                                                C.c_int, C.c_void_p, C.POINTER(SvSigRead), C.POINTER(C.POINTER(SvSig)),
                                                C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_int)]
    reads = np.ascontiguousarray(reads)
    assert reads.dtype == SVREAD_DTYPE
    n = len(reads)
    per = (SvSigRead * max(1, n))()
    sigs, sid, n_sigs = C.POINTER(SvSig)(), C.POINTER(C.c_int)(), C.c_int()
    rc = lib.lesvgpu_svsig_batch_contigs(ctx.h, query_pac.ctypes.data, query_read_offset.ctypes.data, subject_seqs, n_subject,
                                         C.cast(reads.ctypes.data, C.POINTER(SvRead)), n, min_indel, None, per, C.byref(sigs),
                                         C.byref(sid), C.byref(n_sigs))
    if rc != OK:
        raise RuntimeError(f"lesvgpu_svsig_batch_contigs failed ({rc}): {ctx.error()}")
    status = np.frombuffer(per, dtype=np.int32).reshape(-1, 9)[:n, 0].copy()
    if not want_sigs:
        return status, n_sigs.value
    if n_sigs.value == 0:
        return status, 0, np.zeros(0, dtype=SVSIG_DTYPE), np.zeros(0, dtype=np.int32)
    assert C.sizeof(SvSig) == SVSIG_DTYPE.itemsize
    buf = (C.c_char * (n_sigs.value * SVSIG_DTYPE.itemsize)).from_address(C.addressof(sigs.contents))
    sb = (C.c_int * n_sigs.value).from_address(C.addressof(sid.contents))
    return status, n_sigs.value, np.frombuffer(buf, dtype=SVSIG_DTYPE).copy(), np.frombuffer(sb, dtype=np.int32).copy()


def map_batch_raw(ctx: "Context", query_pac, seq_arr, qfrom, qto, hitlist_size=3, max_hsps=5):
    """lesvgpu_map_batch -> HSP array (HSP_DTYPE); seq_arr from _seq_array()."""
    _bind_map(ctx.lib)
    opts = MapOpts(min_ddfs=3, hitlist_size=hitlist_size, max_hsps=max_hsps, want_ops=0)
    h, first, pool = C.POINTER(Hsp)(), C.POINTER(C.c_int)(), C.POINTER(C.c_uint8)()
    rc = ctx.lib.lesvgpu_map_batch(ctx.h, query_pac.ctypes.data, seq_arr, qfrom, qto, C.byref(opts), C.byref(h), C.byref(first),
                                   C.byref(pool))
    if rc != OK:
        raise RuntimeError(f"lesvgpu_map_batch failed ({rc}): {ctx.error()}")
    n = first[qto - qfrom]
    if n == 0:
        return np.zeros(0, dtype=HSP_DTYPE)
    assert C.sizeof(Hsp) == HSP_DTYPE.itemsize
    buf = (C.c_char * (n * HSP_DTYPE.itemsize)).from_address(C.addressof(h.contents))
    return np.frombuffer(buf, dtype=HSP_DTYPE).copy()


def hsps_to_m4(h: np.ndarray) -> np.ndarray:
    """What qx2map prints and qx2m4x parses back (print_one_m4_result, mecat_results.c:256-280; m4x.cpp:13-31):
    reverse-strand query coordinates flipped, identity through its 6-significant-digit text."""
    m = np.zeros(len(h), dtype=M4_DTYPE)
    rev = h["qdir"] == 1
    m["qid"], m["qdir"], m["qsize"] = h["qid"], h["qdir"], h["qsize"]
    m["qoff"] = np.where(rev, h["qsize"] - h["qend"], h["qoff"])
    m["qend"] = np.where(rev, h["qsize"] - h["qoff"], h["qend"])
    m["sid"], m["soff"], m["send"], m["ssize"] = h["sid"], h["soff"], h["send"], h["ssize"]
    ident = 100.0 * h["n_match"] / np.maximum(h["aln_len"], 1)
    m["ident_perc"] = [float("%g" % x) for x in ident]
    m["score"] = 8 * h["aln_len"] - h["n_match"]
    return m


# ---- qx2map candidate stage (include/lesvgpu.h: lesvgpu_index_build .. lesvgpu_map_candidates) --------
class Seq(C.Structure):
    _fields_ = [("seq_offset", C.c_uint64), ("seq_size", C.c_uint64)]


class MapOpts(C.Structure):
    _fields_ = [("min_ddfs", C.c_int), ("hitlist_size", C.c_int), ("max_hsps", C.c_int), ("memsc_kmer_size", C.c_int),
                ("memsc_kmer_window", C.c_int), ("memsc_score", C.c_int), ("qcov_hsp_res", C.c_int),
                ("perc_identity", C.c_double), ("skip_overhang", C.c_int), ("want_ops", C.c_int)]


class Hsp(C.Structure):
    _fields_ = [("qid", C.c_int), ("qdir", C.c_int), ("qoff", C.c_int), ("qend", C.c_int), ("qsize", C.c_int),
                ("sid", C.c_int), ("soff", C.c_int), ("send", C.c_int), ("ssize", C.c_int64),
                ("aln_len", C.c_int), ("n_match", C.c_int), ("chain_score", C.c_int), ("ddf_score", C.c_int),
                ("ops_off", C.c_size_t), ("flags", C.c_int)]


class ExtendTiming(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ihf_ms", "sort_ms", "chain_ms", "trace_ms", "d2h_ms")] + \
               [(n, C.c_uint64) for n in ("n_windows", "n_seeds", "n_candidates", "n_unsupported", "n_overhang_skipped")] + \
               [("d2h_bytes", C.c_size_t)] + [(n, C.c_float) for n in ("prep_ms", "gap_ms", "finish_ms")] + \
               [("n_gap_jobs", C.c_uint64), ("n_gaps_class", C.c_uint64 * 4)]


class SubseqHit(C.Structure):
    _fields_ = [("qid", C.c_int), ("qdir", C.c_int), ("qoff", C.c_int), ("sid", C.c_int), ("soff", C.c_int),
                ("sfrom", C.c_size_t), ("sto", C.c_size_t), ("score", C.c_int)]

    def astuple(self):
        return (self.qid, self.qdir, self.qoff, self.sid, self.soff, self.sfrom, self.sto, self.score)


class InitHit(C.Structure):
    _fields_ = [("qid", C.c_int), ("qdir", C.c_int), ("qoff", C.c_int), ("qbeg", C.c_int), ("qend", C.c_int),
                ("sid", C.c_int), ("soff", C.c_int64), ("sbeg", C.c_int64), ("send", C.c_int64), ("score", C.c_int)]


class KmerMatch(C.Structure):
    _fields_ = [("context", C.c_int), ("qoff", C.c_int), ("soff", C.c_int64)]


class MapTiming(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "lookup_ms", "emit_ms", "sort_ms", "chain_ms", "d2h_ms")] + \
               [(n, C.c_uint64) for n in ("n_kmers", "n_matches", "n_segments", "n_hits")] + \
               [("h2d_bytes", C.c_size_t), ("d2h_bytes", C.c_size_t)]


def _bind_map(lib):
    lib.lesvgpu_index_build.argtypes = [C.c_void_p, C.POINTER(Seq), C.c_int, C.c_int, C.c_int, C.c_int]
    lib.lesvgpu_index_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.lesvgpu_index_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lesvgpu_map_candidates.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Seq), C.c_int, C.c_int, C.POINTER(MapOpts),
                                           C.POINTER(C.POINTER(SubseqHit)), C.POINTER(C.POINTER(C.c_int))]
    lib.lesvgpu_map_last_init_hits.argtypes = [C.c_void_p, C.POINTER(C.POINTER(InitHit)), C.POINTER(C.POINTER(C.c_int))]
    lib.lesvgpu_map_last_matches.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.POINTER(KmerMatch)), C.POINTER(C.c_size_t)]
    lib.lesvgpu_map_last_timing.argtypes = [C.c_void_p, C.POINTER(MapTiming)]
    lib.lesvgpu_map_batch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Seq), C.c_int, C.c_int, C.POINTER(MapOpts),
                                      C.POINTER(C.POINTER(Hsp)), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_uint8))]
    lib.lesvgpu_map_last_extend_timing.argtypes = [C.c_void_p, C.POINTER(ExtendTiming)]


def _seq_array(offsets, sizes):
    n = len(offsets)
    arr = (Seq * max(1, n))()
    for i in range(n):
        arr[i].seq_offset, arr[i].seq_size = int(offsets[i]), int(sizes[i])
    return arr


def index_build(ctx: "Context", seq_offsets, seq_sizes, kmer_size=15, kmer_window=10, max_kmer_occ=200):
    """build_lookup_table over the subject volume loaded with subject_load()."""
    _bind_map(ctx.lib)
    rc = ctx.lib.lesvgpu_index_build(ctx.h, _seq_array(seq_offsets, seq_sizes), len(seq_offsets), kmer_size, kmer_window,
                                     max_kmer_occ)
    if rc != OK:
        raise RuntimeError(f"lesvgpu_index_build failed ({rc}): {ctx.error()}")


def index_export(ctx: "Context"):
    """(offset_list u64[n], hashes u64[d], starts u32[d+1]) -- LookupTable in the reference's order."""
    _bind_map(ctx.lib)
    n, d = C.c_uint64(), C.c_uint64()
    ctx.lib.lesvgpu_index_stats(ctx.h, C.byref(n), C.byref(d))
    off = np.zeros(max(1, n.value), dtype=np.uint64)
    hs = np.zeros(max(1, d.value), dtype=np.uint64)
    st = np.zeros(d.value + 1, dtype=np.uint32)
    rc = ctx.lib.lesvgpu_index_export(ctx.h, off.ctypes.data, hs.ctypes.data, st.ctypes.data)
    if rc != OK:
        raise RuntimeError(f"lesvgpu_index_export failed ({rc}): {ctx.error()}")
    return off[:n.value], hs[:d.value], st


def map_candidates(ctx: "Context", query_pac, q_offsets, q_sizes, qfrom, qto, min_ddfs=3, hitlist_size=3):
    """Candidate windows (HbnSubseqHit tuples) per read of [qfrom, qto): list of lists."""
    _bind_map(ctx.lib)
    query_pac = np.ascontiguousarray(query_pac, dtype=np.uint8)
    opts = MapOpts(min_ddfs=min_ddfs, hitlist_size=hitlist_size)
    w, first = C.POINTER(SubseqHit)(), C.POINTER(C.c_int)()
    rc = ctx.lib.lesvgpu_map_candidates(ctx.h, query_pac.ctypes.data, _seq_array(q_offsets, q_sizes), qfrom, qto,
                                        C.byref(opts), C.byref(w), C.byref(first))
    if rc != OK:
        raise RuntimeError(f"lesvgpu_map_candidates failed ({rc}): {ctx.error()}")
    return [[w[j].astuple() for j in range(first[i], first[i + 1])] for i in range(qto - qfrom)]


def map_last_matches(ctx: "Context", qfrom=0):
    """DDFKmerMatch records of the last map_candidates call as an (n, 3) int64 array (context, qoff, soff),
    sorted by (context, soff, qoff)."""
    _bind_map(ctx.lib)
    m, n = C.POINTER(KmerMatch)(), C.c_size_t()
    rc = ctx.lib.lesvgpu_map_last_matches(ctx.h, qfrom, C.byref(m), C.byref(n))
    if rc != OK:
        raise RuntimeError(f"lesvgpu_map_last_matches failed ({rc}): {ctx.error()}")
    if n.value == 0:
        return np.zeros((0, 3), dtype=np.int64)
    raw = np.ctypeslib.as_array(C.cast(m, C.POINTER(C.c_int32)), shape=(n.value, 4))
    out = np.empty((n.value, 3), dtype=np.int64)
    out[:, 0] = raw[:, 0]; out[:, 1] = raw[:, 1]
    out[:, 2] = raw[:, 2:4].copy().view(np.int64)[:, 0]
    return out


def map_last_timing(ctx: "Context") -> MapTiming:
    _bind_map(ctx.lib)
    t = MapTiming()
    ctx.lib.lesvgpu_map_last_timing(ctx.h, C.byref(t))
    return t


def map_batch(ctx: "Context", query_pac, q_offsets, q_sizes, qfrom, qto, hitlist_size=3, max_hsps=5, want_ops=False,
              min_ddfs=3, skip_overhang=False):
    """chain_and_extend for the reads [qfrom, qto): per read a list of dicts (HSP fields; `ops` = alignment
    columns as a uint8 array when want_ops)."""
    _bind_map(ctx.lib)
    query_pac = np.ascontiguousarray(query_pac, dtype=np.uint8)
    opts = MapOpts(min_ddfs=min_ddfs, hitlist_size=hitlist_size, max_hsps=max_hsps, want_ops=1 if want_ops else 0,
                   skip_overhang=1 if skip_overhang else 0)
    h, first, pool = C.POINTER(Hsp)(), C.POINTER(C.c_int)(), C.POINTER(C.c_uint8)()
    rc = ctx.lib.lesvgpu_map_batch(ctx.h, query_pac.ctypes.data, _seq_array(q_offsets, q_sizes), qfrom, qto, C.byref(opts),
                                   C.byref(h), C.byref(first), C.byref(pool))
    if rc != OK:
        raise RuntimeError(f"lesvgpu_map_batch failed ({rc}): {ctx.error()}")
    out = []
    for i in range(qto - qfrom):
        lst = []
        for j in range(first[i], first[i + 1]):
            d = {n: getattr(h[j], n) for n, _ in Hsp._fields_}
            if want_ops:
                d["ops"] = np.ctypeslib.as_array(pool, shape=(d["ops_off"] + d["aln_len"],))[d["ops_off"]:].copy() if d["aln_len"] else np.zeros(0, np.uint8)
            lst.append(d)
        out.append(lst)
    return out


def map_last_extend_timing(ctx: "Context") -> ExtendTiming:
    _bind_map(ctx.lib)
    t = ExtendTiming()
    ctx.lib.lesvgpu_map_last_extend_timing(ctx.h, C.byref(t))
    return t
