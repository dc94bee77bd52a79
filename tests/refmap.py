"""ctypes views of the compiled reference's mapper internals (oracle/_ref/libhbn_ref.so), used to
generate golden fixtures and -- when oracle/_ref is present -- to check the GPU stages directly:
build_lookup_table (src/algo/lookup_table.c:199) and hbn_find_kmer_matches_mt
(src/algo/hbn_word_finder.c:337).  Test infrastructure only."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

import helpers as H

BIN = os.path.join(H.ROOT, "oracle", "_ref", "bin")


class LookupTable(C.Structure):      # src/algo/lookup_table.h:12-23
    _fields_ = [("offset_list", C.POINTER(C.c_uint64)), ("offset_list_size", C.c_uint64),
                ("lki_array", C.c_void_p), ("lki_count", C.c_uint64)]


def _lib():
    lib, _ = H.ref()
    lib.seqdb_load.restype = C.c_void_p
    lib.seqdb_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    lib.build_lookup_table.restype = C.POINTER(LookupTable)
    lib.build_lookup_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.HbnWordFindDataNew.restype = C.c_void_p
    lib.HbnWordFindDataNew.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.hbn_find_kmer_matches_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return lib


def ref_index(db_dir, k=15, w=10, max_occ=200):
    """(offset_list, hashes, counts) of the reference's LookupTable for <db_dir>/subject."""
    lib = _lib()
    sdb = lib.seqdb_load(db_dir.encode(), b"subject", 0)
    lk = lib.build_lookup_table(sdb, k, w, max_occ, 1)
    n, d = lk.contents.offset_list_size, lk.contents.lki_count
    off = np.ctypeslib.as_array(lk.contents.offset_list, shape=(max(1, n),))[:n].copy()
    raw = np.ctypeslib.as_array(C.cast(lk.contents.lki_array, C.POINTER(C.c_uint64)), shape=(max(1, d), 3))[:d]
    hashes = raw[:, 0].copy()
    starts = raw[:, 1].copy()
    cnt = (raw[:, 2] & 0xFFFFFFFF).astype(np.int64)
    assert np.array_equal(starts, np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.uint64)) or d == 0
    return off, hashes, cnt, (lib, sdb, lk)


def ref_matches(db_dir, k=15, max_occ=200, handles=None):
    """All DDFKmerMatch records of hbn_find_kmer_matches_mt with one thread over the whole query volume,
    as an (n, 3) int64 array (context, qoff, soff) sorted by (context, soff, qoff)."""
    lib = _lib()
    if handles is None:
        handles = ref_index(db_dir, k)[3]
    _, sdb, lk = handles
    qdb = lib.seqdb_load(db_dir.encode(), b"query", 0)
    nq = H.LesvDb(db_dir, "query").num_seqs
    data = lib.HbnWordFindDataNew(lk, sdb, qdb, 1, k, 1, max_occ)
    soff_max = np.full(2 * nq, np.iinfo(np.int64).max, dtype=np.int64)
    lib.hbn_find_kmer_matches_mt(data, 0, nq, soff_max.ctypes.data)
    kmi = C.cast(data + 40, C.POINTER(C.c_void_p))[0]                 # HbnWordFindData.kmi_array
    rows = []
    for q in range(nq):
        ptr = C.cast(kmi + 16 * q, C.POINTER(C.c_void_p))[0]
        cnt = C.cast(kmi + 16 * q + 8, C.POINTER(C.c_int))[0]
        if not ptr or cnt == 0:
            continue
        raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), shape=(cnt, 4)).copy()
        m = np.empty((cnt, 3), dtype=np.int64)
        m[:, 0] = raw[:, 0]; m[:, 1] = raw[:, 1]
        m[:, 2] = raw[:, 2:4].copy().view(np.int64)[:, 0]
        rows.append(m)
    if not rows:
        return np.zeros((0, 3), dtype=np.int64)
    m = np.concatenate(rows)
    order = np.lexsort((m[:, 1], m[:, 2], m[:, 0]))
    return m[order]


def parse_subseq(path):
    """`qx2map -outfmt subseq` lines -> {qid: [(qid, qdir, qoff, sid, soff, sfrom, sto, score), ...]} in file order."""
    out = {}
    with open(path) as f:
        for line in f:
            v = tuple(int(x) for x in re.findall(r"= (\d+)", line))
            if len(v) == 8:
                out.setdefault(v[0], []).append(v)
    return out


def run_qx2map(work, args, out_name, threads=1):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    with open(os.path.join(work, out_name), "wb") as f, open(os.path.join(work, "log.txt"), "ab") as log:
        subprocess.run([os.path.join(BIN, "qx2map")] + args + ["-num_threads", str(threads), "reads.fa", "ref.fa"],
                       cwd=work, check=True, env=env, stdout=f, stderr=log)
