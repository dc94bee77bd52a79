"""chain_core.cuh (the per-segment chaining code of the mapper kernels, built for the host by
tests/emul) against the compiled reference's chaining_find_candidates (src/algo/chain_dp.c:273-395,
incl. join_adjacent_chains): same hits in the same order, same seed lists."""
import ctypes as C

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built")


class ChainSeed(C.Structure):      # src/algo/chain_dp.h:13-19
    _fields_ = [("length", C.c_int), ("qoff", C.c_int64), ("soff", C.c_int64), ("sdir", C.c_int), ("hash", C.c_uint64)]


class InitHit(C.Structure):        # src/corelib/gapped_candidate.h:10-27
    _fields_ = [(n, C.c_int) for n in ("qid", "qdir", "qoff", "qbeg", "qend", "qsize", "sid", "sdir")] + \
               [(n, C.c_size_t) for n in ("soff", "sbeg", "send", "ssize")] + \
               [("score", C.c_int), ("chain_seed_array", C.c_void_p), ("chain_seed_offset", C.c_int), ("chain_seed_count", C.c_int)]


class KVec(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p)]


def ref_chain(q, s, ln, is_mem, min_cnt, min_score):
    lib, _ = H.ref()
    lib.ChainWorkDataNew.restype = C.c_void_p
    lib.ChainWorkDataFree.argtypes = [C.c_void_p]
    lib.chaining_find_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = len(q)
    arr = (ChainSeed * n)()
    for i in range(n):
        arr[i].length, arr[i].qoff, arr[i].soff, arr[i].sdir, arr[i].hash = int(ln[i]), int(q[i]), int(s[i]), 0, i
    data = lib.ChainWorkDataNew(min_cnt, min_score)
    hits, seeds = KVec(), KVec()
    lib.chaining_find_candidates(data, arr, n, int(is_mem), 0, C.byref(hits), C.byref(seeds))
    out = []
    if hits.n:
        ha = C.cast(hits.a, C.POINTER(InitHit))
        for i in range(hits.n):
            h = ha[i]
            cs = C.cast(h.chain_seed_array, C.POINTER(ChainSeed))
            ids = [int(cs[k].hash) for k in range(h.chain_seed_count)]
            first, last = cs[0], cs[h.chain_seed_count - 1]
            out.append((h.score, first.qoff, last.qoff + last.length, first.soff, last.soff + last.length, ids))
    lib.ChainWorkDataFree(data)
    H._libc.free(hits.a)
    H._libc.free(seeds.a)
    return out


_emul = {}


def emul_chain(q, s, ln, is_mem, min_cnt, min_score):
    if "l" not in _emul:
        import os
        _emul["l"] = C.CDLL(os.path.join(H.ROOT, "tests", "emul", "libchain_emul.so"))
    lib = _emul["l"]
    n = len(q)
    q = np.ascontiguousarray(q, dtype=np.int32); s = np.ascontiguousarray(s, dtype=np.int32)
    ln = np.ascontiguousarray(ln, dtype=np.int32)
    hits = np.zeros(6 * 20, dtype=np.int32)
    seeds = np.zeros(max(n, 1), dtype=np.int32)
    nh = lib.lesv_emul_chain(n, H._ptr(q), H._ptr(s), H._ptr(ln), int(is_mem), min_cnt, min_score, H._ptr(hits), H._ptr(seeds))
    out, pos = [], 0
    for h in range(nh):
        sc, qb, qe, sb, se, cnt = [int(x) for x in hits[6 * h:6 * h + 6]]
        out.append((sc, qb, qe, sb, se, [int(x) for x in seeds[pos:pos + cnt]]))
        pos += cnt
    return out


def make_seeds(rng, mode):
    """Seed lists sorted the way the mapper hands them over: by subject offset, then query offset."""
    pts = []
    if mode == 0:      # one noisy diagonal + random hits (DDF stage on a true placement)
        L = int(rng.integers(2000, 30000))
        q = 0; s = int(rng.integers(0, 1000))
        while q < L:
            if rng.random() < 0.35:
                pts.append((q, s))
            step = int(rng.integers(1, 40)); q += step; s += step + int(rng.integers(-2, 3))
        for _ in range(int(rng.integers(0, 60))):
            pts.append((int(rng.integers(0, L)), int(rng.integers(0, L + 2000))))
    elif mode == 1:    # several colinear pieces separated by gaps (join_adjacent_chains) and a repeat copy
        q = 0; s = 5000
        for piece in range(int(rng.integers(2, 6))):
            n = int(rng.integers(80, 400))
            for _ in range(n):
                pts.append((q, s)); step = int(rng.integers(5, 30)); q += step; s += step + int(rng.integers(-1, 2))
            q += int(rng.integers(50, 2500)); s += int(rng.integers(50, 22000))
        if rng.random() < 0.5:
            pts += [(a, b + 40000) for a, b in pts[: len(pts) // 3]]
    elif mode == 2:    # regular lattice: many equal scores, ties everywhere
        n = int(rng.integers(3, 120)); d = int(rng.choice([1, 7, 15, 16, 40]))
        pts = [(i * d, 100 + i * d) for i in range(n)] + [(i * d, 100 + i * d + 3 * d) for i in range(0, n, 2)]
    else:              # sparse random
        n = int(rng.integers(1, 80))
        pts = [(int(rng.integers(0, 3000)), int(rng.integers(0, 3000))) for _ in range(n)]
    pts = sorted(set(pts), key=lambda p: (p[1], p[0]))
    q = np.array([p[0] for p in pts]); s = np.array([p[1] for p in pts])
    return q, s


@pytest.mark.parametrize("is_mem", [0, 1])
def test_chain_core_equals_reference(is_mem):
    rng = np.random.default_rng(11 + is_mem)
    n_hits = 0
    for it in range(400):
        q, s = make_seeds(rng, it % 4)
        if is_mem:
            ln = rng.integers(10, 40, len(q))
            min_cnt, min_score = 1, 30
        else:
            ln = np.full(len(q), 15)
            min_cnt, min_score = 3, 36
        a = ref_chain(q, s, ln, is_mem, min_cnt, min_score)
        b = emul_chain(q, s, ln, is_mem, min_cnt, min_score)
        assert a == b, (it, len(q), a[:2], b[:2])
        n_hits += len(a)
    assert n_hits > 300


def test_warp_schedule_of_the_chaining_dp_equals_the_sequential_loop():
    """chain_warp.cuh scores 32 predecessors at once and replays the order-dependent part (strict '>', t[] marks, skip
    counter, early stop).  Its lane-by-lane host emulation (tests/emul/chain_emul.cpp) must give the f / p / v arrays
    of the sequential loop on random, collinear, tie-heavy and repeat-like seed lists."""
    import os
    lib = C.CDLL(os.path.join(H.ROOT, "tests", "emul", "libchain_emul.so"))
    lib.lesv_emul_chain_warp_diff.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(77)
    total = 0
    for it in range(400):
        mode = it % 5
        n = int(rng.integers(1, 700))
        if mode == 0:                                   # noisy collinear chain (the normal case)
            q = np.cumsum(rng.integers(1, 60, n)); s = q + rng.integers(-30, 30, n) + int(rng.integers(0, 5000))
        elif mode == 1:                                 # several parallel diagonals: many near-ties
            q = np.sort(rng.integers(0, 4000, n)); s = q + rng.choice([0, 7, 300, 1400, 1600], n)
        elif mode == 2:                                 # lattice: exact ties everywhere, long marked runs (skip counter)
            q = (np.arange(n) // 8) * 10; s = q + (np.arange(n) % 8) * 10
        elif mode == 3:                                 # random cloud
            q = rng.integers(0, 20000, n); s = rng.integers(0, 20000, n)
        else:                                           # dense tandem-like: period 37
            q = np.arange(n) * 3; s = (np.arange(n) * 3) % 37 + (np.arange(n) // 40) * 37
        order = np.lexsort((q, s))                      # the kernels feed seeds sorted by (s, q)
        q, s = q[order].astype(np.int32), s[order].astype(np.int32)
        ln = np.full(n, 10 if it % 2 else 15, dtype=np.int32)
        for is_mem in (0, 1):
            bad = lib.lesv_emul_chain_warp_diff(n, H._ptr(q), H._ptr(s), H._ptr(ln), is_mem, 1 if is_mem else 3, 30)
            assert bad == 0, (it, mode, n, is_mem, bad)
            total += n
    assert total > 100000
