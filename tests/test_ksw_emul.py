"""The K-ksw per-thread code (lesv_b200/csrc/ksw_core.cuh), run on the CPU by tests/emul, against
the oracle: CIGARs for random thread counts, and every in-matrix traceback byte."""
import numpy as np
import pytest

import helpers as H
from lesv_b200.synth import mutate


@pytest.mark.parametrize("seed", [0, 1])
def test_emulated_kernel_cigar_equals_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    n = 0
    for it in range(1500):
        q, t, w = H.random_pair(rng, it)
        if not H.oracle_cells(len(q), len(t), w)[1]:
            continue  # band cut: the library answers ok = 0 without launching
        need = H.min_threads(len(q), len(t), w)
        nt = max(need, int(rng.choice([need, need + 1, 64, 128])))
        a, _ = H.oracle_ksw(q, t, w)
        b = H.emul_ksw(q, t, w, nt, eager=bool(it & 1))     # both carry conventions of thread_row_step
        assert a == b, (it, len(q), len(t), w, nt)
        n += 1
    assert n > 800


@pytest.mark.parametrize("k", [2, 8])
def test_emulated_skewed_schedule_equals_oracle(k):
    """kSyncSkew: warps `k` rows apart per 32-block group, a barrier only every k rows; the emulator
    runs the warps of a barrier interval in random order (worst legal interleaving)."""
    rng = np.random.default_rng(300 + k)
    n = 0
    for it in range(1200):
        q, t, w = H.random_pair(rng, it, max_len=2500 if it % 7 == 0 else 700)
        if not H.oracle_cells(len(q), len(t), w)[1]:
            continue
        need = (H.min_threads(len(q), len(t), w) + 31) // 32 * 32
        nt = need + 32 * int(rng.integers(0, 2))
        a, _ = H.oracle_ksw(q, t, w)
        b = H.emul_ksw_skew(q, t, w, nt, k, it)
        assert a == b, (it, len(q), len(t), w, nt)
        n += 1
    assert n > 600


def test_emulated_cluster_skew_equals_oracle():
    """Cluster flavour of the skewed schedule: the CTAs of a cluster are the groups (lock-step inside)."""
    rng = np.random.default_rng(77)
    n = 0
    for it in range(500):
        q, t, w = H.random_pair(rng, it, max_len=3000 if it % 3 == 0 else 700)
        if not H.oracle_cells(len(q), len(t), w)[1]:
            continue
        ncta = int(rng.choice([2, 4]))
        need = H.min_threads(len(q), len(t), w)
        per = ((need + ncta - 1) // ncta + 31) // 32 * 32
        a, _ = H.oracle_ksw(q, t, w)
        b = H.emul_ksw_skew_group(q, t, w, per * ncta, per, int(rng.choice([4, 8])), it)
        assert a == b, (it, len(q), len(t), w, per, ncta)
        n += 1
    assert n > 250


def test_emulated_kernel_traceback_bytes_equal_oracle():
    rng = np.random.default_rng(7)
    cells = 0
    for it in range(400):
        q, t, w = H.random_pair(rng, it, max_len=400)
        ql, tl = len(q), len(t)
        if not H.oracle_cells(ql, tl, w)[1]:
            continue
        ww = w if w >= 0 else max(ql, tl)
        A = H.oracle_tb(q, t, w)
        B = H.emul_tb(q, t, w, H.min_threads(ql, tl, w))
        for r in range(ql + tl - 1):
            st0 = max(0, r - ql + 1, (r - ww + 1) >> 1)
            en0 = min(tl - 1, r, (r + ww) >> 1)
            st, en = st0 // 16 * 16, en0 // 16 * 16 + 15
            lo, hi = max(st, r - ql + 1), min(en, r, tl - 1)   # aligned band ∩ DP matrix
            assert np.array_equal(A[r, lo - st:hi - st + 1], B[r, lo - st:hi - st + 1]), (it, ql, tl, w, r)
            cells += hi - lo + 1
    assert cells > 5_000_000


def test_strip_schedule_equals_oracle():
    """ksw_dp_strip_kernel's schedule (threads keep one block for the whole pass, strips of columns run one after the
    other and hand their column-15 values over through a per-row edge buffer), emulated on the host for several strip
    widths, full-matrix and banded alignments."""
    rng = np.random.default_rng(3)
    lib = H.emul()
    n = 0
    for it in range(24):
        ln = int(rng.integers(200, 1200))
        q = rng.integers(0, 4, ln, dtype=np.uint8)
        t = np.ascontiguousarray(mutate(q, 0.15, rng))
        if it % 3 == 0:
            t = np.ascontiguousarray(np.concatenate([t[:ln // 3], rng.integers(0, 4, int(rng.integers(50, 400)), dtype=np.uint8), t[ln // 3:]]))
        w = max(len(q), len(t)) if it % 2 == 0 else abs(len(q) - len(t)) + int(rng.integers(40, 300))
        exp, _ = H.oracle_ksw(q, t, w)
        for sw in (2, 5, 64):
            lib.lesv_emul_set_strip(sw)
            try:
                got = H.emul_ksw(q, t, w, 32, eager=True)
            finally:
                lib.lesv_emul_set_strip(0)
            assert got == exp, (it, len(q), len(t), w, sw)
            n += 1
    assert n == 72


def test_folded_strip_traceback_has_a_slot_for_every_block():
    """Column strips store row r of a full-matrix alignment at physical row r % qlen, block B at slot B + r / qlen
    (ksw_core.cuh: fold_row_bytes): no two stored blocks may share a slot, for any ratio of the lengths."""
    import ctypes
    lib = H.emul()
    lib.lesv_emul_fold_check.restype = ctypes.c_long
    rng = np.random.default_rng(9)
    shapes = [(1, 1), (1, 500), (500, 1), (16, 16), (17, 33), (1000, 1000), (999, 4001), (4001, 999), (50, 3000), (3000, 50)]
    shapes += [(int(rng.integers(1, 2500)), int(rng.integers(1, 2500))) for _ in range(150)]
    for qlen, tlen in shapes:
        assert lib.lesv_emul_fold_check(qlen, tlen) == 0, (qlen, tlen)


def test_inner_blocks_through_the_band_edge_path_equal_oracle():
    """On the GPU a block strictly inside the band takes the lean path only if all 32 blocks of its warp do; in a warp
    that holds a band edge it runs through the general path with an all-ones fresh-column mask.  The emulator forces
    that for every row (lesv_emul_set_lean(0)): CIGARs must not change."""
    lib = H.emul()
    rng = np.random.default_rng(321)
    lib.lesv_emul_set_lean(0)
    try:
        n = 0
        for it in range(500):
            q, t, w = H.random_pair(rng, it)
            if not H.oracle_cells(len(q), len(t), w)[1]:
                continue
            need = H.min_threads(len(q), len(t), w)
            a, _ = H.oracle_ksw(q, t, w)
            assert a == H.emul_ksw(q, t, w, need + int(it % 3), eager=bool(it & 1)), (it, len(q), len(t), w)
            n += 1
        assert n > 250
    finally:
        lib.lesv_emul_set_lean(1)
