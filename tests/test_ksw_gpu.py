"""K-ksw on the GPU through the C ABI (lesvgpu_ksw_batch) against the oracle: bit-exact CIGARs."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from lesv_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _check(ctx, cases):
    res, cigs = ctx.ksw_batch([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases])
    for i, (q, t, w) in enumerate(cases):
        exp, _ = H.oracle_ksw(q, t, w if w != 0 else int(max(len(q), len(t)) * 0.1))
        ok = H.oracle_cells(len(q), len(t), w if w != 0 else int(max(len(q), len(t)) * 0.1))[1]
        if not ok:
            assert not res[i].ok, i
            continue
        assert res[i].ok, (i, len(q), len(t), w)
        got = [int(x) for x in cigs[i]]
        assert got == exp, (i, len(q), len(t), w, H.cigar_str(exp)[:80], H.cigar_str(got)[:80])
        aln = sum(x >> 4 for x in exp)
        assert res[i].aln_size == aln


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_pairs_bit_exact(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    cases = [H.random_pair(rng, it) for it in range(600)]
    _check(ctx, cases)


def test_noisy_reads_svsig_shape(ctx):
    """10-12 kb reads at 12 % error with planted indels, band = 1.2 * dist as qx2svsig uses."""
    from lesv_b200.synth import mutate
    rng = np.random.default_rng(3)
    cases = []
    for it in range(12):
        q = rng.integers(0, 4, int(rng.integers(3000, 12000)), dtype=np.uint8)
        t = q.copy()
        if it % 3 == 1:
            k = int(rng.integers(500, len(t) - 500)); t = np.concatenate([t[:k], t[k + 300:]])
        if it % 3 == 2:
            k = int(rng.integers(500, len(t) - 500)); t = np.concatenate([t[:k], rng.integers(0, 4, 400, dtype=np.uint8), t[k:]])
        q = mutate(q, 0.12, rng)
        w = int(int((len(q) + len(t)) * 12.0 / 200) * 1.2) + 400
        cases.append((np.ascontiguousarray(q), np.ascontiguousarray(t), w))
    _check(ctx, cases)


def test_identity_counts(ctx):
    rng = np.random.default_rng(9)
    from lesv_b200.synth import mutate
    q = rng.integers(0, 4, 2000, dtype=np.uint8)
    t = mutate(q, 0.1, rng)
    res, cigs = ctx.ksw_batch([q], [t], [300])
    qi = si = m = 0
    for x in cigs[0]:
        n, op = int(x) >> 4, int(x) & 15
        if op == 0:
            m += int((q[qi:qi + n] == t[si:si + n]).sum()); qi += n; si += n
        elif op == 1:
            qi += n
        else:
            si += n
    assert (qi, si) == (len(q), len(t))
    assert res[0].n_match == m


def test_wide_bands_run_on_clusters(ctx):
    """Band windows wider than the single-CTA tiers: without a band (w = -1) these run as column strips or, up to 736
    blocks, in one CTA of 24 warps; with LESV_KSW_NO_STRIPS they take thread-block clusters (carries through DSMEM, see
    test_environment_switches_select_tested_paths); banded windows of that width: test_band_classes_around_the_single_cta_limit."""
    from lesv_b200.synth import mutate
    rng = np.random.default_rng(21)
    cases = []
    for n in (9100, 11000, 15000):                     # 1 CTA x 576 / 704 threads, then 4 strips
        q = rng.integers(0, 4, n, dtype=np.uint8)
        t = np.ascontiguousarray(mutate(q, 0.12, rng))
        cases.append((q, t, -1))                       # no band: max(qlen, tlen)
    q = rng.integers(0, 4, 9000, dtype=np.uint8)
    t = np.ascontiguousarray(np.concatenate([q[:4000], q[5500:]]))      # 1.5 kb deletion
    cases.append((q, t, max(len(q), len(t))))
    q = rng.integers(0, 4, 24000, dtype=np.uint8)       # 6 strips
    cases.append((q, np.ascontiguousarray(mutate(q, 0.1, rng)), -1))
    _check(ctx, cases)


def test_many_small_and_mixed_shapes_in_one_batch(ctx):
    """One batch spanning many launch classes (concurrent launches, shared arena)."""
    from lesv_b200.synth import mutate
    rng = np.random.default_rng(5)
    cases = []
    for it in range(300):
        n = int(rng.integers(200, 6000))
        q = rng.integers(0, 4, n, dtype=np.uint8)
        t = np.ascontiguousarray(mutate(q, float(rng.uniform(0.02, 0.2)), rng))
        if len(t) == 0:
            t = q.copy()
        w = int(abs(len(q) - len(t)) + rng.integers(1, 1 + n // 3))
        cases.append((q, t, w))
    _check(ctx, cases)


def test_full_matrix_alignments_as_column_strips(ctx):
    """Alignments without a band and wider than one CTA run as column strips (ksw_dp_strip_kernel): CIGARs must equal
    the oracle's, here on 12-20 kb pairs (2-3 strips, some with a 2 kb insertion)."""
    from lesv_b200.synth import mutate
    c = ctx
    rng = np.random.default_rng(11)
    qs, ts = [], []
    for i in range(4):
        n = int(rng.integers(11000, 16000))
        t = rng.integers(0, 4, n, dtype=np.uint8)
        q = np.ascontiguousarray(mutate(t, 0.12, rng))
        if i % 2 == 0:
            q = np.ascontiguousarray(np.concatenate([q[:n // 2], rng.integers(0, 4, 2000, dtype=np.uint8), q[n // 2:]]))
        qs.append(q); ts.append(t)
    res, cigs = c.ksw_batch(qs, ts, [-1] * len(qs))
    for q, t, r, cg in zip(qs, ts, res, cigs):
        exp, _ = H.oracle_ksw(q, t, -1)
        assert r.ok and [int(x) for x in cg] == exp


def test_column_strips_fold_rows_for_any_length_ratio(ctx):
    """The strips store row r at physical row r % qlen, block B at slot B + r / qlen (ksw_kernels.cuh: fold_tb_bytes).
    Lengths far apart exercise one to five passes over the physical rows, in both directions; the unrelated pair makes
    the walk leave the diagonal early (long leading / trailing gaps)."""
    from lesv_b200.synth import mutate
    rng = np.random.default_rng(77)
    qs, ts = [], []
    t = rng.integers(0, 4, 13000, dtype=np.uint8)
    qs.append(np.ascontiguousarray(mutate(t, 0.1, rng)[4000:8200])); ts.append(t)           # qlen 4.2 k, tlen 13 k: 5 passes
    t = rng.integers(0, 4, 12500, dtype=np.uint8)
    q = np.ascontiguousarray(np.concatenate([rng.integers(0, 4, 9000, dtype=np.uint8), mutate(t, 0.12, rng),
                                             rng.integers(0, 4, 8000, dtype=np.uint8)]))    # qlen ~29.5 k, tlen 12.5 k: 2 passes
    qs.append(q); ts.append(t)
    t = rng.integers(0, 4, 16001, dtype=np.uint8)
    qs.append(np.ascontiguousarray(mutate(t, 0.15, rng))); ts.append(t)                     # about equal lengths
    qs.append(rng.integers(0, 4, 6000, dtype=np.uint8)); ts.append(rng.integers(0, 4, 12000, dtype=np.uint8))   # unrelated
    res, cigs = ctx.ksw_batch(qs, ts, [-1] * len(qs))
    for q, t, r, cg in zip(qs, ts, res, cigs):
        exp, _ = H.oracle_ksw(q, t, -1)
        assert r.ok and [int(x) for x in cg] == exp, (len(q), len(t))


def test_band_classes_around_the_single_cta_limit(ctx):
    """Banded (not full-matrix) alignments whose band window is just below and above what one CTA holds: tier 608, tier 736
    (24 warps, 80 registers) and the 2-CTA cluster."""
    from lesv_b200.synth import mutate
    rng = np.random.default_rng(31)
    cases = []
    for w in (9500, 10400, 11700, 11900, 13000):
        q = rng.integers(0, 4, 16000, dtype=np.uint8)
        t = np.ascontiguousarray(mutate(q, 0.12, rng))
        cases.append((q, t, w))
    q = rng.integers(0, 4, 17000, dtype=np.uint8)
    t = np.ascontiguousarray(np.concatenate([q[:6000], q[9000:]]))                          # 3 kb deletion inside a 11 k band
    cases.append((q, mutate(t, 0.1, rng), 11000))
    _check(ctx, [(q, np.ascontiguousarray(t), w) for q, t, w in cases])


@pytest.mark.parametrize("switch", ["LESV_KSW_NO_GROUPS", "LESV_KSW_NO_STRIPS"])
def test_environment_switches_select_tested_paths(switch):
    """The two environment switches of the scheduler (lesvgpu.cu: narrow bands without lane groups, full-matrix
    alignments on clusters instead of column strips) select kernels that stay bit-exact: a fresh process per switch
    (they are read once), narrow-band pairs and full-matrix pairs wider than one CTA against the oracle."""
    import os, subprocess, sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests")
import helpers as H
from lesv_b200 import capi
from lesv_b200.synth import mutate
rng = np.random.default_rng(5)
cases = []
for it in range(60):                                   # bands 8 .. 200: lane groups (or the 32-thread tier)
    q = rng.integers(0, 4, int(rng.integers(200, 3000)), dtype=np.uint8)
    t = np.ascontiguousarray(mutate(q, 0.08, rng))
    cases.append((q, t, abs(len(q) - len(t)) + int(rng.integers(8, 200))))
for it in range(3):                                    # no band, 11-12 kb: wider than one CTA (strips, or a cluster)
    q = rng.integers(0, 4, 11000 + 300 * it, dtype=np.uint8)
    t = np.ascontiguousarray(mutate(q, 0.12, rng))
    cases.append((q, t, -1))
ctx = capi.Context(0)
res, cigs = ctx.ksw_batch([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases])
for (q, t, w), r, c in zip(cases, res, cigs):
    exp, _ = H.oracle_ksw(q, t, w)
    assert r.ok and [int(x) for x in c] == exp, (len(q), len(t), w)
print("ok", len(cases))
'''
    out = subprocess.run([sys.executable, "-c", code], cwd=H.ROOT, env=dict(os.environ, **{switch: "1"}), capture_output=True, text=True)
    assert out.returncode == 0 and "ok 63" in out.stdout, out.stderr[-2000:]
