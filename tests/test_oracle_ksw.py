"""The oracle restatement of ksw_extd2_sse against the compiled, unmodified reference."""
import numpy as np
import pytest

import helpers as H

needs_ref = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (make -C oracle ref)")


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_reference_cigar_and_score(seed):
    rng = np.random.default_rng(seed)
    for it in range(1500):
        q, t, w = H.random_pair(rng, it)
        a = H.ref_ksw(q, t, w)
        b = H.oracle_ksw(q, t, w)
        assert a == b, (it, len(q), len(t), w, H.cigar_str(a[0])[:60], H.cigar_str(b[0])[:60])


@needs_ref
def test_oracle_matches_reference_longer_reads():
    rng = np.random.default_rng(77)
    from lesv_b200.synth import mutate
    for it in range(6):
        q = rng.integers(0, 4, 6000, dtype=np.uint8)
        t = mutate(q, 0.12, rng)
        w = int((len(q) + len(t)) * 12 / 200 * 1.2)
        assert H.ref_ksw(q, t, w) == H.oracle_ksw(q, t, w)


def test_cells_closed_form_and_cut_detection():
    """|qlen-tlen| <= w  <=>  the band reaches the last cell (what the host library assumes)."""
    rng = np.random.default_rng(5)
    for _ in range(3000):
        ql, tl = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        w = int(rng.integers(0, 320))
        cells, ok = H.oracle_cells(ql, tl, w)
        expect = abs(ql - tl) <= w and (w >= 1 or ql + tl - 1 <= 1)
        assert ok == expect, (ql, tl, w)
