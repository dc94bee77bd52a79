"""qx2svsig on the GPU: the C ABI (lesvgpu_svsig_batch) and the drop-in binary against the
reference's own output files (golden fixtures + a fresh run of oracle/_ref/bin when present)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(H.ROOT, "tests", "golden", "svsig_small")
BIN = os.path.join(H.ROOT, "lesv_b200", "bin", "qx2svsig")
REF_BIN = os.path.join(H.ROOT, "oracle", "_ref", "bin")


@pytest.fixture(scope="module")
def ctx():
    from lesv_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def test_svsig_batch_equals_golden_and_oracle(ctx):
    from lesv_b200 import capi
    q, s = H.LesvDb(os.path.join(GOLD, "db"), "query"), H.LesvDb(os.path.join(GOLD, "db"), "subject")
    reads = H.load_sv_reads(os.path.join(GOLD, "svr", "subject_00000000.sv_read"))
    capi.subject_load(ctx, s.pac)
    per, sigs = capi.svsig_batch(ctx, q.pac, q.read_offset, int(s.seq_offset[0]), int(s.seq_size[0]), reads, 40)
    ins, dele = H.load_sv_signatures(os.path.join(GOLD, "svsig", "subject_00000000.sv_signature"))
    got = [t + (q.name(t[0]),) for t in sigs]
    assert sorted(x for x in got if x[12] == 6) == ins
    assert sorted(x for x in got if x[12] == 0) == dele
    contig = s.codes(0)
    k = 0
    for i, r in enumerate(reads):          # per read: same count, same order as the oracle's forward scan
        n, exp = H.oracle_discover(r, q.codes(r[0]), contig, 40)
        assert (per[i].n_sig if per[i].status == 1 else -1) == n, (i, per[i].status)
        if n > 0:
            assert sigs[k:k + n] == exp
            k += n


def _compare_dirs(ref_dir, got_dir):
    names = sorted(f for f in os.listdir(ref_dir) if f.endswith(".sv_signature"))
    assert names
    total = 0
    for f in names:
        a, b = H.load_sv_signatures(os.path.join(ref_dir, f)), H.load_sv_signatures(os.path.join(got_dir, f))
        assert a == b, f
        total += len(a[0]) + len(a[1])
    return total


def test_binary_dropin_on_golden(tmp_path):
    out = tmp_path / "svsig"
    subprocess.run([BIN, str(out), os.path.join(GOLD, "svr"), os.path.join(GOLD, "db"), "40", "4"], check=True)
    assert _compare_dirs(os.path.join(GOLD, "svsig"), str(out)) >= 5
    assert os.path.exists(out / "subject_00000000.sv_signature.created")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BIN, "qx2svsig")), reason="oracle/_ref not built")
def test_binary_dropin_against_fresh_reference_run(tmp_path):
    """Config-1 shape, reduced: 400 kb reference, 2 contigs, 240 x 8 kb reads at 12 %, planted INS/DEL."""
    sys.path.insert(0, os.path.join(H.ROOT, "tests", "golden"))
    import make_golden
    from lesv_b200 import synth
    work = str(tmp_path / "w")
    synth.make_dataset(work, ref_len=400000, n_reads=240, read_len=8000, n_events=8, seed=5, n_contigs=2,
                       event_sizes=(40, 60, 100, 300, 1000, 3000))
    make_golden.reference_pipeline(work, threads=os.cpu_count() or 4)
    out = os.path.join(work, "svsig_gpu")
    subprocess.run([BIN, out, os.path.join(work, "svr"), os.path.join(work, "db"), "40", "4"], check=True)
    n = _compare_dirs(os.path.join(work, "svsig"), out)
    assert n >= 10
