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


# ---- qx2svr ------------------------------------------------------------------------------------
SVR_BIN = os.path.join(H.ROOT, "lesv_b200", "bin", "qx2svr")


def _read_lines(path):
    with open(path) as f:
        return [l.rstrip("\n") for l in f]


def _compare_svr(ref_dir, got_dir):
    n = 0
    for f in sorted(os.listdir(ref_dir)):
        if f.startswith("Q") and f.endswith(".sv_read"):            # thread order in the reference: multiset
            assert sorted(_read_lines(os.path.join(ref_dir, f))) == sorted(_read_lines(os.path.join(got_dir, f))), f
            n += len(_read_lines(os.path.join(ref_dir, f)))
        elif f.startswith("subject_") and f.endswith(".sv_read"):    # ordered by query id: exact
            assert _read_lines(os.path.join(ref_dir, f)) == _read_lines(os.path.join(got_dir, f)), f
    assert _read_lines(os.path.join(ref_dir, "subject_sv_read_file_count")) == \
        _read_lines(os.path.join(got_dir, "subject_sv_read_file_count"))
    return n


def test_svr_binary_dropin_on_golden(tmp_path):
    out = tmp_path / "svr"
    subprocess.run([SVR_BIN, str(out), os.path.join(GOLD, "db"), "3000", "70.0", "300", "4"], check=True)
    assert _compare_svr(os.path.join(GOLD, "svr"), str(out)) >= 50


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BIN, "qx2svr")), reason="oracle/_ref not built")
def test_svr_and_svsig_dropin_on_repeats_and_large_svs(tmp_path):
    """A reference with a diverged 20 kb duplication (several complete HSPs per read -> the ksw
    re-scoring branch) and 2-6 kb deletions/insertions (split HSPs -> the dual-chaining branch)."""
    sys.path.insert(0, os.path.join(H.ROOT, "tests", "golden"))
    import make_golden
    from lesv_b200 import synth
    rng = np.random.default_rng(99)
    ref = synth.random_codes(300000, rng)
    dup = synth.mutate(ref[50000:70000], 0.03, rng)
    ref[200000:200000 + len(dup)] = dup
    sample = ref.copy()
    pieces, prev = [], 0
    for k, pos in enumerate(range(30000, 290000, 37000)):
        size = int(rng.integers(2000, 6000))
        pieces.append(sample[prev:pos])
        if k % 2 == 0:
            pieces.append(synth.random_codes(size, rng)); prev = pos
        else:
            prev = pos + size
    pieces.append(sample[prev:])
    sample = np.concatenate(pieces)
    reads = synth.simulate_reads(sample, 260, 9000, np.random.default_rng(7), 0.10, 0.14)
    work = str(tmp_path / "w")
    os.makedirs(work)
    synth.write_fasta(os.path.join(work, "ref.fa"), ["chrA"], [ref], width=80)
    synth.write_fasta(os.path.join(work, "reads.fa"), [f"r{i}" for i in range(len(reads))], reads)
    make_golden.reference_pipeline(work, threads=os.cpu_count() or 4)
    svr_out = os.path.join(work, "svr_gpu")
    subprocess.run([SVR_BIN, svr_out, os.path.join(work, "db"), "3000", "70.0", "300", "4"], check=True)
    assert _compare_svr(os.path.join(work, "svr"), svr_out) >= 100
    sig_out = os.path.join(work, "svsig_gpu")
    subprocess.run([BIN, sig_out, svr_out, os.path.join(work, "db"), "40", "4"], check=True)
    assert _compare_dirs(os.path.join(work, "svsig"), sig_out) >= 20


def test_svr_batch_c_abi_equals_golden(ctx):
    """lesvgpu_svr_batch (process_one_m4_list over one query volume) through the C ABI on the golden M4Record file."""
    from lesv_b200 import capi
    d = GOLD
    Q, S = H.LesvDb(os.path.join(d, "db"), "query"), H.LesvDb(os.path.join(d, "db"), "subject")
    capi.subject_load(ctx, S.pac)
    m4 = np.fromfile(os.path.join(d, "db", "backup_results", "stagebackup_results_Q00000000_D00000000.m4x"), dtype=capi.M4_DTYPE)
    svr, st = capi.svr_batch(ctx, Q.pac, Q.read_offset, S.seq_offset, S.seq_size, m4, 3000, 70.0, 300)
    got = sorted(tuple(int(r[n]) for n in ("query_id", "qdir", "qoff", "qend", "qsize", "subject_id", "soff", "send", "dist")) for r in svr)
    exp = sorted(tuple(int(x) for x in l.split()[:9]) for l in open(os.path.join(d, "svr", "Q00000000.sv_read")))
    assert len(exp) >= 50 and got == exp
    assert st.n_m4 == len(m4)
