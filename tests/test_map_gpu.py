"""qx2map candidate stage on the GPU through the C ABI (lesvgpu_index_build, lesvgpu_map_candidates)
against the reference: the committed golden run (tests/golden/map_small, made by make_golden_map.py
from the unmodified reference) and, when oracle/_ref travelled to the box, fresh reference runs on
seeded synthetic data.  Integer outputs: bit-exact."""
import os
import shutil

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(H.ROOT, "tests", "golden", "map_small")


@pytest.fixture(scope="module")
def ctx():
    from lesv_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _load(ctx, db_dir, k=15, w=10, max_occ=200):
    from lesv_b200 import capi
    S, Q = H.LesvDb(db_dir, "subject"), H.LesvDb(db_dir, "query")
    capi.subject_load(ctx, S.pac)
    capi.index_build(ctx, S.seq_offset, S.seq_size, k, w, max_occ)
    return S, Q


def test_index_equals_reference_lookup_table(ctx):
    from lesv_b200 import capi
    _load(ctx, os.path.join(GOLD, "db"))
    off, hashes, starts = capi.index_export(ctx)
    z = np.load(os.path.join(GOLD, "index.npz"))
    assert np.array_equal(off, z["offset_list"])            # sorted by (hash, offset), repeats removed
    assert np.array_equal(hashes, z["hashes"])
    assert np.array_equal(np.diff(starts.astype(np.int64)), z["counts"])


def test_matches_and_windows_equal_golden(ctx):
    import refmap
    from lesv_b200 import capi
    S, Q = _load(ctx, os.path.join(GOLD, "db"))
    res = capi.map_candidates(ctx, Q.pac, Q.seq_offset, Q.seq_size, 0, Q.num_seqs, 3, 5)
    m = capi.map_last_matches(ctx)
    exp_m = np.load(os.path.join(GOLD, "matches.npy"))
    assert m.shape == exp_m.shape and np.array_equal(m, exp_m)
    exp = refmap.parse_subseq(os.path.join(GOLD, "cand.subseq.txt"))
    assert sum(len(v) for v in exp.values()) > 100
    for i, got in enumerate(res):
        assert got == exp.get(i, []), (i, got, exp.get(i, []))


def test_sub_batches_of_a_volume(ctx):
    """[qfrom, qto) slices: ids stay volume-relative (on this data the batch-wide occurrence filter
    does not bite for the first reads, so the windows are those of the whole-volume run)."""
    import refmap
    from lesv_b200 import capi
    S, Q = _load(ctx, os.path.join(GOLD, "db"))
    exp = refmap.parse_subseq(os.path.join(GOLD, "cand.subseq.txt"))
    res = capi.map_candidates(ctx, Q.pac, Q.seq_offset, Q.seq_size, 7, 19, 3, 5)
    for i, got in enumerate(res):
        assert got == exp.get(7 + i, [])
    assert capi.map_candidates(ctx, Q.pac, Q.seq_offset, Q.seq_size, 4, 4, 3, 5) == []


@pytest.mark.skipif(not os.path.exists(os.path.join(H.ROOT, "oracle", "_ref", "bin", "qx2map")), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["contigs24", "long_reads", "k13w7"])
def test_candidates_against_fresh_reference_run(ctx, tmp_path, case):
    import refmap
    from lesv_b200 import capi, synth
    work = str(tmp_path)
    k, w, occ = 15, 10, 200
    extra = []
    if case == "contigs24":
        synth.make_dataset(work, ref_len=1_200_000, n_reads=120, read_len=8000, n_events=24, seed=31, n_contigs=24)
    elif case == "long_reads":
        synth.make_dataset(work, ref_len=2_000_000, n_reads=24, read_len=50000, n_events=10, seed=32, err_lo=0.10, err_hi=0.15)
    else:
        k, w, occ = 13, 7, 50
        extra = ["-kmer_size", "13", "-kmer_window", "7", "-max_kmer_occ", "50"]
        synth.make_dataset(work, ref_len=400_000, n_reads=60, read_len=5000, n_events=8, seed=33, n_contigs=2)
    refmap.run_qx2map(work, ["-outfmt", "subseq", "-max_target_seqs", "5", "-max_hsps", "10", "-db_dir", "db", "-keep_db"] + extra,
                      "cand.txt")
    exp = refmap.parse_subseq(os.path.join(work, "cand.txt"))
    db = os.path.join(work, "db")
    S, Q = _load(ctx, db, k, w, occ)
    off, hashes, starts = capi.index_export(ctx)
    r_off, r_hashes, r_cnt, handles = refmap.ref_index(db, k, w, occ)
    assert np.array_equal(off, r_off) and np.array_equal(hashes, r_hashes)
    assert np.array_equal(np.diff(starts.astype(np.int64)), r_cnt)
    res = capi.map_candidates(ctx, Q.pac, Q.seq_offset, Q.seq_size, 0, Q.num_seqs, 3, 5)
    assert np.array_equal(capi.map_last_matches(ctx), refmap.ref_matches(db, k, occ, handles))
    n = 0
    for i, got in enumerate(res):
        assert got == exp.get(i, []), (case, i, got, exp.get(i, []))
        n += len(got)
    assert n >= Q.num_seqs // 2
    shutil.rmtree(work, ignore_errors=True)


# ---- full mapping (lesvgpu_map_batch): HSPs and CIGARs -----------------------------------------------
def m4_lines(res, Q, S):
    """Format HSP dicts the way print_one_m4_result does with -cigar (mecat_results.c:27-62, 256-280;
    note the role swap of print_cigar_core: a gap in the reference string prints as D)."""
    lines = []
    for i, lst in enumerate(res):
        for h in lst:
            qoff, qend = h["qoff"], h["qend"]
            if h["qdir"] == 1:                      # dump_m4_hits, chain_and_extend_kmer_matches.c:325-330
                qoff, qend = h["qsize"] - h["qend"], h["qsize"] - h["qoff"]
            ident = 100.0 * h["n_match"] / h["aln_len"]
            score = 8 * h["aln_len"] - h["n_match"]
            ops, cg, a = h["ops"], [], 0
            while a < len(ops):
                b = a
                while b < len(ops) and ops[b] == ops[a]:
                    b += 1
                cg.append(f"{b - a}{'MDI'[ops[a]]}")
                a = b
            lines.append(f"{Q.name(h['qid'])}\t{S.name(h['sid'])}\t{ident:g}\t{score}\t{h['qdir']}\t{qoff}\t{qend}\t{h['qsize']}"
                         f"\t0\t{h['soff']}\t{h['send']}\t{h['ssize']}\tcg:Z:{''.join(cg)}\n")
    return sorted(lines)


def test_map_batch_equals_golden_m4_with_cigars(ctx):
    from lesv_b200 import capi
    S, Q = _load(ctx, os.path.join(GOLD, "db"))
    res = capi.map_batch(ctx, Q.pac, Q.seq_offset, Q.seq_size, 0, Q.num_seqs, hitlist_size=5, max_hsps=10, want_ops=True)
    exp = open(os.path.join(GOLD, "map.cigar.m4")).readlines()
    got = m4_lines(res, Q, S)
    assert len(exp) > 150
    assert got == exp
    tm = capi.map_last_extend_timing(ctx)
    assert tm.n_unsupported == 0


@pytest.mark.skipif(not os.path.exists(os.path.join(H.ROOT, "oracle", "_ref", "bin", "qx2map")), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", ["reads10k", "contigs5_svs", "skip_overhang", "noisy_tails"])
def test_map_batch_against_fresh_reference_run(ctx, tmp_path, case):
    import refmap
    from lesv_b200 import capi, synth
    work = str(tmp_path)
    extra = []
    if case == "reads10k":
        synth.make_dataset(work, ref_len=1_000_000, n_reads=60, read_len=10000, n_events=20, seed=41)
    elif case == "contigs5_svs":
        synth.make_dataset(work, ref_len=1_500_000, n_reads=50, read_len=20000, n_events=30, seed=42, n_contigs=5,
                           err_lo=0.08, err_hi=0.16)
    elif case == "noisy_tails":
        # junk / 30-45 % error read ends: the overhang step of hbn_traceback (left_extend / right_extend with the
        # DALIGNER wave, hbn_traceback.c:170-311) must run on the GPU and reproduce the reference's HSP ends and CIGARs
        ref = synth.random_codes(800_000, np.random.default_rng(44))
        reads = synth.reads_with_noisy_tails(ref, 60, np.random.default_rng(45))
        synth.write_fasta(os.path.join(work, "ref.fa"), ["chrT"], [ref], width=0)
        synth.write_fasta(os.path.join(work, "reads.fa"), [f"r{i}" for i in range(len(reads))], reads)
    else:
        extra = ["-skip_overhang"]
        synth.make_dataset(work, ref_len=600_000, n_reads=40, read_len=6000, n_events=10, seed=43, n_contigs=2)
    refmap.run_qx2map(work, ["-cigar", "-max_target_seqs", "5", "-max_hsps", "10", "-db_dir", "db", "-keep_db"] + extra, "map.m4")
    exp = sorted(open(os.path.join(work, "map.m4")).readlines())
    S, Q = _load(ctx, os.path.join(work, "db"))
    _bind = capi._bind_map(ctx.lib)
    opts_skip = case == "skip_overhang"
    res = capi.map_batch(ctx, Q.pac, Q.seq_offset, Q.seq_size, 0, Q.num_seqs, hitlist_size=5, max_hsps=10, want_ops=True,
                         skip_overhang=opts_skip)
    got = m4_lines(res, Q, S)
    assert len(exp) >= Q.num_seqs // 2
    assert got == exp
    n_wave = sum(1 for lst in res for h in lst if h["flags"] & 3)       # lesvgpu_hsp.flags: the wave ran left (1) / right (2)
    if case == "noisy_tails":
        assert n_wave >= 10, n_wave
        assert any(h["flags"] & 1 for lst in res for h in lst) and any(h["flags"] & 2 for lst in res for h in lst)
    if case == "skip_overhang":
        assert n_wave == 0
    assert capi.map_last_extend_timing(ctx).n_unsupported == 0
    shutil.rmtree(work, ignore_errors=True)


# ---- warp-cooperative edlib_nw (trace_warp.cuh) against the sequential restatement (validated against edlib on CPU) ----
def test_warp_edlib_nw_equals_sequential_restatement(ctx):
    import ctypes as C
    from lesv_b200.synth import mutate
    lib = ctx.lib
    lib.lesvgpu_debug_edlib_nw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    lib.lesvgpu_debug_edlib_nw_block.argtypes = lib.lesvgpu_debug_edlib_nw.argtypes
    em = H.trace_emul()
    rng = np.random.default_rng(21)
    sizes = [1, 2, 63, 64, 65, 127, 128, 129, 500, 1000, 2047, 2048, 2049, 2100, 3000, 4100, 5000, 9000, 17000]
    n_h = 0
    for it, n in enumerate(sizes * 2):
        two = it % 5 == 3                                   # two-letter alphabet: many co-optimal paths and splits
        q = rng.integers(0, 2 if two else 4, n, dtype=np.uint8)
        t = np.ascontiguousarray(mutate(q, float(rng.uniform(0.03, 0.3)), rng))
        if two:
            t &= 1
        if it % 4 == 1 and n > 200:                         # a large deletion / unrelated insert inside
            k = n // 3
            t = np.ascontiguousarray(np.concatenate([t[:k], rng.integers(0, 4, n // 4, dtype=np.uint8), t[k + n // 5:]]))
        if len(t) == 0:
            t = q[:1].copy()
        n_h += 20 * ((len(q) + 63) // 64) * len(t) + 8 * len(t) >= 1 << 20
        exp = np.zeros(len(q) + len(t) + 8, dtype=np.uint8)
        ne = em.lesv_emul_nw(H._ptr(q), len(q), H._ptr(t), len(t), H._ptr(exp), len(exp))
        got = np.zeros(len(q) + len(t) + 8, dtype=np.uint8)
        ng = C.c_int()
        assert lib.lesvgpu_debug_edlib_nw(ctx.h, H._ptr(q), len(q), H._ptr(t), len(t), H._ptr(got), C.byref(ng)) == 0
        assert ng.value == ne and np.array_equal(got[:ne], exp[:ne]), (it, len(q), len(t))
        got[:] = 0
        assert lib.lesvgpu_debug_edlib_nw_block(ctx.h, H._ptr(q), len(q), H._ptr(t), len(t), H._ptr(got), C.byref(ng)) == 0
        assert ng.value == ne and np.array_equal(got[:ne], exp[:ne]), ("block", it, len(q), len(t))
    assert n_h >= 8                                          # Hirschberg route exercised
