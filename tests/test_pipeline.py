"""The drop-in tools chained the way lesv.sh chains the reference's (src/scripts/lesv.sh:105-172):
    qx2map -> qx2m4x -> qx2svr -> qx2svsig
all four from lesv_b200/bin, against the committed golden run of the unmodified reference
(tests/golden/svsig_small) and, when oracle/_ref travelled to the box, a fresh reference run with several
query volumes and contigs.  qx2m4x alone (host I/O only) is checked without a GPU."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

GOLD = os.path.join(H.ROOT, "tests", "golden", "svsig_small")
BIN = os.path.join(H.ROOT, "lesv_b200", "bin")
REF_BIN = os.path.join(H.ROOT, "oracle", "_ref", "bin")


def _m4x(path):
    """M4Record[] without the 4 trailing padding bytes (the reference writes them uninitialised, m4x.cpp:53-58)."""
    raw = np.fromfile(path, dtype=np.uint8).reshape(-1, 56)[:, :52]
    return sorted(bytes(r) for r in raw)


def _lines(path):
    with open(path) as f:
        return f.readlines()


def test_qx2m4x_equals_reference_records(tmp_path):
    subprocess.run(["make", "-s", "-C", os.path.join(H.ROOT, "lesv_b200", "csrc"), "../bin/qx2m4x"], check=True)
    db = str(tmp_path / "db")
    shutil.copytree(os.path.join(GOLD, "db"), db)
    br = os.path.join(db, "backup_results")
    name = "stagebackup_results_Q00000000_D00000000"
    os.rename(os.path.join(br, name + ".m4x"), os.path.join(br, "expected.m4x"))
    shutil.copy(os.path.join(GOLD, "map.m4"), os.path.join(br, name))
    subprocess.run([os.path.join(BIN, "qx2m4x"), db], check=True, stderr=subprocess.DEVNULL)
    got, exp = _m4x(os.path.join(br, name + ".m4x")), _m4x(os.path.join(br, "expected.m4x"))
    assert len(exp) == 60 and got == exp
    # a 13-column (-cigar) file must be refused like the reference's hbn_assert does (m4x.cpp:18)
    with open(os.path.join(br, name), "a") as f:
        f.write(_lines(os.path.join(GOLD, "map.m4"))[0].rstrip("\n") + "\tcg:Z:10M\n")
    assert subprocess.run([os.path.join(BIN, "qx2m4x"), db], stderr=subprocess.DEVNULL).returncode != 0


def _run_ours(work, map_args, threads="4"):
    env = dict(os.environ)
    with open(os.path.join(work, "gpu.map.m4"), "wb") as f, open(os.path.join(work, "gpu.log"), "ab") as log:
        subprocess.run([os.path.join(BIN, "qx2map")] + map_args + ["-db_dir", "db_gpu", "-keep_db", "-num_threads", threads,
                        "split.fa", "ref.fa"], cwd=work, check=True, stdout=f, stderr=log, env=env)
        subprocess.run([os.path.join(BIN, "qx2m4x"), "db_gpu"], cwd=work, check=True, stderr=log)
        subprocess.run([os.path.join(BIN, "qx2svr"), "svr_gpu", "db_gpu", "3000", "70.0", "300", threads], cwd=work, check=True, stderr=log)
        subprocess.run([os.path.join(BIN, "qx2svsig"), "svsig_gpu", "svr_gpu", "db_gpu", "40", threads], cwd=work, check=True, stderr=log)


def _compare_pipeline(work, ref_db, ref_m4, ref_svr, ref_svsig):
    assert sorted(_lines(os.path.join(work, "gpu.map.m4"))) == sorted(_lines(ref_m4))
    n = 0
    br = os.path.join(ref_db, "backup_results")
    for f in sorted(os.listdir(br)):
        if f.endswith(".m4x"):
            exp = _m4x(os.path.join(br, f))
            assert _m4x(os.path.join(work, "db_gpu", "backup_results", f)) == exp, f
            n += len(exp)
    for f in sorted(os.listdir(ref_svr)):
        if f.endswith(".sv_read"):
            a, b = _lines(os.path.join(ref_svr, f)), _lines(os.path.join(work, "svr_gpu", f))
            assert (a == b) if f.startswith("subject_") else (sorted(a) == sorted(b)), f
    n_sig = 0
    for f in sorted(os.listdir(ref_svsig)):
        if f.endswith(".sv_signature"):
            ie, de = H.load_sv_signatures(os.path.join(ref_svsig, f))
            ig, dg = H.load_sv_signatures(os.path.join(work, "svsig_gpu", f))
            assert ie == ig and de == dg, f
            n_sig += len(ie) + len(de)
    return n, n_sig


@pytest.mark.gpu
def test_four_tools_reproduce_the_golden_run(tmp_path):
    work = str(tmp_path)
    shutil.copy(os.path.join(GOLD, "reads.fa"), os.path.join(work, "split.fa"))      # 4 kb reads: qx2splitseq is the identity
    shutil.copy(os.path.join(GOLD, "ref.fa"), os.path.join(work, "ref.fa"))
    _run_ours(work, ["-max_target_seqs", "5", "-max_hsps", "10"])
    n, n_sig = _compare_pipeline(work, os.path.join(GOLD, "db"), os.path.join(GOLD, "map.m4"), os.path.join(GOLD, "svr"),
                                 os.path.join(GOLD, "svsig"))
    assert n == 60 and n_sig >= 5
    # resume: a second run finds the .mapped sentinels and only re-emits the backup files (hbn_job_control.c:31-92)
    first = _lines(os.path.join(work, "gpu.map.m4"))
    with open(os.path.join(work, "again.m4"), "wb") as f:
        subprocess.run([os.path.join(BIN, "qx2map"), "-max_target_seqs", "5", "-max_hsps", "10", "-db_dir", "db_gpu", "-keep_db",
                        "split.fa", "ref.fa"], cwd=work, check=True, stdout=f, stderr=subprocess.DEVNULL)
    assert _lines(os.path.join(work, "again.m4")) == first


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BIN, "qx2map")), reason="oracle/_ref not built")
def test_four_tools_against_fresh_reference_run_with_volumes(tmp_path):
    """3 contigs, 150 x 12 kb reads at 10-15 %, planted 40 bp-3 kb INS/DEL, query database cut into volumes of
    <= 400 kb and mapped in batches of 150 kb (the -max_query_vol_res / -query_batch_size surface)."""
    sys.path.insert(0, os.path.join(H.ROOT, "tests", "golden"))
    from lesv_b200 import synth
    work = str(tmp_path)
    synth.make_dataset(work, ref_len=900_000, n_reads=150, read_len=12000, n_events=18, seed=77, n_contigs=3,
                       err_lo=0.10, err_hi=0.15, event_sizes=(40, 60, 100, 300, 1000, 3000))
    shutil.copy(os.path.join(work, "reads.fa"), os.path.join(work, "split.fa"))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    args = ["-max_target_seqs", "5", "-max_hsps", "10", "-max_query_vol_res", "400k", "-query_batch_size", "150k"]
    t = str(os.cpu_count() or 4)
    with open(os.path.join(work, "ref.map.m4"), "wb") as f, open(os.path.join(work, "ref.log"), "ab") as log:
        subprocess.run([os.path.join(REF_BIN, "qx2map")] + args + ["-db_dir", "db", "-keep_db", "-num_threads", "1", "split.fa", "ref.fa"],
                       cwd=work, check=True, stdout=f, stderr=log, env=env)
        subprocess.run([os.path.join(REF_BIN, "qx2m4x"), "db"], cwd=work, check=True, stdout=log, stderr=log, env=env)
        subprocess.run([os.path.join(REF_BIN, "qx2svr"), "svr", "db", "3000", "70.0", "300", t], cwd=work, check=True, stdout=log, stderr=log, env=env)
        subprocess.run([os.path.join(REF_BIN, "qx2svsig"), "svsig", "svr", "db", "40", t], cwd=work, check=True, stdout=log, stderr=log, env=env)
    _run_ours(work, args)
    n, n_sig = _compare_pipeline(work, os.path.join(work, "db"), os.path.join(work, "ref.map.m4"), os.path.join(work, "svr"),
                                 os.path.join(work, "svsig"))
    assert n >= 140 and n_sig >= 10
    assert len([f for f in os.listdir(os.path.join(work, "db_gpu", "backup_results")) if f.endswith(".m4x")]) >= 4


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BIN, "qx2map")), reason="oracle/_ref not built")
@pytest.mark.parametrize("scenario", ["c1", "k19", "repeats", "c3"])
def test_four_tools_on_named_configs(tmp_path, scenario):
    """tools/parity_probe.py runs both tool chains and diffs every output file:
    c1      = BASELINE.json configs[0] at full size (1 Mbp reference, 2,000 x 10 kb reads at 12 %, 20 planted INS/DEL);
    k19     = the README's ultra-long setting -kmer_size 19 -kmer_window 20 on 30 kb reads at 10-25 % error
              (the int-truncated hash path of hbn_word_finder.c:263);
    repeats = tandem repeats and dispersed 6 kb duplicates (ties in the reference's unstable sorts, Appendix A #1-#4);
    c3      = BASELINE.json configs[2] shape at 1/10 scale: 24 contigs with hg19's length profile (310 Mbp), 50 kb reads
              with 40 bp-3 kb INS/DEL (the full 3.1 Gbp run is recorded in profiles/r1_c3_full_scale.txt)."""
    env = dict(os.environ, C3_SCALE="0.1", C3_READS="48", REF_MAP_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "tools", "parity_probe.py"), scenario, str(tmp_path)],
                         check=True, capture_output=True, text=True, env=env).stdout
    assert "m4: reference" in out and ", identical True" in out, out
    assert "svr:" in out and "svsig:" in out and out.count("all files identical: True") == 2, out
    n_m4 = int(out.split("m4: reference ")[1].split(" ")[0])
    assert n_m4 >= {"c1": 2000, "k19": 100, "repeats": 300, "c3": 40}[scenario]


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BIN, "qx2map")), reason="oracle/_ref not built")
@pytest.mark.parametrize("scenario", ["bigsv", "trf", "minindel", "tails", "c4"])
def test_four_tools_on_reference_branches(tmp_path, scenario):
    """More of tools/parity_probe.py, one reference branch each:
    bigsv    = 50 kb reads across 6-10 kb INS/DEL: qx2svsig bands of 10-20 k, the 2-CTA-cluster launch class at real size;
    trf      = a TRF bed file as 8th / 7th argument of qx2svr / qx2svsig (trf_array.cpp:10-83);
    minindel = qx2svsig with min_indel_size 1: hundreds of records per read (more than the kernel keeps per alignment);
    tails    = junk and 30-45 % error read ends: the overhang step of hbn_traceback (DALIGNER wave);
    c4       = BASELINE.json configs[3] shape: 150 kb+ reads -> qx2splitseq 50000 0 20000 -> -kmer_size 19 -kmer_window 20,
               a third of the reads at 20-30 % error, SVR identity 80.0 (README.md:164-172)."""
    env = dict(os.environ, BIGSV_READS="16", C4_READS="8", C4_REF_MBP="8", REF_MAP_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "tools", "parity_probe.py"), scenario, str(tmp_path)],
                         check=True, capture_output=True, text=True, env=env).stdout
    assert "m4: reference" in out and ", identical True" in out, out
    assert "svr:" in out and "svsig:" in out and out.count("all files identical: True") == 2, out
    import re
    n_m4 = int(out.split("m4: reference ")[1].split(" ")[0])
    n_svr = int(re.search(r"^svr: (\d+) records", out, re.M).group(1))
    n_sig = int(re.search(r"^svsig: (\d+) records", out, re.M).group(1))
    assert n_m4 >= {"bigsv": 16, "trf": 90, "minindel": 90, "tails": 70, "c4": 25}[scenario], out
    if scenario == "bigsv":
        assert n_sig >= 8, out                     # the planted 6-10 kb events are found
    if scenario == "trf":
        assert 40 <= n_svr <= 140, out             # 90 reads without the mask = 180 lines (Q + subject files)
    if scenario == "minindel":
        assert n_sig > 64 * 90, out
    if scenario == "c4":
        assert "splitseq: identical True" in out, out


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BIN, "qx2map")), reason="oracle/_ref not built")
@pytest.mark.parametrize("scenario", ["m4x", "md"])
def test_qx2map_output_formats(tmp_path, scenario):
    """qx2map -outfmt m4x (binary M4Record stream on stdout and in the backup file, mecat_results.c:281-297) and
    -cigar -md (MD:Z: column, mecat_results.c:86-135) against the reference's output."""
    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "tools", "parity_probe.py"), scenario, str(tmp_path)],
                         check=True, capture_output=True, text=True).stdout
    if scenario == "m4x":
        assert "identical True; backup files identical True" in out, out
        assert int(out.split("m4x: reference ")[1].split(" ")[0]) >= 90, out
    else:
        assert "identical True" in out and int(out.split("lines (")[1].split(" ")[0]) >= 90, out


def _tree(work, dirs):
    out = {}
    for d in dirs:
        for f in sorted(os.listdir(os.path.join(work, d))):
            path = os.path.join(work, d, f)
            if os.path.isfile(path):
                out[d + "/" + f] = open(path, "rb").read()
    return out


@pytest.mark.gpu
def test_tools_write_the_same_files_for_any_number_of_feeders(tmp_path):
    """One host process, one feeder thread per GPU pulling sub-batches from a shared queue (lesv_b200/host/lesv_gpus.hpp).
    LESV_GPUS=0 against LESV_GPUS=0,0,0 (three contexts and feeder threads; on a multi-GPU box also 0,1) with small
    sub-batches, so that every tool really has several queue items: all output files must be byte-identical, and
    equal to the golden run of the reference."""
    import ctypes as C
    from lesv_b200 import capi
    n_dev = capi.load().lesvgpu_device_count()
    lists = ["0", "0,0,0"] + (["0,1"] if n_dev >= 2 else [])
    trees = []
    for gi, gl in enumerate(lists):
        work = str(tmp_path / f"g{gi}")
        os.makedirs(work)
        shutil.copy(os.path.join(GOLD, "reads.fa"), os.path.join(work, "split.fa"))
        shutil.copy(os.path.join(GOLD, "ref.fa"), os.path.join(work, "ref.fa"))
        env = dict(os.environ, LESV_GPUS=gl, LESV_SVSIG_BATCH="7", LESV_SVR_CHUNK="5")
        env.pop("LESV_GPU", None)
        with open(os.path.join(work, "gpu.map.m4"), "wb") as f, open(os.path.join(work, "gpu.log"), "ab") as log:
            for cmd, out in (([os.path.join(BIN, "qx2map"), "-max_target_seqs", "5", "-max_hsps", "10", "-query_batch_size", "20k",
                               "-db_dir", "db_gpu", "-keep_db", "split.fa", "ref.fa"], f),
                             ([os.path.join(BIN, "qx2m4x"), "db_gpu"], None),
                             ([os.path.join(BIN, "qx2svr"), "svr_gpu", "db_gpu", "3000", "70.0", "300", "4"], None),
                             ([os.path.join(BIN, "qx2svsig"), "svsig_gpu", "svr_gpu", "db_gpu", "40", "4"], None)):
                rc = subprocess.run(cmd, cwd=work, stdout=out, stderr=log, env=env).returncode
                log.flush()
                assert rc == 0, (gl, cmd[0], rc, open(os.path.join(work, "gpu.log"), errors="replace").read()[-1500:])
        log = open(os.path.join(work, "gpu.log"), errors="replace").read()
        assert f"on {len(gl.split(','))} GPU(s)" in log, log
        t = _tree(work, ["svr_gpu", "svsig_gpu", "db_gpu/backup_results"])
        t["map.m4"] = open(os.path.join(work, "gpu.map.m4"), "rb").read()
        trees.append(t)
        if gi == 0:            # against the reference's golden files (query batches of 20 kb: same reads per batch for every GPU count)
            n, n_sig = _compare_pipeline(work, os.path.join(GOLD, "db"), os.path.join(GOLD, "map.m4"), os.path.join(GOLD, "svr"),
                                         os.path.join(GOLD, "svsig"))
            assert n == 60 and n_sig >= 5
    for t in trees[1:]:
        assert t.keys() == trees[0].keys()
        for k in t:
            assert t[k] == trees[0][k], k


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BIN, "qx2asvr")), reason="oracle/_ref not built")
def test_qx2asvr_equals_reference_sam(tmp_path):
    """consensus reads of SV groups against their reference windows (map_cns_sv_read.c:170-292), SAM
    records with CIGAR / MD / identity tags (map_results.c:63-323).  Both tools run on the same
    subject_%08d.normal.cns.fasta files; the reference writes records in thread order, so records compare as multisets;
    the header is identical apart from the @PG line, which carries the program path."""
    from lesv_b200 import synth
    work = str(tmp_path)
    rng = np.random.default_rng(3)
    contigs = [synth.random_codes(400_000, rng) for _ in range(3)]
    synth.write_fasta(os.path.join(work, "ref.fa"), [f"chr{i}" for i in range(3)], contigs, width=0)
    synth.write_fasta(os.path.join(work, "dummy.fa"), ["q0"], [contigs[0][:5000]])
    env = dict(os.environ, OMP_NUM_THREADS="1")
    subprocess.run([os.path.join(REF_BIN, "qx2map"), "-db_dir", "db", "-keep_db", "-num_threads", "1", "dummy.fa", "ref.fa"], cwd=work,
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
    per = synth.consensus_reads(contigs, 60, np.random.default_rng(4), max_len=52000)      # up to 52 kb: band 0.2 x len > one CTA
    for d in ("grp_ref", "grp_gpu"):
        os.makedirs(os.path.join(work, d))
        for sid, lst in enumerate(per):
            synth.write_fasta(os.path.join(work, d, f"subject_{sid:08d}.normal.cns.fasta"), [n for n, _ in lst], [c for _, c in lst], width=0)
    subprocess.run([os.path.join(REF_BIN, "qx2asvr"), "db", "grp_ref", str(os.cpu_count() or 4)], cwd=work, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
    subprocess.run([os.path.join(BIN, "qx2asvr"), "db", "grp_gpu", "4"], cwd=work, check=True, stderr=subprocess.DEVNULL,
                   env=dict(os.environ, LESV_ASVR_BATCH="9"))
    n_rec = 0
    for sid in range(3):
        name = f"subject_{sid:08d}.normal.cns.fasta"
        a, b = _lines(os.path.join(work, "grp_ref", name + ".sam")), _lines(os.path.join(work, "grp_gpu", name + ".sam"))
        head = lambda x: [l for l in x if l.startswith("@") and not l.startswith("@PG")]
        body = lambda x: sorted(l for l in x if not l.startswith("@"))
        assert head(a) == head(b) and len(head(a)) == 5
        assert body(a) == body(b), name
        n_rec += len(body(a))
        assert os.path.exists(os.path.join(work, "grp_gpu", name + ".mapped"))
        for f in ("query.pac", "query.seq_info", "query.hdr"):                      # the packed consensus reads, byte for byte
            pa, pb = (os.path.join(work, d, f"s{sid:08d}_svr_pdb", f) for d in ("grp_ref", "grp_gpu"))
            assert open(pa, "rb").read() == open(pb, "rb").read(), f
    assert 30 <= n_rec < 60            # the reads at 10-22 % error fall below 85 % effective identity and are dropped
