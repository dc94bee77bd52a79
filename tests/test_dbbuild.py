"""The packed-database builder of the GPU qx2map driver (lesv_b200/host/lesv_dbbuild.hpp) against the
reference's build_db (src/corelib/build_db.c): every file byte-identical, including ambiguous runs,
lower case, gzip input and multi-volume splitting."""
import filecmp
import gzip
import os
import subprocess

import numpy as np
import pytest

import helpers as H

REFBIN = os.path.join(H.ROOT, "oracle", "_ref", "bin", "qx2map")
pytestmark = pytest.mark.skipif(not os.path.exists(REFBIN), reason="oracle/_ref not built")


def test_database_files_equal_reference(tmp_path):
    rng = np.random.default_rng(3)
    work = str(tmp_path)

    def seq(n):
        s = rng.integers(0, 4, n)
        return "".join("ACGT"[x] for x in s)

    reads = []
    for i in range(40):
        s = seq(int(rng.integers(600, 4000)))
        if i % 5 == 1:
            s = s[:100] + "N" * int(rng.integers(1, 30)) + s[100:300] + "RRYYN" + s[300:]
        if i % 5 == 2:
            s = s.lower()
        reads.append((f"read{i} some description {i}", s))
    with gzip.open(os.path.join(work, "reads.fa.gz"), "wt") as f:
        for n, s in reads:
            f.write(f">{n}\n")
            for k in range(0, len(s), 70):
                f.write(s[k:k + 70] + "\n")
    with open(os.path.join(work, "ref.fa"), "w") as f:
        for c in range(3):
            f.write(f">chr{c}\n{seq(30000 + c)}\n")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    with open(os.path.join(work, "out.m4"), "wb") as out, open(os.path.join(work, "log"), "wb") as log:
        subprocess.run([REFBIN, "-db_dir", "db_ref", "-keep_db", "-max_query_vol_res", "20k", "-max_subject_vol_res", "50k",
                        "-num_threads", "1", "reads.fa.gz", "ref.fa"], cwd=work, check=True, stdout=out, stderr=log, env=env)
    subprocess.run(["make", "-s", "-C", os.path.join(H.ROOT, "tests", "emul"), "dbbuild_tool"], check=True)
    tool = os.path.join(H.ROOT, "tests", "emul", "dbbuild_tool")
    os.makedirs(os.path.join(work, "db_ours"))
    subprocess.run([tool, "reads.fa.gz", "db_ours", "query", "0", str(2**31 - 1), "20000"], cwd=work, check=True)
    subprocess.run([tool, "ref.fa", "db_ours", "subject", "0", str(2**31 - 1), "50000"], cwd=work, check=True)
    n = 0
    for title in ("query", "subject"):
        for ext in ("pac", "hdr", "seq_info", "ambig", "volume_info.bin", "volume_info.txt", "hbndb_is_built"):
            a = os.path.join(work, "db_ref", f"{title}.{ext}")
            b = os.path.join(work, "db_ours", f"{title}.{ext}")
            if ext == "ambig":          # CAmbigSubseq{size_t; char; int}: the reference leaves the 3 padding bytes uninitialised
                ra = np.fromfile(a, dtype=np.uint8).reshape(-1, 16)
                rb = np.fromfile(b, dtype=np.uint8).reshape(-1, 16)
                keep = [0, 1, 2, 3, 4, 5, 6, 7, 8, 12, 13, 14, 15]
                assert ra.shape == rb.shape and np.array_equal(ra[:, keep], rb[:, keep]), title
            else:
                assert filecmp.cmp(a, b, shallow=False), (title, ext)
            n += 1
    assert n == 14
    assert os.path.getsize(os.path.join(work, "db_ref", "query.volume_info.bin")) > 3 * 64      # several volumes
    assert os.path.getsize(os.path.join(work, "db_ref", "query.ambig")) > 0


def test_qx2splitseq_equals_reference(tmp_path):
    """Subread splitting (src/app/split_seq/main.c): same pieces, same names, byte-identical output; FASTA, gzip and
    FASTQ input, with and without overlap / last-piece merging."""
    rng = np.random.default_rng(8)
    work = str(tmp_path)
    subprocess.run(["make", "-s", "-C", os.path.join(H.ROOT, "lesv_b200", "csrc"), "../bin/qx2splitseq"], check=True)
    ours = os.path.join(H.ROOT, "lesv_b200", "bin", "qx2splitseq")
    ref = os.path.join(H.ROOT, "oracle", "_ref", "bin", "qx2splitseq")
    lens = [500, 4999, 5000, 5001, 6999, 7000, 7001, 12000, 26000, 10, 15000]
    seqs = ["".join("ACGT"[x] for x in rng.integers(0, 4, n)) for n in lens]
    with open(os.path.join(work, "a.fa"), "w") as f:
        for i, s in enumerate(seqs):
            f.write(f">read{i} comment {i}\n")
            for k in range(0, len(s), 60):
                f.write(s[k:k + 60] + "\n")
    with gzip.open(os.path.join(work, "b.fa.gz"), "wt") as f:
        for i, s in enumerate(seqs[:5]):
            f.write(f">gz{i}\n{s}\n")
    with open(os.path.join(work, "c.fq"), "w") as f:
        for i, s in enumerate(seqs[5:9]):
            f.write(f"@fq{i} x\n{s}\n+\n{'I' * len(s)}\n")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    n = 0
    for inp in ("a.fa", "b.fa.gz", "c.fq"):
        for args in (("5000", "0", "2000"), ("5000", "300", "0"), ("3000", "100", "2999")):
            a = subprocess.run([ref, *args, inp], cwd=work, check=True, capture_output=True, env=env).stdout
            b = subprocess.run([ours, *args, inp], cwd=work, check=True, capture_output=True).stdout
            assert a == b, (inp, args)
            n += a.count(b">")
    assert n > 100


def test_trf_loader_rejects_what_the_reference_asserts(tmp_path):
    """TrfArrayBuild (app/necat2sv/trf_array.cpp:31-46) asserts three tab-separated columns, from < to and to <= contig size;
    the shared loader of qx2svr / qx2svsig (host/lesv_db.hpp) dies on the same lines instead of writing past the mask.
    The mask is loaded before any GPU is touched, so this runs without one."""
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(H.ROOT, "lesv_b200", "csrc"), "../bin/qx2svsig"], check=True)
    gold = os.path.join(H.ROOT, "tests", "golden", "svsig_small")
    S = H.LesvDb(os.path.join(gold, "db"), "subject")
    name, size = S.name(0), int(S.seq_size[0])
    tool = os.path.join(H.ROOT, "lesv_b200", "bin", "qx2svsig")
    for bad, msg in ((f"{name}\t10\t20\textra\n", "3 tab-separated columns"), (f"{name}\t30\t30\n", "from >= to"),
                     (f"{name}\t5\t{size + 1}\n", "behind the contig"), (f"{name} 10 20\n", "3 tab-separated columns"),
                     (f"{name}\tx\t20\n", "bad interval")):
        bed = tmp_path / "bad.bed"
        bed.write_text(f"{name}\t0\t100\nunknown_contig\t5\t50\n" + bad)
        r = subprocess.run([tool, str(tmp_path / "out"), os.path.join(gold, "svr"), os.path.join(gold, "db"), "40", "1", str(bed)],
                           capture_output=True, text=True)
        assert r.returncode != 0 and msg in r.stderr, (bad, r.stderr[-300:])
