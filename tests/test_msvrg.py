"""qx2msvrg (the step after the hot path; host-only) against the compiled reference on the committed
golden signatures and on synthetic signature files with dense, sparse and length-heterogeneous clusters."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers as H

BIN = os.path.join(H.ROOT, "lesv_b200", "bin")
REF_BIN = os.path.join(H.ROOT, "oracle", "_ref", "bin")
GOLD = os.path.join(H.ROOT, "tests", "golden", "svsig_small")


def _make_signatures(path, rng, ssize=2_000_000, n_sites=160):
    """INS block then DEL block, each sorted by sfrom (what qx2svsig writes)."""
    blocks = []
    for typ in (6, 0):
        recs = []
        for _ in range(n_sites):
            pos = int(rng.integers(1000, ssize - 20000))
            size = int(np.exp(rng.uniform(np.log(40), np.log(8000))))
            depth = int(rng.integers(1, 14))
            spread = int(rng.choice([0, 2, 8, 30, 90]))               # breakpoint jitter of the site
            lenvar = float(rng.choice([0.0, 0.03, 0.2]))
            for _ in range(depth):
                sfrom = pos + int(rng.integers(-spread, spread + 1))
                n = max(40, int(size * (1 + rng.uniform(-lenvar, lenvar))))
                qid, qsize = int(rng.integers(0, 100000)), 50000
                qfrom = int(rng.integers(100, 40000))
                fs = max(0, sfrom - 20000)
                if typ == 6:
                    rec = (qid, int(rng.integers(0, 2)), qfrom, qfrom + n, 10, qsize - 10, qsize, sfrom, sfrom + 1, fs, fs + 45000, ssize, 6)
                else:
                    rec = (qid, int(rng.integers(0, 2)), qfrom, qfrom + 1, 10, qsize - 10, qsize, sfrom, sfrom + n, fs, fs + 45000, ssize, 0)
                recs.append(rec)
        recs.sort(key=lambda r: r[7])
        blocks += recs
    with open(path, "w") as f:
        for r in blocks:
            f.write("\t".join(str(x) for x in r) + f"\tread{r[0]}\n")


def _run(tool, work, out_dir, svsig_dir, db):
    subprocess.run([tool, out_dir, svsig_dir, db], cwd=work, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return {f: open(os.path.join(work, out_dir, f), "rb").read() for f in sorted(os.listdir(os.path.join(work, out_dir)))}


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BIN, "qx2msvrg")), reason="oracle/_ref not built")
def test_qx2msvrg_equals_reference(tmp_path):
    subprocess.run(["make", "-s", "-C", os.path.join(H.ROOT, "lesv_b200", "csrc"), "../bin/qx2msvrg"], check=True)
    work = str(tmp_path)
    db = os.path.join(GOLD, "db")
    # (1) the golden pipeline run of the reference
    a = _run(os.path.join(REF_BIN, "qx2msvrg"), work, "grp_ref", os.path.join(GOLD, "svsig"), db)
    b = _run(os.path.join(BIN, "qx2msvrg"), work, "grp_ours", os.path.join(GOLD, "svsig"), db)
    assert a == b and any(k.endswith(".created") for k in a)
    # (2) synthetic signature files, one per contig of the golden database
    n_contigs = H.LesvDb(db, "subject").num_seqs
    os.makedirs(os.path.join(work, "sig"))
    rng = np.random.default_rng(7)
    for sid in range(n_contigs):
        _make_signatures(os.path.join(work, "sig", f"subject_{sid:08d}.sv_signature"), rng)
    a = _run(os.path.join(REF_BIN, "qx2msvrg"), work, "grp2_ref", "sig", db)
    b = _run(os.path.join(BIN, "qx2msvrg"), work, "grp2_ours", "sig", db)
    assert a == b
    text = a["subject_00000000_sv_read_group"].decode()
    n_groups = sum(1 for l in text.splitlines() if l.count("\t") == 1)
    assert n_groups >= 60                                              # strict and relaxed groups, INS and DEL
    # resume: with the .created sentinel in place nothing is rewritten
    os.remove(os.path.join(work, "grp2_ours", "subject_00000000_sv_read_group"))
    _run(os.path.join(BIN, "qx2msvrg"), work, "grp2_ours", "sig", db)
    assert not os.path.exists(os.path.join(work, "grp2_ours", "subject_00000000_sv_read_group"))
