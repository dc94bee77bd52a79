#!/usr/bin/env python
"""Regenerate tests/golden/ from the compiled, UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference/src).  The reference ships no golden vectors of its own
(SURVEY.md §4), so these fixtures are outputs of the reference itself on seeded synthetic inputs:

  ksw_kat.json          ksw_extd2_sse (flag 0, zdrop -1) CIGAR + score for a few (query, target, band)
  svsig_small/          a complete tiny run of  qx2splitseq -> qx2map -> qx2m4x -> qx2svr -> qx2svsig :
      db/               packed database written by qx2map (query.* / subject.*) + backup_results/*.m4x (qx2m4x)
      map.m4            qx2map output (sorted: line order is thread-schedule dependent)
      svr/              qx2svr output (Q*.sv_read, subject_*.sv_read)
      svsig/            qx2svsig output (subject_*.sv_signature)

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lesv_b200 import synth  # noqa: E402
import helpers as H  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
OUT = os.path.join(ROOT, "tests", "golden")


def run(cmd, cwd, stdout=None):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    with open(os.path.join(cwd, "log.txt"), "ab") as log:
        subprocess.run(cmd, cwd=cwd, check=True, env=env, stdout=stdout or log, stderr=log)


def reference_pipeline(work, threads=4, min_indel=40):
    """qx2splitseq .. qx2svsig in `work` (expects ref.fa, reads.fa); returns nothing."""
    with open(os.path.join(work, "split.fa"), "wb") as f:
        run([os.path.join(BIN, "qx2splitseq"), "50000", "0", "20000", "reads.fa"], work, stdout=f)
    with open(os.path.join(work, "map.raw.m4"), "wb") as f:
        run([os.path.join(BIN, "qx2map"), "-max_target_seqs", "5", "-max_hsps", "10", "-db_dir", "db", "-keep_db",
             "-num_threads", str(threads), "split.fa", "ref.fa"], work, stdout=f)
    run([os.path.join(BIN, "qx2m4x"), "db"], work)
    run([os.path.join(BIN, "qx2svr"), "svr", "db", "3000", "70.0", "300", str(threads)], work)
    run([os.path.join(BIN, "qx2svsig"), "svsig", "svr", "db", str(min_indel), str(threads)], work)
    with open(os.path.join(work, "map.raw.m4")) as f:
        lines = sorted(f.readlines())
    with open(os.path.join(work, "map.m4"), "w") as f:
        f.writelines(lines)
    os.remove(os.path.join(work, "map.raw.m4"))


def main():
    assert os.path.exists(os.path.join(BIN, "qx2map")), "build the reference first: make -C oracle ref"
    # ---- ksw known-answer vectors -------------------------------------------------------------
    rng = np.random.default_rng(20261018)
    kat = []
    for it in range(40):
        q, t, w = H.random_pair(rng, it, max_len=300, small=50)
        cig, score = H.ref_ksw(q, t, w)
        kat.append({"q": "".join("ACGT"[x] for x in q), "t": "".join("ACGT"[x] for x in t), "w": int(w),
                    "cigar": H.cigar_str(cig), "score": int(score)})
    with open(os.path.join(OUT, "ksw_kat.json"), "w") as f:
        json.dump(kat, f, indent=0)
    # ---- tiny end-to-end run ---------------------------------------------------------------------
    work = os.path.join(OUT, "svsig_small")
    shutil.rmtree(work, ignore_errors=True)
    synth.make_dataset(work, ref_len=120000, n_reads=60, read_len=4000, n_events=4, seed=11,
                       event_sizes=(40, 60, 100, 300))
    reference_pipeline(work)
    # keep what the tests need, drop bulky intermediates
    for name in ("split.fa", "log.txt"):
        os.remove(os.path.join(work, name))
    br = os.path.join(work, "db", "backup_results")
    for f in os.listdir(br):
        if not f.endswith(".m4x"):
            os.remove(os.path.join(br, f))
    for f in os.listdir(os.path.join(work, "db")):
        if f.endswith(".txt") or f.endswith("_is_built"):
            os.remove(os.path.join(work, "db", f))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
