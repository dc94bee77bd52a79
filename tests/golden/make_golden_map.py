#!/usr/bin/env python
"""Regenerate tests/golden/map_small/ from the compiled, UNMODIFIED reference (oracle/_ref):

  db/                 packed database written by qx2map (3 contigs with a planted repeat family, 48 reads)
  cand.subseq.txt     `qx2map -outfmt subseq -num_threads 1`: the candidate windows of every read
                      (HbnSubseqHit, app/map/hbn_subseq_hit.h) = output of index + seeding + chaining +
                      find_candidate_subject_subseqs
  index.npz           the reference's LookupTable (build_lookup_table called through ctypes)
  matches.npy         the reference's DDFKmerMatch records (hbn_find_kmer_matches_mt, one thread),
                      sorted by (context, soff, qoff)
  map.cigar.m4        `qx2map -cigar` output, sorted (pins the extension stage)

Run from the repo root:  python tests/golden/make_golden_map.py
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lesv_b200 import synth  # noqa: E402
import refmap  # noqa: E402


def make_inputs(work, seed=7):
    rng = np.random.default_rng(seed)
    contigs = [synth.random_codes(n, rng) for n in (150000, 90000, 60001)]
    # a dispersed repeat family: 260 copies of one 70-mer (index-side occurrence filter) ...
    rep = synth.random_codes(70, rng)
    for _ in range(260):
        c = contigs[int(rng.integers(0, 3))]
        p = int(rng.integers(0, len(c) - 100))
        c[p:p + 70] = rep
    # two tandem repeats on contig 2: period 10 (= the index window: one k-mer 500 times -> removed from
    # the index, lookup_table.c:95-138) and period 25 (5 distinct sampled k-mers x 100 copies: kept in the
    # index, but the reads over it carry each of them hundreds of times -> query-side filter,
    # hbn_word_finder.c:292-294, and the "> 4 query positions per subject offset" filter,
    # chain_and_extend_kmer_matches.c:105-112)
    contigs[2][10000:15000] = np.tile(synth.random_codes(10, rng), 500)
    contigs[2][30000:35000] = np.tile(synth.random_codes(25, rng), 200)
    # ... and a 3 kb segmental duplication between contig 0 and contig 1 (two candidate contigs per read)
    contigs[1][20000:23000] = contigs[0][50000:53000]
    reads = []
    for i in range(48):
        c = contigs[i % 3]
        L = int(rng.integers(3000, 9000))
        s = int(rng.integers(0, len(c) - L))
        if i == 5:
            s = 49000                                   # spans the duplicated segment
            c = contigs[0]
        if i >= 36:                                     # over / next to the period-25 tandem repeat
            c = contigs[2]
            s = int(rng.integers(26000, 34000))
            L = min(L, len(c) - s)
        sample = c[s:s + L].copy()
        if i % 7 == 3:                                  # a deletion in the sample
            sample = np.concatenate([sample[:L // 2], sample[L // 2 + 300:]])
        r = synth.mutate(sample, 0.12, rng)
        if rng.random() < 0.5:
            r = synth.revcomp_codes(r)
        reads.append(np.ascontiguousarray(r))
    synth.write_fasta(os.path.join(work, "ref.fa"), [f"contig{c}" for c in range(3)], contigs, width=80)
    synth.write_fasta(os.path.join(work, "reads.fa"), [f"read{i}" for i in range(len(reads))], reads)


def main():
    work = os.path.join(ROOT, "tests", "golden", "map_small")
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    make_inputs(work)
    common = ["-max_target_seqs", "5", "-max_hsps", "10"]
    refmap.run_qx2map(work, ["-outfmt", "subseq"] + common + ["-db_dir", "db", "-keep_db"], "cand.subseq.txt")
    refmap.run_qx2map(work, ["-cigar"] + common + ["-db_dir", "db_cigar"], "map.cigar.raw")
    with open(os.path.join(work, "map.cigar.raw")) as f:
        lines = sorted(f.readlines())
    with open(os.path.join(work, "map.cigar.m4"), "w") as f:
        f.writelines(lines)
    db = os.path.join(work, "db")
    off, hashes, cnt, handles = refmap.ref_index(db)
    np.savez_compressed(os.path.join(work, "index.npz"), offset_list=off, hashes=hashes, counts=cnt)
    np.save(os.path.join(work, "matches.npy"), refmap.ref_matches(db, handles=handles))
    for name in ("map.cigar.raw", "log.txt"):
        os.remove(os.path.join(work, name))
    shutil.rmtree(os.path.join(work, "db_cigar"), ignore_errors=True)
    shutil.rmtree(os.path.join(db, "backup_results"), ignore_errors=True)
    for f in os.listdir(db):
        if f.endswith(".txt") or f.endswith("_is_built"):
            os.remove(os.path.join(db, f))
    print("map_small written:", len(off), "index offsets,", len(hashes), "distinct k-mers")


if __name__ == "__main__":
    main()
