"""Seeded synthetic inputs of the shapes BASELINE.json names.

Everything here is plain numpy and deterministic for a given seed, so the reference binaries
(oracle/_ref), the oracle restatement and the CUDA path all see byte-identical inputs.

Vocabulary follows lesv: a *reference* (subject) of one or more contigs, a *sample* genome
that carries planted INS/DEL events relative to the reference, and noisy *reads* drawn from
the sample (ONT-like: substitutions / insertions / deletions, random strand).
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def random_codes(n: int, rng: np.random.Generator) -> np.ndarray:
    """n i.i.d. uniform residues as codes 0..3 (A,C,G,T)."""
    return rng.integers(0, 4, size=n, dtype=np.uint8)


def codes_to_ascii(codes: np.ndarray) -> bytes:
    return _ACGT[codes].tobytes()


def revcomp_codes(codes: np.ndarray) -> np.ndarray:
    return _COMP[codes[::-1]]


def plant_events(ref: np.ndarray, rng: np.random.Generator, n_events: int,
                 sizes=(40, 60, 100, 300, 1000, 3000), min_gap: int = 20000):
    """Return (sample, events): the sample genome = ref with alternating INS/DEL planted.

    events = list of (kind, ref_pos, size); kind 'INS' means the sample has `size` extra bases
    after ref_pos, 'DEL' means ref[ref_pos:ref_pos+size] is missing from the sample.
    """
    n = len(ref)
    if n_events == 0:
        return ref.copy(), []
    slot = n // (n_events + 1)
    assert slot >= min_gap or n_events <= 1, "reference too short for that many events"
    pieces, events, prev = [], [], 0
    for i in range(n_events):
        pos = slot * (i + 1) + int(rng.integers(-slot // 8, slot // 8 + 1))
        size = int(sizes[int(rng.integers(0, len(sizes)))])
        if i % 2 == 0:
            pieces.append(ref[prev:pos])
            pieces.append(random_codes(size, rng))
            prev = pos
            events.append(("INS", pos, size))
        else:
            pieces.append(ref[prev:pos])
            prev = pos + size
            events.append(("DEL", pos, size))
    pieces.append(ref[prev:])
    return np.concatenate(pieces), events


def mutate(seq: np.ndarray, err: float, rng: np.random.Generator,
           mix=(0.4, 0.3, 0.3)) -> np.ndarray:
    """ONT-like noise: per-base event with probability `err`, split sub/ins/del by `mix`."""
    n = len(seq)
    u = rng.random(n)
    sub = u < err * mix[0]
    ins = (u >= err * mix[0]) & (u < err * (mix[0] + mix[1]))
    dele = (u >= err * (mix[0] + mix[1])) & (u < err)
    out = seq.copy()
    # substitution: add 1..3 mod 4 so the base always changes
    out[sub] = (out[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    keep = ~dele
    # insertion: emit one random base after the position
    reps = keep.astype(np.int64) + ins.astype(np.int64)
    idx = np.repeat(np.arange(n), reps)
    res = out[idx]
    # positions that are the *second* copy of an inserted site get a random base
    first = np.ones(len(idx), dtype=bool)
    first[1:] = idx[1:] != idx[:-1]
    second = ~first
    # when a site is deleted AND flagged ins (impossible: masks are disjoint) - nothing to do
    res[second] = rng.integers(0, 4, size=int(second.sum()), dtype=np.uint8)
    return res


def mutate_with_map(seq: np.ndarray, err: float, rng: np.random.Generator, mix=(0.4, 0.3, 0.3)):
    """mutate() that also returns where every source position starts in the result (len(seq) + 1 entries), so that
    a slice seq[a:b] and its noisy copy result[pos[a]:pos[b]] can be cut out of one long sequence without copying."""
    n = len(seq)
    u = rng.random(n)
    sub = u < err * mix[0]
    ins = (u >= err * mix[0]) & (u < err * (mix[0] + mix[1]))
    dele = (u >= err * (mix[0] + mix[1])) & (u < err)
    out = seq.copy()
    out[sub] = (out[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    reps = (~dele).astype(np.int64) + ins.astype(np.int64)
    pos = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(reps, out=pos[1:])
    idx = np.repeat(np.arange(n), reps)
    res = out[idx]
    second = np.zeros(len(idx), dtype=bool)
    second[1:] = idx[1:] == idx[:-1]
    res[second] = rng.integers(0, 4, size=int(second.sum()), dtype=np.uint8)
    return res, pos


def simulate_reads(sample: np.ndarray, n_reads: int, read_len: int, rng: np.random.Generator,
                   err_lo: float = 0.12, err_hi: float = 0.12):
    """Draw `n_reads` windows of `read_len` sample bases, mutate, random strand."""
    reads = []
    hi = len(sample) - read_len
    for _ in range(n_reads):
        s = int(rng.integers(0, hi + 1))
        err = err_lo if err_hi <= err_lo else float(rng.uniform(err_lo, err_hi))
        r = mutate(sample[s:s + read_len], err, rng)
        if rng.random() < 0.5:
            r = revcomp_codes(r)
        reads.append(r)
    return reads


def reads_with_noisy_tails(ref: np.ndarray, n_reads: int, rng: np.random.Generator, core_len: int = 8000):
    """Reads whose ends are unrelated sequence (150-2,600 bp) or the true flank at 30-45 % error: edlib's end extension
    (reference src/algo/hbn_traceback.c:436-486) gives up inside the tail, which leaves an overhang of <= 3 kb, so
    hbn_traceback's left_extend / right_extend run the DALIGNER wave (:214-311)."""
    reads = []
    for i in range(n_reads):
        s0 = int(rng.integers(1200, len(ref) - core_len - 1000))
        core = mutate(ref[s0:s0 + core_len], float(rng.uniform(0.08, 0.14)), rng)

        def tail(seg):
            if int(rng.integers(0, 3)) == 0:
                return random_codes(int(rng.integers(150, 2600)), rng)
            return mutate(seg, float(rng.uniform(0.30, 0.45)), rng)
        left = tail(ref[s0 - 1200:s0]) if i % 3 != 2 else np.zeros(0, np.uint8)
        right = tail(ref[s0 + core_len:s0 + core_len + 1000]) if i % 3 != 1 else np.zeros(0, np.uint8)
        rd = np.concatenate([left, core, right])
        reads.append(revcomp_codes(rd) if i % 2 else rd)
    return reads


def consensus_reads(contigs, n_reads: int, rng: np.random.Generator, max_len: int = 30000):
    """Consensus-like reads as qx2csvrg hands them to qx2asvr (reference src/app/necat2sv/map_cns_sv_read.c:170-292):
    a reference window of 3-`max_len` bp carrying one planted INS/DEL of 40 bp-3 kb, 0.5-4 % residual error, stored
    reverse-complemented when the header says REV, named
    <name>_svr:<dir>:<sid>:<gid>:<sfrom>:<sto>_cns:<a>:<b>:<from>:<to>:<size>:<raw> (sv_read_group_file_name.cpp:102-196,
    corelib/cns_read_header.c:3-95).  Returns per contig a list of (name, codes)."""
    out = [[] for _ in contigs]
    for i in range(n_reads):
        sid = int(rng.integers(0, len(contigs)))
        c = contigs[sid]
        L = int(rng.integers(3000, max_len))
        sfrom = int(rng.integers(0, len(c) - L))
        sto = sfrom + L
        seg = c[sfrom:sto]
        size = int(np.exp(rng.uniform(np.log(40), np.log(3000))))
        k = int(rng.integers(L // 4, 3 * L // 4))
        if i % 2 == 0:
            seg = np.concatenate([seg[:k], random_codes(size, rng), seg[k:]])              # insertion in the sample
        elif size < L // 3:
            seg = np.concatenate([seg[:k], seg[k + size:]])                                # deletion
        err = float(rng.uniform(0.005, 0.04)) if i % 7 else float(rng.uniform(0.10, 0.22))   # a few bad ones (< 85 % eff. identity)
        rd = mutate(seg, err, rng)
        n = len(rd)
        cns_from, cns_to = int(rng.integers(0, n // 10)), n - int(rng.integers(0, n // 10))
        rdir = int(rng.integers(0, 2))
        if rdir:
            rd = revcomp_codes(rd)
        name = f"read{i}_svr:{rdir}:{sid}:{i % 50}:{sfrom}:{sto}_cns:{3 + i % 5}:{10 + i % 9}:{cns_from}:{cns_to}:{n}:{n + 37}"
        out[sid].append((name, np.ascontiguousarray(rd)))
    return out


def write_fasta(path: str, names, seqs, width: int = 0) -> None:
    with open(path, "wb") as f:
        for name, s in zip(names, seqs):
            f.write(b">" + name.encode() + b"\n")
            a = codes_to_ascii(s)
            if width:
                for i in range(0, len(a), width):
                    f.write(a[i:i + width] + b"\n")
            else:
                f.write(a + b"\n")


def make_dataset(outdir: str, ref_len: int, n_reads: int, read_len: int, n_events: int,
                 seed: int = 1, err_lo: float = 0.12, err_hi: float = 0.12,
                 n_contigs: int = 1, event_sizes=(40, 60, 100, 300, 1000, 3000)):
    """Write ref.fa + reads.fa under outdir; return dict with paths and planted events."""
    import os
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    contigs, samples, all_events = [], [], []
    per = ref_len // n_contigs
    ev_per = n_events // n_contigs
    for c in range(n_contigs):
        ref = random_codes(per, rng)
        sample, events = plant_events(ref, rng, ev_per, sizes=event_sizes,
                                      min_gap=min(20000, per // (ev_per + 2)))
        contigs.append(ref)
        samples.append(sample)
        all_events.append(events)
    rrng = np.random.default_rng(seed + 1)
    reads = []
    for i in range(n_reads):
        c = i % n_contigs
        reads.extend(simulate_reads(samples[c], 1, read_len, rrng, err_lo, err_hi))
    ref_path = os.path.join(outdir, "ref.fa")
    reads_path = os.path.join(outdir, "reads.fa")
    write_fasta(ref_path, [f"contig{c}" for c in range(n_contigs)], contigs, width=80)
    write_fasta(reads_path, [f"read{i}" for i in range(n_reads)], reads)
    return {"ref": ref_path, "reads": reads_path, "events": all_events,
            "contigs": contigs, "reads_codes": reads}
