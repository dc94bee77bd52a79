#!/usr/bin/env python
"""bench.py -- headline benchmark of the lesv hot path on B200 (contract: see the task statement).

Metric (BASELINE.json): aligned bases/s and GCUPS per B200 on configs[1] -- 50 kb subreads at 10-15 %
ONT-like error against a 50 Mbp synthetic contig with planted 40 bp-10 kb INS/DEL, "mapping + SV
signatures on 1 B200".  One *step* takes one batch of reads through the whole hot path, calling the
reference-facing C ABI (include/lesvgpu.h) with HOST buffers exactly as the drop-in tools do:

    lesvgpu_map_batch    qx2map   : seeding, chaining, candidate windows, MEM stage, edlib/DALIGNER alignment
    (m4 text round trip) qx2m4x   : identity through its 6-significant-digit text, reverse-strand flip
    lesvgpu_svr_batch    qx2svr   : one trusted placement per read (ksw2 re-alignment where HSPs compete/split)
    lesvgpu_svsig_batch  qx2svsig : banded dual-affine DP + traceback + signature scan per placed read

The subject .pac and its k-mer index are resident (built once per reference, outside the step, like the
reference's per-volume build_lookup_table; the build time is reported as index_build_ms).

  value      aligned read bases (sum of qend-qoff over emitted HSPs) / summed CUDA-event time of the GPU work
             of all stages (inputs resident; host<->device copies and host logic excluded)
  e2e        the same bases / wall time of the step: H2D of the packed reads, all host logic, D2H of results
  gcups      DP cells (ksw_extd2_sse's count, reference src/algo/ksw2_extd2_sse.c:124-138) of the qx2svr +
             qx2svsig alignments / their kernel time
  roofline   integer-issue roofline of ksw_dp_kernel, the dominant kernel, plus the HBM view
  cpu_baseline / --impl reference : the UNMODIFIED reference tools (oracle/_ref/bin, compiled from
             /root/reference/src) on all host cores over a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from lesv_b200 import shard, synth  # noqa: E402

READ_LEN = 50000
REF_LEN = 50_000_000
MIN_INDEL = 40
OPS_PER_CELL = 46            # scalar formulation of ksw2_extd2_sse.c:31-58,227-265,318-349
LANES_PACKED = 2             # two DP cells per 32-bit integer instruction (16-bit lanes)
METRIC = "aligned bases/s (qx2map + qx2svr + qx2svsig stages, C2)"
METRIC_C3 = "aligned bases/s (qx2map + qx2svr + qx2svsig stages, C3)"


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def pack_pac(seqs):
    """codes -> lesv .pac bytes (base l at byte l>>2, bits (~l&3)<<1; each sequence 4-base aligned)."""
    offs, total = [], 0
    for s in seqs:
        offs.append(total)
        total += (len(s) + 3) // 4 * 4
    buf = np.zeros(total, dtype=np.uint8)
    for o, s in zip(offs, seqs):
        buf[o:o + len(s)] = s
    b = buf.reshape(-1, 4)
    pac = (b[:, 0] << 6) | (b[:, 1] << 4) | (b[:, 2] << 2) | b[:, 3]
    return pac.astype(np.uint8), np.asarray(offs, dtype=np.int64)


def make_genomes(seed: int, ref_len: int = REF_LEN):
    """(reference contig, sample genome): the sample carries INS/DEL of 40 bp-10 kb (log-uniform), ~1 per 250 kb."""
    rng = np.random.default_rng(seed)
    ref = synth.random_codes(ref_len, rng)
    n_events = max(1, ref_len // 250_000)
    slot = ref_len // (n_events + 1)
    pieces, prev = [], 0
    for i in range(n_events):
        pos = slot * (i + 1) + int(rng.integers(-slot // 8, slot // 8 + 1))
        size = int(np.exp(rng.uniform(np.log(40), np.log(10000))))
        pieces.append(ref[prev:pos])
        if rng.random() < 0.5:
            pieces.append(synth.random_codes(size, rng))         # insertion in the sample
            prev = pos
        else:
            prev = pos + size                                     # deletion from the sample
    pieces.append(ref[prev:])
    return ref, np.concatenate(pieces)


def make_reads(sample, n_reads: int, seed: int, read_len: int = READ_LEN):
    """n_reads subreads of the sample genome, 10-15 % error (0.4 sub / 0.3 ins / 0.3 del), random strand."""
    rng = np.random.default_rng(seed)
    reads = []
    for _ in range(n_reads):
        s = int(rng.integers(0, len(sample) - read_len))
        r = synth.mutate(sample[s:s + read_len], float(rng.uniform(0.10, 0.15)), rng)
        if rng.integers(0, 2):
            r = synth.revcomp_codes(r)
        reads.append(np.ascontiguousarray(r))
    return reads


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, gpu_index: int, active: bool = True):
        self.active = active                             # only rank 0 polls nvidia-smi (its line is the one reported)
        self.idx = gpu_index
        self.samples, self.reasons, self.smax = [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.smax = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        if self.active:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        hi = [x for x in s if self.smax and x > 0.5 * self.smax] or s
        return {"sm_mhz": (hi[len(hi) // 2] if hi else None), "sm_max_mhz": self.smax, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference tools on the host cores
# ---------------------------------------------------------------------------------------------
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def reference_pipeline(ref, reads, cores: int, keep=None):
    """qx2map -> qx2m4x -> qx2svr -> qx2svsig of the compiled reference (oracle/_ref/bin) on `reads`.
    Returns a dict of seconds per stage, aligned bases, DP cells and signature count.  The mapping time is the
    reference's own 'mapping query ...' timer (hbn_align_one_volume.c:124,141): database packing and the
    index build are one-off costs per reference and are reported separately.  keep = dict: the m4x records, SvReads and
    signatures the reference wrote are left in it (for the parity block)."""
    work = tempfile.mkdtemp(prefix="lesv_ref_")
    try:
        synth.write_fasta(os.path.join(work, "ref.fa"), ["chr1"], [ref], width=0)
        synth.write_fasta(os.path.join(work, "reads.fa"), [f"read{i}" for i in range(len(reads))], reads, width=0)
        env = dict(os.environ, OMP_NUM_THREADS="1")
        t = str(cores)
        res = {}

        def run(name, cmd, stdout=None):
            t0 = time.perf_counter()
            with open(os.path.join(work, name + ".log"), "wb") as log:
                subprocess.run(cmd, cwd=work, check=True, env=env, stdout=stdout or log, stderr=log)
            res[name + "_wall_s"] = time.perf_counter() - t0

        with open(os.path.join(work, "map.m4"), "wb") as f:
            run("qx2map", [os.path.join(REF_BIN, "qx2map"), "-max_target_seqs", "5", "-max_hsps", "10", "-db_dir", "db",
                           "-keep_db", "-num_threads", t, "reads.fa", "ref.fa"], stdout=f)
        log = open(os.path.join(work, "qx2map.log"), errors="replace").read()
        res["map_s"] = sum(float(x) for x in re.findall(r"'mapping query [0-9 -]+' takes ([0-9.]+) secs", log))
        if res["map_s"] <= 0:
            res["map_s"] = res["qx2map_wall_s"]
        run("qx2m4x", [os.path.join(REF_BIN, "qx2m4x"), "db"])
        run("qx2svr", [os.path.join(REF_BIN, "qx2svr"), "svr", "db", "3000", "70.0", "300", t])
        run("qx2svsig", [os.path.join(REF_BIN, "qx2svsig"), "svsig", "svr", "db", str(MIN_INDEL), t])
        bases = 0
        with open(os.path.join(work, "map.m4")) as f:
            for line in f:
                c = line.split("\t")
                bases += int(c[6]) - int(c[5])
        sig_lines, svr_lines = [], []
        for fn in sorted(os.listdir(os.path.join(work, "svsig"))):
            if fn.endswith(".sv_signature"):
                sig_lines += open(os.path.join(work, "svsig", fn)).readlines()
        for fn in sorted(os.listdir(os.path.join(work, "svr"))):
            if fn.startswith("subject_") and fn.endswith(".sv_read"):
                svr_lines += open(os.path.join(work, "svr", fn)).readlines()
        res["aligned_bases"] = bases
        res["signatures"] = len(sig_lines)
        res["path_s"] = res["map_s"] + res["qx2svr_wall_s"] + res["qx2svsig_wall_s"]
        if keep is not None:
            br = os.path.join(work, "db", "backup_results")
            raw = [np.fromfile(os.path.join(br, fn), dtype=np.uint8) for fn in sorted(os.listdir(br)) if fn.endswith(".m4x")]
            keep["m4x"] = np.concatenate(raw).reshape(-1, 56)[:, :52] if raw else np.zeros((0, 52), np.uint8)
            keep["svr"] = [tuple(int(x) for x in l.split("\t")[:9]) for l in svr_lines]
            keep["sig"] = [tuple(int(x) for x in l.split("\t")[:13]) for l in sig_lines]
        return res
    finally:
        shutil.rmtree(work, ignore_errors=True)


def cpu_baseline(ref, reads, cores, keep=None):
    if not os.path.exists(os.path.join(REF_BIN, "qx2map")):
        return None
    r = reference_pipeline(ref, reads, cores, keep)
    return {"value": r["aligned_bases"] / r["path_s"], "unit": "bases/s", "cores": cores, "kind": "reference",
            "sample": f"{len(reads)} reads (16 per host thread) of the same workload through oracle/_ref/bin qx2map, qx2svr, qx2svsig on {cores} threads: "
                      f"mapping {r['map_s']:.2f} s (the tool's own timer, without database/index build; whole qx2map "
                      f"{r['qx2map_wall_s']:.2f} s), qx2svr {r['qx2svr_wall_s']:.2f} s, qx2svsig {r['qx2svsig_wall_s']:.2f} s",
            "stages_s": {"map": r["map_s"], "svr": r["qx2svr_wall_s"], "svsig": r["qx2svsig_wall_s"]},
            "signatures": r["signatures"]}


def sample_size(cores: int) -> int:
    """Reads in the reference leg: 16 per host thread, so that qx2map's 5-read chunks and qx2svr's few full-matrix
    alignments spread over all threads instead of leaving most of them idle behind one straggler."""
    return max(16 * cores, 64)


# ---------------------------------------------------------------------------------------------
_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, the reference tools) write there
    too, so fd 1 is pointed at stderr for the whole run and the result line goes to the original stdout."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


HG19_MBP = [249, 243, 198, 191, 181, 171, 159, 146, 141, 136, 135, 134, 115, 107, 103, 90, 81, 78, 59, 63, 48, 51, 155, 59]


def make_c3_reference(seed: int, scale: float = 1.0):
    """BASELINE.json configs[2]: 24 contigs with hg19's length profile (3.1 Gbp at scale 1), i.i.d. uniform residues."""
    rng = np.random.default_rng(seed)
    return [synth.random_codes(int(mb * 1_000_000 * scale), rng) for mb in HG19_MBP]


def make_reads_c3(contigs, n_reads: int, seed: int, read_len: int = READ_LEN):
    """50 kb subreads at 10-15 % error drawn from the contigs in proportion to their length; every fourth read carries a
    planted 40 bp-10 kb insertion or deletion (the sample genome of a 3.1 Gbp reference is not materialised)."""
    rng = np.random.default_rng(seed)
    sizes = np.array([len(c) for c in contigs], dtype=np.float64)
    reads = []
    for i in range(n_reads):
        c = contigs[int(rng.choice(len(contigs), p=sizes / sizes.sum()))]
        p0 = int(rng.integers(0, len(c) - read_len - 10000))
        seg = c[p0:p0 + read_len]
        if i % 4 == 0:
            size = int(np.exp(rng.uniform(np.log(40), np.log(10000))))
            k = int(rng.integers(5000, read_len - 15000))
            seg = np.concatenate([seg[:k], c[p0 + k + size:p0 + read_len + size]]) if i % 8 == 0 else \
                np.concatenate([seg[:k], synth.random_codes(size, rng), seg[k:read_len - size]])
        r = synth.mutate(seg, float(rng.uniform(0.10, 0.15)), rng)
        if rng.integers(0, 2):
            r = synth.revcomp_codes(r)
        reads.append(np.ascontiguousarray(r))
    return reads


def make_reads_c4(contigs, n_subreads: int, seed: int):
    """BASELINE.json configs[3]: ultra-long source reads (log-normal, clipped to 150 kb-2.4 Mbp) cut like `qx2splitseq 50000 0
    20000` cuts them (reference src/app/split_seq/main.c:28-45: 50 kb pieces, a rest below 20 kb stays with the last piece);
    35 % of the source reads carry 20-30 % error, the others 8-15 % (README.md:90-95); an INS/DEL of 40 bp-10 kb about every
    250 kb.  Returns exactly n_subreads subreads."""
    rng = np.random.default_rng(seed)
    sizes = np.array([len(c) for c in contigs], dtype=np.float64)
    out = []
    while len(out) < n_subreads:
        c = contigs[int(rng.choice(len(contigs), p=sizes / sizes.sum()))]
        L = int(min(2_400_000, max(150_000, np.exp(rng.normal(np.log(200_000), 0.6)))))
        L = min(L, len(c) - 20_000)
        p0 = int(rng.integers(0, len(c) - L - 10_000))
        pieces, prev = [], p0
        for _ in range(max(1, L // 250_000)):
            pos = int(rng.integers(prev + 1000, p0 + L - 1000)) if prev + 1000 < p0 + L - 1000 else None
            if pos is None:
                break
            size = int(np.exp(rng.uniform(np.log(40), np.log(10000))))
            pieces.append(c[prev:pos])
            if rng.random() < 0.5:
                pieces.append(synth.random_codes(size, rng))
                prev = pos
            else:
                prev = min(pos + size, p0 + L)
        pieces.append(c[prev:p0 + L])
        err = float(rng.uniform(0.20, 0.30)) if rng.random() < 0.35 else float(rng.uniform(0.08, 0.15))
        long_read = synth.mutate(np.concatenate(pieces), err, rng)
        if rng.integers(0, 2):
            long_read = synth.revcomp_codes(long_read)
        n, a = len(long_read), 0
        while a < n:
            b = min(a + 50_000, n)
            if n - b < 20_000:
                b = n
            out.append(np.ascontiguousarray(long_read[a:b]))
            a = b
    return out[:n_subreads]


class Batch:
    """One query batch as the drop-in tools hand it to the library: packed reads (.pac layout) in HOST memory."""

    def __init__(self, reads, capi):
        self.reads = reads
        self.query_pac, qoffs = pack_pac(reads)
        self.qsizes = np.array([len(r) for r in reads], dtype=np.int64)
        self.read_offset = (qoffs + 1).astype(np.uint64)
        self.seq_arr = capi._seq_array(qoffs, self.qsizes)
        self.nq = len(reads)


def run_path(ctx, capi, B: Batch, subj, collect: bool = False, svr_ident: float = 70.0):
    """One batch through the path: qx2map -> (m4 text round trip) -> qx2svr -> qx2svsig, all through the C ABI with host
    buffers.  Returns per-stage device ms, byte counts and work counts (collect: also the records themselves)."""
    w0 = time.perf_counter()
    h = capi.map_batch_raw(ctx, B.query_pac, B.seq_arr, 0, B.nq, hitlist_size=5, max_hsps=10)
    w1 = time.perf_counter()
    mt, et = capi.map_last_timing(ctx), capi.map_last_extend_timing(ctx)
    m4 = capi.hsps_to_m4(h)
    w2 = time.perf_counter()
    svr, st = capi.svr_batch(ctx, B.query_pac, B.read_offset, subj["offsets"], subj["sizes"], m4, 3000, svr_ident, 300)
    w3 = time.perf_counter()
    t_svr = ctx.timing() if (st.n_refine_jobs or st.n_chain_jobs) else capi.Timing()
    svr_classes = ctx.class_timing() if (st.n_refine_jobs or st.n_chain_jobs) else []
    out = capi.svsig_contigs_raw(ctx, B.query_pac, B.read_offset, subj["seq_arr"], len(subj["offsets"]), svr, MIN_INDEL, want_sigs=collect)
    status, n_sig = out[0], out[1]
    w4 = time.perf_counter()
    t_sig = ctx.timing()
    dev = {"map_seed_chain": mt.lookup_ms + mt.emit_ms + mt.sort_ms + mt.chain_ms,
           "map_extend": et.ihf_ms + et.sort_ms + et.chain_ms + et.trace_ms,
           "svr": t_svr.pack_ms + t_svr.dp_ms + t_svr.scan_ms,
           "svsig": t_sig.pack_ms + t_sig.dp_ms + t_sig.scan_ms}
    wall = {"map": (w1 - w0) * 1e3, "m4": (w2 - w1) * 1e3, "svr": (w3 - w2) * 1e3, "svsig": (w4 - w3) * 1e3}
    r = {"dev": dev, "wall": wall, "dp_ms": t_svr.dp_ms + t_sig.dp_ms, "sig_dp_ms": t_sig.dp_ms, "sig_cells": int(t_sig.dp_cells),
         "svr_dp_ms": t_svr.dp_ms, "svr_cells": int(t_svr.dp_cells),
         "cells": int(t_svr.dp_cells) + int(t_sig.dp_cells), "dp_launches": t_svr.dp_launches + t_sig.dp_launches,
         "sig_launches": t_sig.dp_launches, "classes": {"svsig": ctx.class_timing(), "svr_last_round": svr_classes},
         "h2d": int(mt.h2d_bytes + t_svr.h2d_bytes + t_sig.h2d_bytes),
         "d2h": int(mt.d2h_bytes + et.d2h_bytes + t_svr.d2h_bytes + t_sig.d2h_bytes),
         "bases": int((h["qend"] - h["qoff"]).sum()), "hsps": len(h), "sv_reads": len(svr),
         "reads_ok": int((status == 1).sum()), "signatures": n_sig, "n_unsupported": int(et.n_unsupported),
         "svr_jobs": int(st.n_refine_jobs + st.n_chain_jobs)}
    if collect:
        r["records"] = {"m4": m4, "svr": svr, "sig": out[2]}
    return r


def parity_block(ours, ref_keep):
    """Our records for the sample reads against the files the reference leg wrote for the same reads (outside the
    timed region): M4Record multiset (first 52 bytes: the 4 trailing pad bytes are uninitialised in the reference),
    subject_*.sv_read and subject_*.sv_signature records as multisets of integer tuples."""
    from collections import Counter

    def diff(a, b):
        ca, cb = Counter(a), Counter(b)
        return sum(((ca - cb) + (cb - ca)).values())
    m4 = ours["m4"]
    mine_m4 = [bytes(x) for x in np.frombuffer(m4.tobytes(), dtype=np.uint8).reshape(-1, 56)[:, :52]]
    ref_m4 = [bytes(x) for x in ref_keep["m4x"]]
    svr = ours["svr"]
    mine_svr = [tuple(int(r[k]) for k in ("query_id", "qdir", "qoff", "qend", "qsize", "subject_id", "soff", "send", "dist")) for r in svr]
    mine_sig = [tuple(int(v) for v in r) for r in ours["sig"].tolist()]
    d = {"m4": diff(mine_m4, ref_m4), "sv_read": diff(mine_svr, ref_keep["svr"]), "sv_signature": diff(mine_sig, ref_keep["sig"])}
    return {"checked": {"m4": len(ref_m4), "sv_read": len(ref_keep["svr"]), "sv_signature": len(ref_keep["sig"])},
            "n_diff": sum(d.values()), "diff_by_file": d,
            "note": "reference leg on all host threads, GPU with the whole sample as one batch: the query-side k-mer occurrence "
                    "filter is per thread slice in the reference (SURVEY.md Appendix A #5), identical on non-repetitive references"}


def c5_block(ctx, capi, cores: int, seed: int, n_sm: int, f_mhz: float):
    """BASELINE.json configs[4] on a bounded sample: query/target pairs with lengths log-uniform in 1-50 kb (target = query
    mutated at 12 %), band w in {64 .. 1024} passed as max_distance; pairs whose length difference exceeds the band fail
    up front in both implementations and are left out.  GPU = lesvgpu_ksw_batch kernel time (CUDA events); reference =
    nw_ksw2_extd2 of oracle/_ref/libhbn_ref.so called from one Python thread per host core (ctypes releases the GIL)."""
    import ctypes as C
    rng = np.random.default_rng(seed + 5)
    n_pairs = int(os.environ.get("LESV_C5_PAIRS", "32768"))
    # the pairs are slices of one long random sequence and of its noisy copy (no per-pair generation, no copies): enough of
    # them that the batch is several waves of alignments deep, as BASELINE's 1 M pairs are -- with a few thousand pairs the time
    # is the critical path of the longest alignment (100 k anti-diagonals one after the other), not throughput
    lens = np.exp(rng.uniform(np.log(1000), np.log(50000), n_pairs)).astype(np.int64)
    big = synth.random_codes(int(min(lens.sum() + 50000, 96_000_000)), rng)
    noisy, pos = synth.mutate_with_map(big, 0.12, rng)
    base, a = [], 0
    for L in lens:
        if a + L > len(big):
            a = int(rng.integers(0, len(big) - 50000))
        base.append((big[a:a + L], noisy[pos[a]:pos[a + L]]))
        a += int(L)
    peak = n_sm * 128 * f_mhz * 1e6 / OPS_PER_CELL * LANES_PACKED / 1e9
    so = os.path.join(ROOT, "oracle", "_ref", "libhbn_ref.so")
    lib = None
    if os.path.exists(so):
        lib = C.CDLL(so)
        lib.Ksw2DataNew.restype = C.c_void_p
        lib.ksw2_extd2_set_params.argtypes = [C.c_void_p]
        lib.nw_ksw2_extd2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 7

    class KString(C.Structure):
        _fields_ = [("l", C.c_size_t), ("m", C.c_size_t), ("s", C.c_void_p)]

    def cpu_gcups(pairs, w):
        it, lock = iter(range(len(pairs))), threading.Lock()

        def worker():
            data = lib.Ksw2DataNew()
            lib.ksw2_extd2_set_params(data)
            qa, sa, ints, ident = KString(), KString(), [C.c_int() for _ in range(4)], C.c_double()
            while True:
                with lock:
                    i = next(it, None)
                if i is None:
                    return
                q, t = pairs[i]
                lib.nw_ksw2_extd2(data, 0, q.ctypes.data, 0, len(q), len(q), 0, t.ctypes.data, 0, len(t), len(t), 0,
                                  C.c_double(0.0), w, *[C.addressof(x) for x in ints], C.addressof(ident), C.addressof(qa), C.addressof(sa))
        ths = [threading.Thread(target=worker) for _ in range(cores)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        sec = time.perf_counter() - t0
        return sum(capi.ksw_cells(len(q), len(t), w)[0] for q, t in pairs) / sec / 1e9

    rows = []
    for w in (64, 128, 256, 512, 1024):
        good = [(q, t) for q, t in base if abs(len(q) - len(t)) <= w]
        if not good:
            continue
        qs, ts = [p[0] for p in good], [p[1] for p in good]
        ctx.ksw_batch(qs[:64], ts[:64], [w] * min(64, len(good)), want_cigars=False)          # warm-up
        res, _ = ctx.ksw_batch(qs, ts, [w] * len(good), want_cigars=False)
        tm = ctx.timing()
        g = tm.dp_cells / (tm.dp_ms / 1e3) / 1e9
        row = {"band": w, "pairs": len(good), "ok": int(sum(r.ok for r in res)), "gcells": tm.dp_cells / 1e9, "gpu_ms": tm.dp_ms,
               "gcups": g, "roofline_frac": g / peak}
        if lib is not None:
            n_cpu = min(len(good), 4 * cores)
            row["reference_gcups"] = cpu_gcups(good[:n_cpu], w)
            row["reference_pairs"] = n_cpu
        rows.append(row)
    return {"workload": f"C5 sample: {n_pairs} of BASELINE's 1 M pairs (lengths log-uniform 1-50 kb, target = query at 12 % error), every band "
                        "over all pairs whose length difference fits it", "reference_threads": cores if lib is not None else 0, "bands": rows}


def c3_c4_block(capi, device: int, seed: int, n_reads: int):
    """BASELINE configs[2] / configs[3] at the stage where they differ from configs[1]: the index of the 24-contig 3.1 Gbp reference
    (replicated per GPU in a multi-GPU run) and the mapping stage against it -- at 3.1 Gbp a 50 kb read draws ~16,500 random seed matches
    instead of ~270.  C3: k 15 / w 10, 50 kb reads at 10-15 %; C4: k 19 / w 20, ultra-long reads cut into 50 kb
    subreads with a third at 20-30 % error.  The DP stages behind the mapper are the kernels the headline measures; whole-path lines of
    both shapes (`--workload c3|c4`, 1 and 8 GPUs) are under profiles/."""
    t0 = time.perf_counter()
    contigs = make_c3_reference(seed, 1.0)
    pac, offs = pack_pac(contigs)
    sizes = [len(c) for c in contigs]
    gen_s = time.perf_counter() - t0
    ctx = capi.Context(device)
    capi.subject_load(ctx, pac)
    out = {"reference": {"contigs": len(sizes), "bases": int(sum(sizes)), "pac_bytes": int(pac.size), "host_generation_s": gen_s}}
    for name, k, w, gen in (("c3", 15, 10, make_reads_c3), ("c4", 19, 20, make_reads_c4)):
        t0 = time.perf_counter()
        capi.index_build(ctx, [int(x) for x in offs], sizes, k, w, 200)
        build_ms = (time.perf_counter() - t0) * 1e3
        B = Batch(gen(contigs, n_reads, seed + 900), capi)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            h = capi.map_batch_raw(ctx, B.query_pac, B.seq_arr, 0, B.nq, hitlist_size=5, max_hsps=10)
            wall = (time.perf_counter() - t0) * 1e3
            mt, et = capi.map_last_timing(ctx), capi.map_last_extend_timing(ctx)
            dev = mt.lookup_ms + mt.emit_ms + mt.sort_ms + mt.chain_ms + et.ihf_ms + et.sort_ms + et.chain_ms + et.trace_ms
            if best is None or dev < best["device_ms"]:
                best = {"device_ms": dev, "wall_ms": wall, "seed_chain_ms": mt.lookup_ms + mt.emit_ms + mt.sort_ms + mt.chain_ms,
                        "extend_ms": et.ihf_ms + et.sort_ms + et.chain_ms + et.trace_ms, "kmer_matches": int(mt.n_matches)}
        bases = int((h["qend"] - h["qoff"]).sum())
        out[name] = {"kmer_size": k, "kmer_window": w, "index_build_ms": build_ms, "reads": B.nq, "read_bases": int(B.qsizes.sum()),
                     "hsps": len(h), "aligned_bases": bases, "matches_per_read": best["kmer_matches"] / max(1, B.nq),
                     "map_device_ms": best["device_ms"], "map_wall_ms": best["wall_ms"], "seed_chain_ms": best["seed_chain_ms"],
                     "extend_ms": best["extend_ms"], "map_aligned_bases_per_s": bases / (best["device_ms"] / 1e3),
                     "map_read_bases_per_s": float(B.qsizes.sum()) / (best["device_ms"] / 1e3),
                     "n_unsupported": int(capi.map_last_extend_timing(ctx).n_unsupported)}
    ctx.close()
    return out


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=2368)
    ap.add_argument("--batches", type=int, default=4, help="distinct read batches cycled through the steps")
    ap.add_argument("--ref-len", type=int, default=REF_LEN)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2 = BASELINE configs[1] (headline); c3 = configs[2]: 24 contigs / 3.1 Gbp (x --c3-scale), index replicated per GPU; "
                         "c4 = configs[3]: the c3 reference, ultra-long reads split into 50 kb subreads, 35 %% of them at 20-30 %% error, "
                         "-kmer_size 19 -kmer_window 20, SVR identity 80.0")
    ap.add_argument("--c3-scale", type=float, default=1.0)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank has its own batches; strong: one fixed read set, sub-batches pulled by all ranks from a shared queue")
    ap.add_argument("--extra", default="c5,c3", help="comma list of extra blocks on rank 0 at N=1: c5 (band sweep sample), c3 (index + mapping "
                                                      "stage against the 3.1 Gbp reference with the C3 and the C4 settings), none")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    c4 = args.workload == "c4"
    c3 = args.workload == "c3" or c4                    # the 24-contig reference
    kmer, window, svr_ident = (19, 20, 80.0) if c4 else (15, 10, 70.0)
    if c4:
        what = ("C4: 24 contigs with hg19's length profile, %.2f Gbp; ultra-long reads (150 kb-2.4 Mbp) cut into 50 kb subreads, 35 %% of "
                "the source reads at 20-30 %% error, the others at 8-15 %%, INS/DEL of 40 bp-10 kb about every 250 kb" % (3.098 * args.c3_scale))
    elif c3:
        what = ("C3: 24 contigs with hg19's length profile, %.2f Gbp, 50 kb subreads at 10-15 %% error with planted 40 bp-10 kb INS/DEL "
                "(every 4th read)" % (3.098 * args.c3_scale))
    else:
        what = ("C2: 50 Mbp synthetic contig, 50 kb subreads at 10-15 % error with planted 40 bp-10 kb INS/DEL (~1 per 250 kb of the "
                "sample genome)")
    config = {"workload": what + f"; mapping + SvRead selection + SV signatures (-max_target_seqs 5 -max_hsps 10 -kmer_size {kmer} "
                                 f"-kmer_window {window}, SVR 3000/{svr_ident}/300, min indel 40)",
              "reads_per_step_per_gpu": args.reads_per_step, "read_len": READ_LEN,
              "ref_len": int(sum(HG19_MBP) * 1e6 * args.c3_scale) if c3 else args.ref_len,
              "batches": (f"a pool of {args.batches} distinct read batches; steps x {args.batches} sub-batches for the whole job, pulled by all ranks from a shared queue"
                          if args.scaling == "strong" else f"a pool of {args.batches} distinct read batches, cycled step by step by every GPU from its own "
                                                           "starting point (warm-up included): equal work per GPU when steps is a multiple of the pool size"),
              "l2": "every step maps another batch and writes > 100 GB of traceback: far beyond the 126 MB L2",
              "seed": args.seed}

    # ---- reference arm: the reference's own CPU tools on the host cores --------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        if c3:
            _emit({"impl": "reference", "unavailable": "the reference arm runs the C2 workload (its index build for 3.1 Gbp alone takes minutes)"})
            return
        ref, sample = make_genomes(args.seed, args.ref_len)
        n_sample = sample_size(cores)
        K = max(1, min(args.steps, 2))
        vals = []
        base = None
        for k in range(K):
            base = cpu_baseline(ref, make_reads(sample, n_sample, args.seed + 77 + k), cores)
            if base is None:
                _emit({"impl": "reference", "unavailable": "oracle/_ref/bin is missing (make -C oracle ref needs /root/reference)"})
                return
            vals.append(base)
        v = sum(b["value"] for b in vals) / len(vals)
        sec = sum(sum(b["stages_s"].values()) for b in vals) / len(vals)
        base = dict(vals[-1], value=v)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "bases/s", "n_gpus": args.gpus, "steps": K, "warmup": 0,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int8", "data": "synthetic", "config": dict(config, reads_per_step_per_gpu=n_sample, batches=f"{K} samples"),
                "cpu_baseline": base,
                "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _emit(line)
        return

    # ---- our arm ------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    from lesv_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # keep NCCL's banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    strong = args.scaling == "strong"
    n_batches = max(1, args.batches)
    if c3:
        contigs = make_c3_reference(args.seed, args.c3_scale)                    # replicated on every GPU
        ref = sample = None
    else:
        ref, sample = make_genomes(args.seed, args.ref_len)
        contigs = [ref]
    subject_pac, s_offs = pack_pac(contigs)
    subj = {"offsets": [int(x) for x in s_offs], "sizes": [len(c) for c in contigs]}
    subj["seq_arr"] = capi._seq_array(subj["offsets"], subj["sizes"])

    def reads_of(seed, n):
        return make_reads_c4(contigs, n, seed) if c4 else make_reads_c3(contigs, n, seed) if c3 else make_reads(sample, n, seed)
    # One pool of read batches, the same on every rank.  weak: every rank cycles through the whole pool, each from its own
    # starting point, so all ranks do the same work (the slowest rank's time then shows the system, not its draw of reads: with
    # a private read set per rank the spread between draws alone cost 4-5 % at 8 GPUs); strong: steps x batches sub-batches
    # for the whole job, pulled by all ranks from a shared queue.
    batches = [Batch(reads_of(args.seed + 500 + b, args.reads_per_step), capi) for b in range(n_batches)]
    if c3:
        subject_bases = sum(subj["sizes"])
        del contigs

    ctx = capi.Context(local_rank)
    lib = ctx.lib
    lib.lesvgpu_launch_count.restype = C.c_uint64
    capi.subject_load(ctx, subject_pac)
    t0 = time.perf_counter()
    capi.index_build(ctx, subj["offsets"], subj["sizes"], kmer, window, 200)
    index_ms = (time.perf_counter() - t0) * 1e3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        run_path(ctx, capi, batches[(rank + k) % len(batches)], subj, svr_ident=svr_ident)
    barrier()
    acc = None
    n_done = 0
    launches0 = lib.lesvgpu_launch_count()
    queue = shard.WorkQueue(args.steps * n_batches, world) if strong else None
    with ClockSampler(local_rank, active=(rank == 0)) as clk:
        t0 = time.perf_counter()
        k = 0
        while True:
            if strong:
                item = queue.next()
                if item is None:
                    break
                B = batches[item % len(batches)]
            else:
                if k >= args.steps:
                    break
                B = batches[(rank + args.warmup + k) % len(batches)]
            k += 1
            r = run_path(ctx, capi, B, subj, svr_ident=svr_ident)
            n_done += 1
            if acc is None:
                acc = {key: (dict(v) if isinstance(v, dict) and key in ("dev", "wall") else v) for key, v in r.items()}
                acc["class_sum"], acc["class_sum_svr"] = {}, {}
            else:
                for key in ("dp_ms", "sig_dp_ms", "sig_cells", "svr_dp_ms", "svr_cells", "cells", "dp_launches", "sig_launches", "bases",
                            "n_unsupported"):
                    acc[key] += r[key]
                for key in acc["dev"]:
                    acc["dev"][key] += r["dev"][key]
                for key in acc["wall"]:
                    acc["wall"][key] += r["wall"][key]
            for stage, key in (("svsig", "class_sum"), ("svr_last_round", "class_sum_svr")):
                for c in r["classes"][stage]:
                    name = ("strips" if c["ctas_per_alignment"] < 0 else "global-memory fallback" if c["ctas_per_alignment"] == 0 else
                            f"{c['ctas_per_alignment']}-CTA cluster" if c["ctas_per_alignment"] > 1 else
                            f"{32 // c['threads']} alignments per warp" if c["threads"] <= 16 else f"single CTA, {c['threads']} threads")
                    e = acc.setdefault(key, {}).setdefault(name, {"alignments": 0, "cells": 0, "done_ms": 0.0, "ctas": c["n_ctas"]})
                    e["alignments"] += c["n_alignments"]; e["cells"] += c["cells"]; e["done_ms"] += c["done_ms"]
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
    launches = lib.lesvgpu_launch_count() - launches0
    if acc is None:                                     # strong scaling with more ranks than sub-batches
        acc = {"dev": {"map_seed_chain": 0.0, "map_extend": 0.0, "svr": 0.0, "svsig": 0.0}, "wall": {}, "bases": 0, "cells": 0, "dp_ms": 0.0,
               "sig_dp_ms": 0.0, "sig_cells": 0, "svr_dp_ms": 0.0, "svr_cells": 0, "sig_launches": 0, "n_unsupported": 0, "class_sum": {}}
    dev_ms = sum(acc["dev"].values())
    if world > 1:                                       # per-rank view for diagnosis (stderr; rank 0's JSON line is the result)
        print(f"[bench rank {rank}] {n_done} batches, device ms/batch " + ", ".join(f"{k} {v / max(1, n_done):.1f}" for k, v in acc["dev"].items()) +
              f"; wall {wall_ms / max(1, n_done):.1f}; clocks {clk.summary()}", file=sys.stderr, flush=True)

    dev_ms_max, wall_ms_max, all_bases, all_cells, all_launches = shard.reduce_stats(
        dev_ms, wall_ms, float(acc["bases"]), float(acc["cells"]), float(launches), world, device="cuda")
    dev_hi, dev_mean, dev_lo = shard.spread(dev_ms, world, device="cuda")

    if rank == 0:
        K = args.steps * n_batches / world if strong else args.steps       # steps per rank
        clocks = clk.summary()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        f_mhz = clocks["sm_mhz"] or float(peaks.get("sm_max_mhz", 1965.0))
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        int_peak = n_sm * 128 * f_mhz * 1e6 / OPS_PER_CELL * LANES_PACKED / 1e9      # GCUPS, SURVEY.md 8d
        kern_gcups = acc["sig_cells"] / (max(acc["sig_dp_ms"], 1e-9) / 1e3) / 1e9    # rank 0's qx2svsig DP launches
        def class_table(cs):
            return {name: {"alignments_per_step": e["alignments"] / max(1, n_done), "gcells_per_step": e["cells"] / max(1, n_done) / 1e9,
                           "ctas": e["ctas"], "done_ms_per_step": e["done_ms"] / max(1, n_done),
                           "gcups_over_done_ms": e["cells"] / max(e["done_ms"], 1e-9) / 1e6} for name, e in cs.items()}
        classes = class_table(acc.get("class_sum", {}))
        line = {
            "metric": METRIC.replace("C2", "C4") if c4 else METRIC_C3 if c3 else METRIC, "value": all_bases / (dev_ms_max / 1e3), "unit": "bases/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / K,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int8",
            "data": "synthetic", "config": config,
            "gcups": all_cells / (acc["dp_ms"] / 1e3) / 1e9 if world == 1 else None,
            "e2e": {"value": all_bases / (wall_ms_max / 1e3), "unit": "bases/s", "h2d_bytes_per_step": r["h2d"] if n_done else 0,
                    "d2h_bytes_per_step": r["d2h"] if n_done else 0, "wall_ms_per_step": wall_ms_max / K},
            "gpu_launches": int(all_launches),
            "clocks": clocks,
            "rank_spread": {"dev_ms_max": dev_hi, "dev_ms_mean": dev_mean, "dev_ms_min": dev_lo,
                            "note": "summed device time of the timed steps per rank; the job is as slow as dev_ms_max"},
            "stages_ms_per_step": {k: v / max(1, n_done) for k, v in acc["dev"].items()},
            "call_wall_ms_per_step": {k: v / max(1, n_done) for k, v in acc["wall"].items()},
            "stage_throughput": {"map_bases_per_s": acc["bases"] / (max(acc["dev"]["map_seed_chain"] + acc["dev"]["map_extend"], 1e-9) / 1e3),
                                 "svr_gcups": acc["svr_cells"] / (max(acc["svr_dp_ms"], 1e-9) / 1e3) / 1e9,
                                 "svsig_gcups": kern_gcups,
                                 "svr_svsig_gcups": acc["cells"] / (max(acc["dp_ms"], 1e-9) / 1e3) / 1e9},
            "svsig_launch_classes": classes,
            "svr_launch_classes": class_table(acc.get("class_sum_svr", {})),
            "roofline": {"bound": "int-issue", "kernel": "ksw_dp_overlap_kernel (single-CTA classes; ksw_dp_kernel for cluster classes)", "achieved": kern_gcups, "peak": int_peak,
                         "unit": "GCUPS", "frac": kern_gcups / int_peak,
                         "peak_def": f"{n_sm} SMs x 128 lanes x {f_mhz:.0f} MHz (median under load) / {OPS_PER_CELL} ops per cell x "
                                     f"{LANES_PACKED} cells per 32-bit op (SURVEY.md 8d)",
                         "avg_launch_ms": acc["sig_dp_ms"] / max(1, acc["sig_launches"]), "launches": int(acc["sig_launches"]),
                         "launch_note": "the launch classes of one batch run concurrently (chained by programmatic dependent launch); achieved = "
                                        "cells of the qx2svsig batch / CUDA-event time around all of its K-ksw launches; per class: svsig_launch_classes",
                         # profiles/r2_ncu_summary.md: 203.4 GB of DRAM traffic for 197.75 G cells in one ncu --set full capture
                         "traffic": 1.028 * acc["sig_cells"] / max(1, acc["sig_launches"]) / 1e9, "traffic_unit": "GB per launch",
                         "traffic_note": "dram__bytes_read+write per cell from the ncu capture (1.028 B/cell) x cells per launch; algorithmic = 1 B/cell",
                         "hbm": {"achieved": kern_gcups, "peak": hbm_peak, "unit": "GB/s", "frac": kern_gcups / hbm_peak,
                                 "note": "1 traceback byte per DP cell, written once (of measured copy bandwidth)"}},
            "index_build_ms": index_ms,
            "n_unsupported": int(acc["n_unsupported"]),
            "per_step": {k: r[k] for k in ("hsps", "sv_reads", "svr_jobs", "reads_ok", "signatures")} if n_done else {},
        }
        if c3:
            line["index"] = {"subject_bases": int(subject_bases), "contigs": len(subj["sizes"]), "build_ms": index_ms}
        if not args.no_cpu_baseline and world == 1 and not c3:
            n_sample = sample_size(cores)
            keep = {}
            sample_reads = batches[0].reads[:n_sample]
            base = cpu_baseline(ref, sample_reads, cores, keep)
            if base is not None:
                line["cpu_baseline"] = base
                ours = run_path(ctx, capi, Batch(sample_reads, capi), subj, collect=True)      # outside the timed region
                line["parity"] = parity_block(ours["records"], keep)
        extras = [x for x in args.extra.split(",") if x and x != "none"]
        if world == 1 and "c5" in extras:
            line.setdefault("extra", {})["c5"] = c5_block(ctx, capi, cores, args.seed, n_sm, f_mhz)
        if world == 1 and "c3" in extras and not c3:
            ctx.close()                                  # the 3.1 Gbp index wants the memory the headline context holds
            ctx = None
            line.setdefault("extra", {})["c3_c4_mapping"] = c3_c4_block(capi, local_rank, args.seed, 1184)
        _emit(line)
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
