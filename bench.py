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


def reference_pipeline(ref, reads, cores: int):
    """qx2map -> qx2m4x -> qx2svr -> qx2svsig of the compiled reference (oracle/_ref/bin) on `reads`.
    Returns a dict of seconds per stage, aligned bases and signature count.  The mapping time is the
    reference's own 'mapping query ...' timer (hbn_align_one_volume.c:124,141): database packing and the
    index build are one-off costs per reference and are reported separately."""
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
        n_sig = 0
        for fn in os.listdir(os.path.join(work, "svsig")):
            if fn.endswith(".sv_signature"):
                n_sig += sum(1 for _ in open(os.path.join(work, "svsig", fn)))
        res["aligned_bases"] = bases
        res["signatures"] = n_sig
        res["path_s"] = res["map_s"] + res["qx2svr_wall_s"] + res["qx2svsig_wall_s"]
        return res
    finally:
        shutil.rmtree(work, ignore_errors=True)


def cpu_baseline(ref, reads, cores):
    if not os.path.exists(os.path.join(REF_BIN, "qx2map")):
        return None
    r = reference_pipeline(ref, reads, cores)
    return {"value": r["aligned_bases"] / r["path_s"], "unit": "bases/s", "cores": cores, "kind": "reference",
            "sample": f"{len(reads)} reads of the same workload through oracle/_ref/bin qx2map, qx2svr, qx2svsig on {cores} threads: "
                      f"mapping {r['map_s']:.2f} s (the tool's own timer, without database/index build; whole qx2map "
                      f"{r['qx2map_wall_s']:.2f} s), qx2svr {r['qx2svr_wall_s']:.2f} s, qx2svsig {r['qx2svsig_wall_s']:.2f} s",
            "stages_s": {"map": r["map_s"], "svr": r["qx2svr_wall_s"], "svsig": r["qx2svsig_wall_s"]},
            "signatures": r["signatures"]}


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


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=2368)
    ap.add_argument("--ref-len", type=int, default=REF_LEN)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    config = {"workload": "C2: 50 Mbp synthetic contig, 50 kb subreads at 10-15 % error sampled from a genome with planted "
                          "40 bp-10 kb INS/DEL (~1 per 250 kb); mapping + SvRead selection + SV signatures "
                          "(-max_target_seqs 5 -max_hsps 10, SVR 3000/70.0/300, min indel 40)",
              "reads_per_step_per_gpu": args.reads_per_step, "read_len": READ_LEN, "ref_len": args.ref_len,
              "l2": "each step streams new candidate windows and > 100 GB of traceback: far beyond the 126 MB L2",
              "seed": args.seed}

    # ---- reference arm: the reference's own CPU tools on the host cores --------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        ref, sample = make_genomes(args.seed, args.ref_len)
        n_sample = max(4 * cores, 16)
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
                "dtype": "int8", "data": "synthetic", "config": dict(config, reads_per_step_per_gpu=n_sample),
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

    ref, sample = make_genomes(args.seed, args.ref_len)                          # replicated on every GPU
    reads = make_reads(sample, args.reads_per_step, shard.read_seed_of_rank(args.seed, rank))   # this rank's shard
    subject_pac, _ = pack_pac([ref])
    query_pac, qoffs = pack_pac(reads)
    qsizes = np.array([len(r) for r in reads], dtype=np.int64)
    read_offset = (qoffs + 1).astype(np.uint64)
    seq_arr = capi._seq_array(qoffs, qsizes)
    nq = len(reads)

    ctx = capi.Context(local_rank)
    lib = ctx.lib
    lib.lesvgpu_launch_count.restype = C.c_uint64
    capi.subject_load(ctx, subject_pac)
    t0 = time.perf_counter()
    capi.index_build(ctx, [0], [len(ref)], 15, 10, 200)
    index_ms = (time.perf_counter() - t0) * 1e3

    def step():
        """one batch through the path; returns per-stage device ms, byte counts, work counts"""
        w0 = time.perf_counter()
        h = capi.map_batch_raw(ctx, query_pac, seq_arr, 0, nq, hitlist_size=5, max_hsps=10)
        w1 = time.perf_counter()
        mt, et = capi.map_last_timing(ctx), capi.map_last_extend_timing(ctx)
        m4 = capi.hsps_to_m4(h)
        w2 = time.perf_counter()
        svr, st = capi.svr_batch(ctx, query_pac, read_offset, [0], [len(ref)], m4, 3000, 70.0, 300)
        w3 = time.perf_counter()
        t_svr = ctx.timing() if (st.n_refine_jobs or st.n_chain_jobs) else capi.Timing()
        status, n_sig = capi.svsig_batch_raw(ctx, query_pac, read_offset, 0, len(ref), svr, MIN_INDEL)
        w4 = time.perf_counter()
        t_sig = ctx.timing()
        dev = {"map_seed_chain": mt.lookup_ms + mt.emit_ms + mt.sort_ms + mt.chain_ms,
               "map_extend": et.ihf_ms + et.sort_ms + et.chain_ms + et.trace_ms,
               "svr": t_svr.pack_ms + t_svr.dp_ms + t_svr.scan_ms,
               "svsig": t_sig.pack_ms + t_sig.dp_ms + t_sig.scan_ms}
        wall = {"map": (w1 - w0) * 1e3, "m4": (w2 - w1) * 1e3, "svr": (w3 - w2) * 1e3, "svsig": (w4 - w3) * 1e3}
        return {"dev": dev, "wall": wall, "dp_ms": t_svr.dp_ms + t_sig.dp_ms, "sig_dp_ms": t_sig.dp_ms, "sig_cells": int(t_sig.dp_cells),
                "cells": int(t_svr.dp_cells) + int(t_sig.dp_cells), "dp_launches": t_svr.dp_launches + t_sig.dp_launches,
                "sig_launches": t_sig.dp_launches,
                "h2d": int(mt.h2d_bytes + t_svr.h2d_bytes + t_sig.h2d_bytes),
                "d2h": int(mt.d2h_bytes + et.d2h_bytes + t_svr.d2h_bytes + t_sig.d2h_bytes),
                "bases": int((h["qend"] - h["qoff"]).sum()), "hsps": len(h), "sv_reads": len(svr),
                "reads_ok": int((status == 1).sum()), "signatures": n_sig,
                "svr_jobs": int(st.n_refine_jobs + st.n_chain_jobs)}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    acc = None
    launches0 = lib.lesvgpu_launch_count()
    with ClockSampler(local_rank, active=(rank == 0)) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            r = step()
            if acc is None:
                acc = r
            else:
                for k in ("dp_ms", "sig_dp_ms", "sig_cells", "cells", "dp_launches", "sig_launches", "bases"):
                    acc[k] += r[k]
                for k in acc["dev"]:
                    acc["dev"][k] += r["dev"][k]
                for k in acc["wall"]:
                    acc["wall"][k] += r["wall"][k]
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
    launches = lib.lesvgpu_launch_count() - launches0
    dev_ms = sum(acc["dev"].values())
    if world > 1:                                       # per-rank view for diagnosis (stderr; rank 0's JSON line is the result)
        print(f"[bench rank {rank}] device ms/step " + ", ".join(f"{k} {v / args.steps:.1f}" for k, v in acc["dev"].items()) +
              f"; wall {wall_ms / args.steps:.1f}; clocks {clk.summary()}", file=sys.stderr, flush=True)

    dev_ms_max, wall_ms_max, all_bases, all_cells, all_launches = shard.reduce_stats(
        dev_ms, wall_ms, float(acc["bases"]), float(acc["cells"]), float(launches), world, device="cuda")

    if rank == 0:
        K = args.steps
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
        kern_gcups = acc["sig_cells"] / (acc["sig_dp_ms"] / 1e3) / 1e9               # rank 0's qx2svsig DP kernel launches
        line = {
            "metric": METRIC, "value": all_bases / (dev_ms_max / 1e3), "unit": "bases/s",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": dev_ms_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic", "config": config,
            "gcups": all_cells / (acc["dp_ms"] / 1e3) / 1e9 if world == 1 else None,
            "e2e": {"value": all_bases / (wall_ms_max / 1e3), "unit": "bases/s", "h2d_bytes_per_step": r["h2d"],
                    "d2h_bytes_per_step": r["d2h"], "wall_ms_per_step": wall_ms_max / K},
            "gpu_launches": int(all_launches),
            "clocks": clocks,
            "stages_ms_per_step": {k: v / K for k, v in acc["dev"].items()},
            "call_wall_ms_per_step": {k: v / K for k, v in acc["wall"].items()},
            "stage_throughput": {"map_bases_per_s": acc["bases"] / ((acc["dev"]["map_seed_chain"] + acc["dev"]["map_extend"]) / 1e3),
                                 "svr_svsig_gcups": acc["cells"] / (acc["dp_ms"] / 1e3) / 1e9},
            "roofline": {"bound": "int-issue", "kernel": "ksw_dp_overlap_kernel (single-CTA classes; ksw_dp_kernel for cluster classes)", "achieved": kern_gcups, "peak": int_peak,
                         "unit": "GCUPS", "frac": kern_gcups / int_peak,
                         "peak_def": f"{n_sm} SMs x 128 lanes x {f_mhz:.0f} MHz (median under load) / {OPS_PER_CELL} ops per cell x "
                                     f"{LANES_PACKED} cells per 32-bit op (SURVEY.md 8d)",
                         "avg_launch_ms": acc["sig_dp_ms"] / max(1, acc["sig_launches"]), "launches": int(acc["sig_launches"]),
                         "launch_note": "the launch classes of one batch run concurrently on side streams; achieved = cells of the "
                                        "qx2svsig batch / CUDA-event time around all of its ksw_dp_kernel launches",
                         # profiles/r1_ncu_summary.md: 406.8 GB of DRAM traffic for 395.5 G cells in one ncu --set full capture
                         "traffic": 1.029 * acc["sig_cells"] / max(1, acc["sig_launches"]) / 1e9, "traffic_unit": "GB per launch",
                         "traffic_note": "dram__bytes_read+write per cell from the ncu capture (1.029 B/cell) x cells per launch; algorithmic = 1 B/cell",
                         "hbm": {"achieved": kern_gcups, "peak": hbm_peak, "unit": "GB/s", "frac": kern_gcups / hbm_peak,
                                 "note": "1 traceback byte per DP cell, written once (of measured copy bandwidth)"}},
            "index_build_ms": index_ms,
            "per_step": {k: r[k] for k in ("hsps", "sv_reads", "svr_jobs", "reads_ok", "signatures")},
        }
        if not args.no_cpu_baseline and world == 1:
            n_sample = max(4 * cores, 16)
            base = cpu_baseline(ref, reads[:n_sample], cores)
            if base is not None:
                line["cpu_baseline"] = base
        _emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
