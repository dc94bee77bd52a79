#!/usr/bin/env python
"""bench.py -- headline benchmark of the lesv hot path on B200 (contract: see the task statement).

Metric (BASELINE.json): GCUPS of the banded dual-affine DP (cells as ksw_extd2_sse counts them,
reference src/algo/ksw2_extd2_sse.c:124-138) and aligned bases/s, per B200, on the SV-signature
stage (qx2svsig) that consumes the mapper's alignments -- ~90 % of the reference's CPU time on
this path.  Workload = BASELINE.json configs[1] shape: 50 kb subreads at 10-15 %
ONT-like error against a 50 Mbp synthetic contig with planted 40 bp-10 kb INS/DEL; one *step* is
one lesvgpu_svsig_batch call (the reference-facing C ABI; discover_func's body) over a batch of
SvReads with HOST buffers.

  value      whole-job GCUPS with inputs resident (pack + DP/traceback kernels, CUDA events)
  e2e        the same through the C ABI including H2D of the reads' .pac bytes and D2H of results
  roofline   integer-issue roofline of ksw_dp_kernel, plus the HBM view
  cpu_baseline / --impl reference : the reference's own nw_ksw2_extd2 (oracle/_ref/libhbn_ref.so,
             compiled from the unmodified sources) on all host cores, bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from lesv_b200 import synth  # noqa: E402

READ_LEN = 50000
REF_LEN = 50_000_000
MIN_INDEL = 40
OPS_PER_CELL = 46            # scalar formulation of ksw2_extd2_sse.c:31-58,227-265,318-349
LANES_PACKED = 2             # two DP cells per 32-bit integer instruction (16-bit lanes)


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


def make_workload(n_reads: int, seed: int, read_len: int = READ_LEN, ref_len: int = REF_LEN, sv_frac: float = 0.2):
    """C2-shaped SvRead batch: (reference contig, reads, SvRead tuples)."""
    rng = np.random.default_rng(seed)
    ref = synth.random_codes(ref_len, rng)
    reads, svreads, pairs = [], [], []
    for i in range(n_reads):
        s = int(rng.integers(0, ref_len - read_len - 12000))
        tgt_len = read_len
        sample = ref[s:s + tgt_len]
        sv = 0
        if rng.random() < sv_frac:                   # ~1 event per 250 kb of read
            size = int(np.exp(rng.uniform(np.log(40), np.log(10000))))
            pos = int(rng.integers(2000, read_len - 2000))
            if rng.random() < 0.5:                   # insertion in the sample
                sample = np.concatenate([sample[:pos], synth.random_codes(size, rng), sample[pos:]])
                sv = size
            else:                                    # deletion in the sample: the window on the reference grows
                tgt_len = read_len + size
                sample = np.concatenate([ref[s:s + pos], ref[s + pos + size:s + tgt_len]])
                sv = size
        err = float(rng.uniform(0.10, 0.15))
        read = synth.mutate(sample, err, rng)
        qdir = int(rng.integers(0, 2))
        fwd = synth.revcomp_codes(read) if qdir else read
        qlen = len(read)
        # identity as the mapper would report it (matches / alignment columns), 6 significant digits
        cols = len(sample) * (1 + 0.3 * err) + sv
        ident = float(f"{100.0 * len(sample) * (1 - 0.7 * err) / cols:.6g}")
        dist = int((qlen + tgt_len) * (100.0 - ident) / 200.0)       # find_sv_reads.c:49
        reads.append(np.ascontiguousarray(fwd))
        svreads.append((i, 0, qdir, 0, qlen, qlen, s, s + tgt_len, dist))
        pairs.append((np.ascontiguousarray(read), s, tgt_len, dist))
    return ref, reads, svreads, pairs


def band_of(qn, sn, dist):
    """discover_func's band policy (find_sv_signature.c:272-316)."""
    w = int(dist * 1.2)
    if w == 0:
        w = int(max(qn, sn) * 0.1)
    if abs(qn - sn) > w:
        w = max(qn, sn)
    return w


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, gpu_index: int):
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
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        hi = [x for x in s if self.smax and x > 0.5 * self.smax] or s
        return {"sm_mhz": (hi[len(hi) // 2] if hi else None), "sm_max_mhz": self.smax, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the compiled reference's nw_ksw2_extd2 on all host cores
# ---------------------------------------------------------------------------------------------
class KString(C.Structure):
    _fields_ = [("l", C.c_size_t), ("m", C.c_size_t), ("s", C.c_void_p)]


def _cpu_lib():
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libhbn_ref.so")
    if os.path.exists(ref_so):
        lib = C.CDLL(ref_so)
        lib.Ksw2DataNew.restype = C.c_void_p
        lib.ksw2_extd2_set_params.argtypes = [C.c_void_p]
        lib.nw_ksw2_extd2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 7
        return "reference", lib
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    return "port", C.CDLL(os.path.join(ROOT, "oracle", "liblesv_oracle.so"))


def cpu_align_pairs(kind, lib, jobs, n_threads):
    """jobs: list of (query codes, target codes, band). Runs them on n_threads host threads
    (ctypes releases the GIL). Returns wall seconds."""
    it = iter(range(len(jobs)))
    lock = threading.Lock()

    def worker():
        if kind == "reference":
            data = lib.Ksw2DataNew()
            lib.ksw2_extd2_set_params(data)
            qa, sa = KString(), KString()
            ints = [C.c_int() for _ in range(4)]
            ident = C.c_double()
        while True:
            with lock:
                i = next(it, None)
            if i is None:
                return
            q, t, w = jobs[i]
            if kind == "reference":
                lib.nw_ksw2_extd2(data, 0, q.ctypes.data, 0, len(q), len(q), 0, t.ctypes.data, 0, len(t), len(t), 0,
                                  C.c_double(0.0), w, *[C.addressof(x) for x in ints], C.addressof(ident),
                                  C.addressof(qa), C.addressof(sa))
            else:
                outs = [C.c_int() for _ in range(4)]
                ident2, n = C.c_double(), C.c_int()
                pa, ps = C.c_void_p(), C.c_void_p()
                lib.lesv_oracle_nw_ksw2_extd2(q.ctypes.data, 0, len(q), t.ctypes.data, 0, len(t), C.c_double(0.0), w,
                                              *[C.byref(x) for x in outs], C.byref(ident2), C.byref(pa), C.byref(ps), C.byref(n))

    ths = [threading.Thread(target=worker) for _ in range(n_threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    return time.perf_counter() - t0


def cpu_jobs_from(ref, pairs, k):
    jobs = []
    for read, s, tl, dist in pairs[:k]:
        t = np.ascontiguousarray(ref[s:s + tl])
        jobs.append((read, t, band_of(len(read), tl, dist)))
    return jobs


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=1184)
    ap.add_argument("--no-sv", action="store_true", help="profiling aid: no planted INS/DEL (no cluster launches)")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    config = {"workload": "C2: 50 Mbp synthetic contig, 50 kb subreads at 10-15 % error, planted 40 bp-10 kb INS/DEL; "
                          "qx2svsig stage (banded dual-affine DP + traceback + signature scan)",
              "reads_per_step_per_gpu": args.reads_per_step, "read_len": READ_LEN, "min_indel": MIN_INDEL,
              "l2": "inputs and traceback per step far exceed the 126 MB L2", "seed": args.seed}

    # ---- reference arm: the reference's own CPU implementation on the host cores ------------------
    if args.impl == "reference":
        if rank != 0:
            return
        kind, lib = _cpu_lib()
        n_sample = max(cores, 8)
        ref, reads, svreads, pairs = make_workload(n_sample, args.seed, ref_len=4_000_000)
        jobs = cpu_jobs_from(ref, pairs, n_sample)
        from lesv_b200 import capi
        cells = sum(capi.ksw_cells(len(q), len(t), w)[0] for q, t, w in jobs)
        bases = sum(len(q) for q, _, _ in jobs)
        for _ in range(min(args.warmup, 1)):
            cpu_align_pairs(kind, lib, jobs[:cores], cores)
        times = [cpu_align_pairs(kind, lib, jobs, cores) for _ in range(max(1, min(args.steps, 3)))]
        sec = sum(times) / len(times)
        gcups = cells / sec / 1e9
        line = {"impl": "reference", "metric": "GCUPS (banded dual-affine DP cells/s, qx2svsig stage)", "value": gcups,
                "unit": "GCUPS", "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 1),
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int8", "data": "synthetic", "config": config,
                "aligned_bases_per_s": bases / sec,
                "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": kind,
                                 "sample": f"{n_sample} reads of the same workload, nw_ksw2_extd2 on {cores} threads"},
                "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---- our arm ------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    from lesv_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ref, reads, svreads, pairs = make_workload(args.reads_per_step, args.seed + 1000 * rank,
                                               sv_frac=0.0 if args.no_sv else 0.2)
    subject_pac, _ = pack_pac([ref])
    query_pac, qoffs = pack_pac(reads)
    read_offset = (qoffs + 1).astype(np.uint64)
    bands = [band_of(r[4] - r[3], r[7] - r[6], r[8]) for r in svreads]
    cells = sum(capi.ksw_cells(r[4] - r[3], r[7] - r[6], w)[0] for r, w in zip(svreads, bands))
    bases = sum(r[4] - r[3] for r in svreads)
    tb_bytes = cells                                   # 1 traceback byte per cell, written once

    ctx = capi.Context(local_rank)
    capi.subject_load(ctx, subject_pac)

    def step():
        per, sigs = capi.svsig_batch(ctx, query_pac, read_offset, 0, len(ref), svreads, MIN_INDEL)
        return per, sigs, ctx.timing()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    dev_ms = e2e_ms = dp_ms = 0.0
    launches = dp_launches = 0
    h2d = d2h = 0
    n_sig = 0
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            per, sigs, tm = step()
            dev_ms += tm.pack_ms + tm.dp_ms + tm.scan_ms
            e2e_ms += tm.h2d_ms + tm.pack_ms + tm.dp_ms + tm.scan_ms + tm.d2h_ms
            dp_ms += tm.dp_ms
            launches += tm.kernel_launches
            dp_launches += tm.dp_launches
            h2d, d2h = tm.h2d_bytes, tm.d2h_bytes
            n_sig = len(sigs)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
    n_ok = sum(1 for i in range(len(svreads)) if per[i].status == 1)

    stats = torch.tensor([dev_ms, e2e_ms, wall_ms, dp_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(cells), float(bases)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, wall_ms, dp_ms_max = [float(x) for x in stats.tolist()]
    all_cells, all_bases = [float(x) for x in tot.tolist()]

    if rank == 0:
        K = args.steps
        value = all_cells * K / (dev_ms / 1e3) / 1e9
        e2e_val = all_cells * K / (wall_ms / 1e3) / 1e9
        clocks = clk.summary()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        f_mhz = clocks["sm_mhz"] or float(peaks.get("sm_max_mhz", 1965.0))
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        int_peak = n_sm * 128 * f_mhz * 1e6 / OPS_PER_CELL * LANES_PACKED / 1e9      # GCUPS, SURVEY.md §8d
        kern_gcups = cells * K / (dp_ms / 1e3) / 1e9                                # rank 0's kernel
        line = {
            "metric": "GCUPS (banded dual-affine DP cells/s, qx2svsig stage)", "value": value, "unit": "GCUPS",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": dev_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic", "config": config,
            "aligned_bases_per_s": all_bases * K / (dev_ms / 1e3),
            "e2e": {"value": e2e_val, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "device_event_value": all_cells * K / (e2e_ms / 1e3) / 1e9, "wall_ms_per_step": wall_ms / K,
                    "aligned_bases_per_s": all_bases * K / (wall_ms / 1e3)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "int-issue", "kernel": "ksw_dp_kernel", "achieved": kern_gcups, "peak": int_peak,
                         "unit": "GCUPS", "frac": kern_gcups / int_peak,
                         "peak_def": f"{n_sm} SMs x 128 lanes x {f_mhz:.0f} MHz (median under load) / {OPS_PER_CELL} ops per cell x "
                                     f"{LANES_PACKED} cells per 32-bit op (SURVEY.md 8d)",
                         "avg_launch_ms": dp_ms / max(1, dp_launches), "launches": int(dp_launches),
                         "traffic": None,
                         "hbm": {"achieved": tb_bytes * K / (dp_ms / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": tb_bytes * K / (dp_ms / 1e3) / 1e9 / hbm_peak,
                                 "note": "1 traceback byte per DP cell, written once (of measured copy bandwidth)"}},
            "reads_ok": n_ok, "signatures_per_step": n_sig,
        }
        if not args.no_cpu_baseline and world == 1:
            kind, lib = _cpu_lib()
            n_sample = max(cores, 8)
            jobs = cpu_jobs_from(ref, pairs, n_sample)
            c_cells = sum(capi.ksw_cells(len(q), len(t), w)[0] for q, t, w in jobs)
            sec = cpu_align_pairs(kind, lib, jobs, cores)
            line["cpu_baseline"] = {"value": c_cells / sec / 1e9, "unit": "GCUPS", "cores": cores, "kind": kind,
                                    "sample": f"first {n_sample} reads of the step, nw_ksw2_extd2 on {cores} threads, {sec:.1f} s"}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
