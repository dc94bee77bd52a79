"""Quick qx2map-stage timing on the GPU (development aid; bench.py is the contract).
usage: map_bench.py [n_reads] [read_len] [ref_len] [reps]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi, synth
from bench import pack_pac

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200
read_len = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
ref_len = int(sys.argv[3]) if len(sys.argv) > 3 else 50_000_000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
sv = int(sys.argv[5]) if len(sys.argv) > 5 else 1

rng = np.random.default_rng(7)
ref = synth.random_codes(ref_len, rng)
reads = []
for i in range(n_reads):
    s = int(rng.integers(0, ref_len - read_len))
    seg = ref[s:s + read_len]
    if sv and i % 5 == 0:                      # every 5th read spans a 40 bp - 10 kb insertion or deletion
        size = int(np.exp(rng.uniform(np.log(40), np.log(10000))))
        pos = int(rng.integers(2000, read_len - 2000 - size))
        seg = np.concatenate([seg[:pos], synth.random_codes(size, rng), seg[pos:]]) if i % 10 == 0 else np.concatenate([seg[:pos], seg[pos + size:]])
    r = synth.mutate(seg, float(rng.uniform(0.10, 0.15)), rng)
    if rng.integers(0, 2):
        r = synth.revcomp_codes(r)
    reads.append(np.ascontiguousarray(r))
spac, _ = pack_pac([ref])
qpac, qoffs = pack_pac(reads)
qsizes = np.array([len(r) for r in reads], dtype=np.int64)
bases = int(qsizes.sum())

ctx = capi.Context(0)
capi.subject_load(ctx, spac)
t0 = time.time()
capi.index_build(ctx, [0], [ref_len], 15, 10, 200)
print(f"index build {ref_len/1e6:.0f} Mbp: {(time.time()-t0)*1e3:.0f} ms wall", flush=True)
for rep in range(reps):
    t0 = time.time()
    res = capi.map_batch(ctx, qpac, qoffs, qsizes, 0, n_reads, hitlist_size=5, max_hsps=10)
    wall = time.time() - t0
    mt, et = capi.map_last_timing(ctx), capi.map_last_extend_timing(ctx)
    aligned = sum(h["qend"] - h["qoff"] for l in res for h in l)
    nh = sum(len(l) for l in res)
    print(f"rep {rep}: {n_reads} x {read_len}: wall {wall*1e3:.0f} ms -> {bases/wall/1e6:.1f} Mbases/s (aligned {aligned/wall/1e6:.1f}), hsps {nh}", flush=True)
    print("   cand: " + " ".join(f"{n}={getattr(mt, n):.1f}" for n in ("h2d_ms", "lookup_ms", "emit_ms", "sort_ms", "chain_ms", "d2h_ms")) +
          f" kmers={mt.n_kmers} matches={mt.n_matches} segs={mt.n_segments} hits={mt.n_hits}", flush=True)
    print("   ext:  " + " ".join(f"{n}={getattr(et, n):.1f}" for n in ("ihf_ms", "sort_ms", "chain_ms", "trace_ms", "d2h_ms")) +
          f" windows={et.n_windows} seeds={et.n_seeds} cands={et.n_candidates} unsupported={et.n_unsupported}", flush=True)
    print(f"   trace: prep={et.prep_ms:.1f} gap={et.gap_ms:.1f} finish={et.finish_ms:.1f} gaps={et.n_gap_jobs} classes={list(et.n_gaps_class)}", flush=True)
ctx.close()
