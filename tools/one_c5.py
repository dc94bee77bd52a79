"""One narrow-band K-ksw batch (BASELINE configs[4] length mix) for ncu captures: python tools/one_c5.py [band] [pairs]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi, synth
band = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
rng = np.random.default_rng(7)
lens = np.exp(rng.uniform(np.log(1000), np.log(50000), n)).astype(np.int64)
big = synth.random_codes(64_000_000, rng)
noisy, pos = synth.mutate_with_map(big, 0.12, rng)
qs, ts, a = [], [], 0
for L in lens:
    if a + L > len(big):
        a = int(rng.integers(0, len(big) - 50000))
    q, t = big[a:a + L], noisy[pos[a]:pos[a + L]]
    a += int(L)
    if abs(len(q) - len(t)) <= band:
        qs.append(q); ts.append(t)
ctx = capi.Context(0)
res, _ = ctx.ksw_batch(qs, ts, [band] * len(qs), want_cigars=False)
tm = ctx.timing()
print(f"band {band}: {len(qs)} pairs, {tm.dp_cells / 1e9:.2f} Gcells, {tm.dp_ms:.1f} ms -> {tm.dp_cells / tm.dp_ms / 1e6:.1f} GCUPS", flush=True)
print(ctx.class_timing())
