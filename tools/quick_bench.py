"""Quick K-ksw timing on the GPU (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from lesv_b200.synth import mutate

def run(ctx, n, length, err, rng, reps=2, full=False):
    qs, ts, ws = [], [], []
    for _ in range(n):
        t = rng.integers(0, 4, length, dtype=np.uint8)
        q = np.ascontiguousarray(mutate(t, err, rng))
        dist = int((len(q) + len(t)) * (100 - 100 * (1 - err)) / 200)
        ws.append(-1 if full else int(dist * 1.2)); qs.append(q); ts.append(t)
    cells = sum(capi.ksw_cells(len(q), len(t), w)[0] for q, t, w in zip(qs, ts, ws))
    for rep in range(reps):
        t0 = time.time()
        res, _ = ctx.ksw_batch(qs, ts, ws, want_cigars=False)
        wall = time.time() - t0
        tm = ctx.timing()
        print(f"n={n} len={length} err={err} band={ws[0]} cells={cells/1e9:.2f}G ok={sum(r.ok for r in res)} "
              f"dp={tm.dp_ms:.1f}ms -> {cells/tm.dp_ms/1e6:.1f} GCUPS | h2d {tm.h2d_ms:.1f} pack {tm.pack_ms:.2f} d2h {tm.d2h_ms:.1f} wall {wall*1e3:.0f}ms", flush=True)

if __name__ == "__main__":
    ctx = capi.Context(0)
    rng = np.random.default_rng(1)
    run(ctx, 592, 10000, 0.12, rng)
    run(ctx, 148, 50000, 0.10, rng, reps=1)
    run(ctx, 148, 50000, 0.12, rng)
    run(ctx, 148, 50000, 0.15, rng, reps=1)
    run(ctx, 16, 20000, 0.12, rng, reps=1, full=True)
