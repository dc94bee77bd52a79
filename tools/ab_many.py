"""A/B timing of several builds on the same inputs in one process (development aid):
python tools/ab_many.py [--n 296] [--len 50000] [--err 0.12] a.so b.so ...   (a missing result check is fine for ablation builds)"""
import sys, os, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from lesv_b200.synth import mutate
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=296); ap.add_argument("--len", type=int, default=50000)
ap.add_argument("--err", type=float, default=0.12); ap.add_argument("--reps", type=int, default=2)
ap.add_argument("libs", nargs="+")
a = ap.parse_args()
rng = np.random.default_rng(1)
qs, ts, ws = [], [], []
for _ in range(a.n):
    t = rng.integers(0, 4, a.len, dtype=np.uint8)
    q = np.ascontiguousarray(mutate(t, a.err, rng))
    ws.append(int(int((len(q) + len(t)) * (100 - 100 * (1 - a.err)) / 200) * 1.2)); qs.append(q); ts.append(t)
for path in a.libs:
    capi._lib = None
    capi.LIB_PATH = os.path.abspath(path)
    ctx = capi.Context(0)
    cells = sum(capi.ksw_cells(len(q), len(t), w)[0] for q, t, w in zip(qs, ts, ws))
    out = []
    for rep in range(a.reps):
        try:
            res, _ = ctx.ksw_batch(qs, ts, ws, want_cigars=False)
            ok = sum(r.ok for r in res)
        except RuntimeError as e:
            print('   ', e, flush=True); ok = -1; out.append('failed'); continue
        tm = ctx.timing()
        out.append(f"{cells / tm.dp_ms / 1e6:.1f}")
    print(f"{os.path.basename(path):24s} n={a.n} len={a.len} err={a.err} band={ws[0]} ok={ok} GCUPS " + " ".join(out), flush=True)
    ctx.close()
