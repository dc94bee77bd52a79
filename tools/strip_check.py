"""Full-matrix alignments: strip kernel vs cluster kernel CIGARs (development aid). Run twice: with and without LESV_KSW_NO_STRIPS."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from lesv_b200.synth import mutate
ctx = capi.Context(0)
rng = np.random.default_rng(11)
qs, ts, ws = [], [], []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    n = int(rng.integers(11000, 30000))
    t = rng.integers(0, 4, n, dtype=np.uint8)
    q = np.ascontiguousarray(mutate(t, 0.12, rng))
    if i % 3 == 0:
        k = n // 2
        q = np.ascontiguousarray(np.concatenate([q[:k], rng.integers(0, 4, 2000, dtype=np.uint8), q[k:]]))
    qs.append(q); ts.append(t); ws.append(-1)
res, cigs = ctx.ksw_batch(qs, ts, ws)
for i, (r, c) in enumerate(zip(res, cigs)):
    h = hashlib.md5(np.asarray(c, dtype=np.uint32).tobytes()).hexdigest()[:12]
    print(i, len(qs[i]), len(ts[i]), r.ok, r.aln_size, r.n_match, len(c), h)
