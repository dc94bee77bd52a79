import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from lesv_b200.synth import mutate
ctx = capi.Context(0)
rng = np.random.default_rng(11)
pairs = []
for i in range(12):
    n = int(rng.integers(11000, 30000))
    t = rng.integers(0, 4, n, dtype=np.uint8)
    q = np.ascontiguousarray(mutate(t, 0.12, rng))
    if i % 3 == 0:
        k = n // 2
        q = np.ascontiguousarray(np.concatenate([q[:k], rng.integers(0, 4, 2000, dtype=np.uint8), q[k:]]))
    pairs.append((q, t))
keep = [int(x) for x in sys.argv[1].split(",")]
for rep in range(3):
    qs = [pairs[i][0] for i in keep]; ts = [pairs[i][1] for i in keep]
    res, cigs = ctx.ksw_batch(qs, ts, [-1] * len(keep))
    print(rep, [(i, r.aln_size, r.n_match) for i, r in zip(keep, res)])
