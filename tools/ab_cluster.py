"""Timing of the 2-CTA cluster class on its own (bands of 10-20 k): python tools/ab_cluster.py [liblesvgpu.so]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
if len(sys.argv) > 1:
    capi.LIB_PATH = os.path.abspath(sys.argv[1])
from tools.quick_bench import run
ctx = capi.Context(0)
rng = np.random.default_rng(1)
for err in (0.17, 0.20, 0.25, 0.30):           # band = 1.2 x err x 50 kb: 10.2 k, 12 k, 15 k, 18 k
    run(ctx, 296, 50000, err, rng, reps=1)
    print("   ", [(c["ctas_per_alignment"], c["threads"], c["n_ctas"], round(c["cells"] / max(c["done_ms"], 1e-9) / 1e6)) for c in ctx.class_timing()], flush=True)
