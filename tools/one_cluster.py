"""One K-ksw batch of the 2-CTA cluster class (24 x 20 kb pairs, band 10.8 k), for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from tools.quick_bench import run
ctx = capi.Context(0)
rng = np.random.default_rng(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
L = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
run(ctx, n, L, 0.45, rng, reps=1)
print([(c["ctas_per_alignment"], c["threads"], c["n_ctas"], round(c["cells"] / max(c["done_ms"], 1e-9) / 1e6)) for c in ctx.class_timing()], flush=True)
