"""One K-ksw batch (148 x 50 kb at 12 %), for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from tools.quick_bench import run
ctx = capi.Context(0)
rng = np.random.default_rng(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
L = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
run(ctx, n, L, 0.12, rng, reps=1)
