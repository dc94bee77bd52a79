"""K-ksw rate against band width at 50 kb (development aid): err -> band -> threads per alignment -> CTAs per SM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from tools.quick_bench import run
ctx = capi.Context(0)
rng = np.random.default_rng(1)
for err in (0.05, 0.07, 0.085, 0.10, 0.12, 0.15):
    run(ctx, 296, 50000, err, rng, reps=1)
