import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from tools.quick_bench import run
ctx = capi.Context(0)
rng = np.random.default_rng(1)
run(ctx, 592, 10000, 0.12, rng, reps=2)
run(ctx, 148, 50000, 0.12, rng, reps=2)
