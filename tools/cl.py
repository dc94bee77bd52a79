"""Cluster-path timing: wide bands (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from tools.quick_bench import run
ctx = capi.Context(0)
rng = np.random.default_rng(1)
run(ctx, 74, 30000, 0.12, rng, reps=2, full=True)      # 1878 blocks -> clusters of 4
run(ctx, 74, 50000, 0.22, rng, reps=2)                 # band 13200 -> 828 blocks -> clusters of 2
