"""A/B timing of the full-matrix (column strip) path: python tools/ab_strips.py <liblesvgpu.so> [label]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
if len(sys.argv) > 1:
    capi.LIB_PATH = os.path.abspath(sys.argv[1])
label = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(capi.LIB_PATH)
from tools.quick_bench import run
ctx = capi.Context(0)
print(f"== {label}", flush=True)
rng = np.random.default_rng(1)
run(ctx, 32, 50000, 0.12, rng, reps=2, full=True)      # qx2svr's dual-HSP chaining shape: ~32 full-matrix alignments per batch
run(ctx, 8, 50000, 0.12, rng, reps=1, full=True)
