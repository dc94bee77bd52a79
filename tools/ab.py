"""A/B timing of K-ksw builds (development aid): python tools/ab.py <liblesvgpu.so> [label]
Loads the given build of the library instead of lesv_b200/liblesvgpu.so and times fixed batches."""
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
run(ctx, 592, 10000, 0.12, rng, reps=2)
run(ctx, 1184, 50000, 0.12, rng, reps=2)      # 8 alignments per CTA, tier 512
run(ctx, 592, 50000, 0.15, rng, reps=1)       # tier 640
run(ctx, 592, 30000, 0.12, rng, reps=1)       # tier 320
