"""Timing of the warp-cooperative edlib_nw test hook on single pairs (development aid)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from lesv_b200.synth import mutate
ctx = capi.Context(0)
lib = ctx.lib
lib.lesvgpu_debug_edlib_nw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
rng = np.random.default_rng(1)
for n, m in [(300, 0), (1000, 0), (1800, 0), (2500, 0), (5000, 0), (50, 10000), (10000, 50), (1500, 100)]:
    q = rng.integers(0, 4, n, dtype=np.uint8)
    t = np.ascontiguousarray(mutate(q, 0.15, rng)) if m == 0 else rng.integers(0, 4, m, dtype=np.uint8)
    out = np.zeros(len(q) + len(t) + 8, dtype=np.uint8)
    ng = C.c_int()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        lib.lesvgpu_debug_edlib_nw(ctx.h, q.ctypes.data, len(q), t.ctypes.data, len(t), out.ctypes.data, C.byref(ng))
        best = min(best, time.perf_counter() - t0)
    print(f"n={len(q)} m={len(t)} ops={ng.value} wall {best*1e3:.2f} ms", flush=True)
