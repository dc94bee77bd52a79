"""BASELINE.json configs[4]: banded-DP microbenchmark -- query/target pairs of 1-50 kb across band widths 64-1024,
GCUPS of lesvgpu_ksw_batch (kernel time, CUDA events) next to the reference's nw_ksw2_extd2 on all host cores.
usage: ksw_sweep.py [out.txt]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import capi
from lesv_b200.synth import mutate
import bench

ctx = capi.Context(0)
rng = np.random.default_rng(5)
cores = os.cpu_count() or 1
lines = [f"# banded dual-affine DP sweep (B200, kernel time; reference = nw_ksw2_extd2 on {cores} host threads, 2 x cores pairs)",
         f"{'len':>6s} {'band':>5s} {'pairs':>7s} {'Gcells':>8s} {'gpu ms':>9s} {'GPU GCUPS':>10s} {'CPU GCUPS':>10s} {'ok':>6s}"]
try:
    kind, lib = bench_cpu = None, None
    from bench import REF_BIN
    import ctypes as C
    so = os.path.join(bench.ROOT, "oracle", "_ref", "libhbn_ref.so")
    lib = C.CDLL(so) if os.path.exists(so) else None
    if lib is not None:
        lib.Ksw2DataNew.restype = C.c_void_p
        lib.ksw2_extd2_set_params.argtypes = [C.c_void_p]
        lib.nw_ksw2_extd2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int] + [C.c_void_p] * 7
except Exception:
    lib = None


class KString(C.Structure):
    _fields_ = [("l", C.c_size_t), ("m", C.c_size_t), ("s", C.c_void_p)]


def cpu_gcups(pairs, w):
    import threading
    if lib is None:
        return float("nan")
    it = iter(range(len(pairs)))
    lock = threading.Lock()

    def worker():
        data = lib.Ksw2DataNew()
        lib.ksw2_extd2_set_params(data)
        qa, sa = KString(), KString()
        ints = [C.c_int() for _ in range(4)]
        ident = C.c_double()
        while True:
            with lock:
                i = next(it, None)
            if i is None:
                return
            q, t = pairs[i]
            lib.nw_ksw2_extd2(data, 0, q.ctypes.data, 0, len(q), len(q), 0, t.ctypes.data, 0, len(t), len(t), 0,
                              C.c_double(0.0), w, *[C.addressof(x) for x in ints], C.addressof(ident), C.addressof(qa), C.addressof(sa))
    ths = [threading.Thread(target=worker) for _ in range(cores)]
    t0 = time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    sec = time.perf_counter() - t0
    cells = sum(capi.ksw_cells(len(q), len(t), w)[0] for q, t in pairs)
    return cells / sec / 1e9


LENS = tuple(int(x) for x in os.environ.get('SWEEP_LENS', '1000,5000,10000,20000,50000').split(','))
for L in LENS:
    base = []
    for _ in range(64):
        t = rng.integers(0, 4, L, dtype=np.uint8)
        # substitutions only near the ends keep |qlen - tlen| small enough for the narrow bands; 6 % indels elsewhere
        q = np.ascontiguousarray(mutate(t, 0.06, rng))
        base.append((q, t))
    for w in (64, 128, 256, 512, 1024):
        good = [(q, t) for q, t in base if abs(len(q) - len(t)) <= w]
        if not good:
            continue
        per = capi.ksw_cells(len(good[0][0]), len(good[0][1]), w)[0]
        n = int(min(200000, max(len(good), 40e9 // per)))
        pairs = [good[i % len(good)] for i in range(n)]
        qs, ts = [p[0] for p in pairs], [p[1] for p in pairs]
        cells = sum(capi.ksw_cells(len(q), len(t), w)[0] for q, t in good) * (n // len(good)) + \
            sum(capi.ksw_cells(len(q), len(t), w)[0] for q, t in good[:n % len(good)])
        ctx.ksw_batch(qs[:len(good)], ts[:len(good)], [w] * len(good), want_cigars=False)      # warm-up
        res, _ = ctx.ksw_batch(qs, ts, [w] * n, want_cigars=False)
        tm = ctx.timing()
        cpu = cpu_gcups(good[:2 * cores] if len(good) >= 2 * cores else good, w)
        lines.append(f"{L:6d} {w:5d} {n:7d} {cells / 1e9:8.2f} {tm.dp_ms:9.2f} {cells / tm.dp_ms / 1e6:10.1f} {cpu:10.2f} {sum(r.ok for r in res):6d}")
        print(lines[-1], flush=True)
out = sys.argv[1] if len(sys.argv) > 1 else None
if out:
    open(out, "w").write("\n".join(lines) + "\n")
ctx.close()
