"""BASELINE configs[4] sample (bench.py's c5 block) on its own (development aid): python tools/c5.py [pairs]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    os.environ["LESV_C5_PAIRS"] = sys.argv[1]
import bench
from lesv_b200 import capi
ctx = capi.Context(0)
out = bench.c5_block(ctx, capi, os.cpu_count() or 1, 2026, 148, 1965.0)
for r in out["bands"]:
    print(json.dumps(r))
ctx.close()
