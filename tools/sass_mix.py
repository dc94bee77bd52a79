"""Static opcode mix of one kernel's SASS (development aid): python tools/sass_mix.py <obj|so> <kernel substring>
Counts instructions between the first and the last packed DP instruction (the block update is inlined once)."""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, body = None, collections.defaultdict(list)
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        body[cur].append(m.group(2))
for fn, ops in body.items():
    if pat not in fn:
        continue
    idx = [i for i, o in enumerate(ops) if o.startswith("VIMNMX3")]
    if not idx:
        continue
    seg = ops[idx[0] - 40:idx[-1] + 120]
    c = collections.Counter(o.split(".")[0] + (".MOV" if ".MOV" in o else "") for o in seg)
    print(fn[:60], "total", len(ops), "| DP segment", len(seg))
    print("   " + ", ".join(f"{k} {v}" for k, v in c.most_common(16)))
