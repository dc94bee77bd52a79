"""Summaries of ncu outputs kept under profiles/ (development aid).
  python tools/ncu_summaries.py launches <launch-list.csv> "<command line>"      -> per-kernel totals of an `--metrics gpu__time_duration.sum` pass
  python tools/ncu_summaries.py opcodes <report.ncu-rep> <kernel substring> <cells> -> opcode mix of one kernel from the source page"""
import collections, csv, io, re, subprocess, sys


def launches(path, cmd):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
    per = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*", "", r[4])
        val = float(r[14].replace(",", ""))
        unit = r[13]
        ms = val / 1e6 if unit in ("ns", "nsecond") else val / 1e3 if unit in ("us", "usecond") else val * 1e3 if unit in ("s", "second") else val
        per[name][0] += 1; per[name][1] += ms
    total = sum(v[1] for v in per.values())
    print(f"# ncu launch list of `{cmd}`")
    print("# (--metrics gpu__time_duration.sum --clock-control none: kernels serialised and cold, so shares matter, not absolutes)")
    print(f"# {sum(v[0] for v in per.values())} launches, {total:.1f} ms in kernels")
    print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {v[0]:8d} {v[1]:10.2f} {100 * v[1] / total:6.1f}%")


def opcodes(rep, pat, cells):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    isrc, iex = hdr.index("Source"), hdr.index("Instructions Executed")
    per = collections.Counter()
    for r in rows[hi + 1:]:
        try:
            n = int(r[iex])
        except (ValueError, IndexError):
            continue
        t = r[isrc].split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        base = op.split(".")[0] + (".MOV" if ".MOV" in op else "")
        per[base] += n
    total = sum(per.values())
    print(f"# {pat}: warp instructions executed {total / 1e9:.2f} G for {cells / 1e9:.2f} G DP cells = {32 * total / cells:.2f} thread-instructions per cell")
    print(f"# {'opcode':20s} {'warp-instr (G)':>14s} {'share':>7s}")
    for k, v in per.most_common(30):
        print(f"{k:22s} {v / 1e9:14.2f} {100 * v / total:6.1f}%")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        opcodes(sys.argv[2], sys.argv[3], float(sys.argv[4]))
