"""Exploratory parity runs of the four drop-in tools against oracle/_ref/bin on bigger / nastier inputs than the
test-suite uses (development aid).  usage: parity_probe.py <scenario> [workdir]"""
import os, sys, shutil, subprocess, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesv_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF, OURS = os.path.join(ROOT, "oracle", "_ref", "bin"), os.path.join(ROOT, "lesv_b200", "bin")
scen = sys.argv[1]
work = sys.argv[2] if len(sys.argv) > 2 else tempfile.mkdtemp(prefix="probe_")
os.makedirs(work, exist_ok=True)
map_args = ["-max_target_seqs", "5", "-max_hsps", "10"]
rng = np.random.default_rng(123)
if scen == "c1":                       # BASELINE.json configs[0]
    synth.make_dataset(work, ref_len=1_000_000, n_reads=2000, read_len=10000, n_events=20, seed=1, err_lo=0.12, err_hi=0.12,
                       event_sizes=(40, 60, 100, 300, 1000, 3000))
elif scen == "k19":                    # README's ultra-long setting
    synth.make_dataset(work, ref_len=2_000_000, n_reads=120, read_len=30000, n_events=10, seed=2, err_lo=0.10, err_hi=0.25)
    map_args += ["-kmer_size", "19", "-kmer_window", "20"]
elif scen == "repeats":                # tandem repeats + dispersed duplicates: ties in the unstable sorts
    ref = synth.random_codes(600_000, rng)
    unit = synth.random_codes(37, rng)
    for pos in (50_000, 210_000, 400_000):
        n = int(rng.integers(40, 120))
        ref[pos:pos + 37 * n] = np.tile(unit, n)
    seg = ref[100_000:106_000].copy()
    for pos in (300_000, 500_000):
        ref[pos:pos + 6000] = synth.mutate(seg, 0.01, rng)[:6000]
    reads = synth.simulate_reads(ref, 300, 9000, np.random.default_rng(5), 0.10, 0.15)
    synth.write_fasta(os.path.join(work, "ref.fa"), ["chrR"], [ref], width=80)
    synth.write_fasta(os.path.join(work, "reads.fa"), [f"r{i}" for i in range(len(reads))], reads)
elif scen == "c3":                     # BASELINE.json configs[2] shape: 24 contigs, 3.1 Gbp, 50 kb reads (a sample of them)
    hg19 = [249, 243, 198, 191, 181, 171, 159, 146, 141, 136, 135, 134, 115, 107, 103, 90, 81, 78, 59, 63, 48, 51, 155, 59]
    scale = float(os.environ.get("C3_SCALE", "1.0"))
    names, contigs = [], []
    for i, mb in enumerate(hg19):
        names.append(f"chr{i + 1}")
        contigs.append(synth.random_codes(int(mb * 1_000_000 * scale), rng))
    synth.write_fasta(os.path.join(work, "ref.fa"), names, contigs, width=0)
    reads = []
    r2 = np.random.default_rng(9)
    for i in range(int(os.environ.get("C3_READS", "64"))):
        c = contigs[int(r2.integers(0, len(contigs)))]
        p0 = int(r2.integers(0, len(c) - 50000))
        seg = c[p0:p0 + 50000]
        if i % 4 == 0:                 # a 40 bp - 3 kb deletion or insertion inside
            size = int(np.exp(r2.uniform(np.log(40), np.log(3000))))
            k = int(r2.integers(5000, 40000))
            seg = np.concatenate([seg[:k], seg[k + size:]]) if i % 8 == 0 else np.concatenate([seg[:k], synth.random_codes(size, r2), seg[k:]])
        rd = synth.mutate(seg, float(r2.uniform(0.10, 0.15)), r2)
        reads.append(synth.revcomp_codes(rd) if i % 2 else rd)
    del contigs
    synth.write_fasta(os.path.join(work, "reads.fa"), [f"r{i}" for i in range(len(reads))], reads)
else:
    raise SystemExit("unknown scenario")
env = dict(os.environ, OMP_NUM_THREADS="1")
t = str(os.cpu_count() or 4)


def run(cmd, out=None, tag=""):
    t0 = time.time()
    with open(os.path.join(work, "log.txt"), "ab") as log:
        subprocess.run(cmd, cwd=work, check=True, env=env, stdout=out or log, stderr=log)
    print(f"  {tag or os.path.basename(cmd[0])}: {time.time() - t0:.1f} s", flush=True)


print("reference:")
with open(os.path.join(work, "ref.m4"), "wb") as f:
    run([os.path.join(REF, "qx2map")] + map_args + ["-db_dir", "db", "-keep_db", "-num_threads", os.environ.get("REF_MAP_THREADS", "1"),
         "reads.fa", "ref.fa"], f, "qx2map")
run([os.path.join(REF, "qx2m4x"), "db"])
run([os.path.join(REF, "qx2svr"), "svr", "db", "3000", "70.0", "300", t])
run([os.path.join(REF, "qx2svsig"), "svsig", "svr", "db", "40", t])
print("ours:")
env["LESV_TIMING"] = "1"
with open(os.path.join(work, "gpu.m4"), "wb") as f:
    run([os.path.join(OURS, "qx2map")] + map_args + ["-db_dir", "db_gpu", "-keep_db", "reads.fa", "ref.fa"], f)
for l in open(os.path.join(work, "log.txt"), errors="replace"):
    if "[lesv-b200] '" in l: print("     " + l.strip())
run([os.path.join(OURS, "qx2m4x"), "db_gpu"])
run([os.path.join(OURS, "qx2svr"), "svr_gpu", "db_gpu", "3000", "70.0", "300", t])
run([os.path.join(OURS, "qx2svsig"), "svsig_gpu", "svr_gpu", "db_gpu", "40", t])


def lines(p):
    return sorted(open(p).readlines())


a, b = lines(os.path.join(work, "ref.m4")), lines(os.path.join(work, "gpu.m4"))
sa, sb = set(a), set(b)
print(f"m4: reference {len(a)} lines, ours {len(b)}, identical {a == b}; only-ref {len(sa - sb)}, only-ours {len(sb - sa)}")
for x in sorted(sa - sb)[:3]: print("   ref :", x.strip()[:160])
for x in sorted(sb - sa)[:3]: print("   ours:", x.strip()[:160])
for d1, d2, pat in (("svr", "svr_gpu", ".sv_read"), ("svsig", "svsig_gpu", ".sv_signature")):
    n = same = 0
    for fn in sorted(os.listdir(os.path.join(work, d1))):
        if fn.endswith(pat):
            x, y = lines(os.path.join(work, d1, fn)), lines(os.path.join(work, d2, fn))
            n += len(x); same += x == y
            if x != y:
                print(f"   {fn}: differs ({len(set(x) - set(y))} only-ref, {len(set(y) - set(x))} only-ours)")
    print(f"{d1}: {n} records, all files identical: {same == len([f for f in os.listdir(os.path.join(work, d1)) if f.endswith(pat)])}")
