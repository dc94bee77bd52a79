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
elif scen == "bigsv":                  # 50 kb reads across 6-10 kb INS/DEL: bands of 10-20 k (the 2-CTA cluster class) at real size
    n = int(os.environ.get("BIGSV_READS", "24"))
    ref = synth.random_codes(3_000_000, rng)
    sample, events = synth.plant_events(ref, rng, 12, sizes=(6000, 7000, 8000, 9000, 10000), min_gap=100_000)
    shift, spos = 0, []                # event positions in sample coordinates
    for kind, pos, size in events:
        spos.append(pos + shift)
        shift += size if kind == "INS" else -size
    r2 = np.random.default_rng(11)
    reads = []
    for i in range(n):
        c = spos[i % len(spos)] + int(r2.integers(-15000, 15000))
        s0 = max(0, min(len(sample) - 50000, c - 25000))
        rd = synth.mutate(sample[s0:s0 + 50000], float(r2.uniform(0.10, 0.15)), r2)
        reads.append(synth.revcomp_codes(rd) if i % 2 else rd)
    synth.write_fasta(os.path.join(work, "ref.fa"), ["chrB"], [ref], width=0)
    synth.write_fasta(os.path.join(work, "reads.fa"), [f"r{i}" for i in range(n)], reads)
elif scen in ("trf", "minindel", "m4x", "md"):
    # trf      = a TRF bed file through qx2svr and qx2svsig (trf_array.cpp:10-83): masked stretches cover some SV reads
    #            entirely (dropped), some partly (kept), plus a contig name the database does not know (skipped)
    # minindel = qx2svsig with min_indel_size 1: hundreds of gap runs per read (no cap in find_sv_signature.c:126-220)
    # m4x      = qx2map -outfmt m4x (binary M4Record stream, mecat_results.c:281-297) on stdout and in the backup file
    # md       = qx2map -cigar -md (MD:Z: column, mecat_results.c:86-135)
    info = synth.make_dataset(work, ref_len=900_000, n_reads=90, read_len=9000, n_events=18, seed=61, n_contigs=3,
                              err_lo=0.10, err_hi=0.15, event_sizes=(40, 60, 100, 300, 1000, 3000))
    if scen == "trf":
        with open(os.path.join(work, "trf.bed"), "w") as f:
            for c in range(3):
                for a, b in ((20_000, 60_000), (100_000, 103_000), (150_000, 230_000)):
                    f.write(f"contig{c}\t{a}\t{b}\n")
            f.write("chrUnknown\t5\t500\n")
elif scen == "tails":                  # junk / very noisy read ends: edlib's end extension gives up, the DALIGNER wave runs
    ref = synth.random_codes(1_000_000, rng)
    reads = synth.reads_with_noisy_tails(ref, 80, np.random.default_rng(21))
    synth.write_fasta(os.path.join(work, "ref.fa"), ["chrT"], [ref], width=0)
    synth.write_fasta(os.path.join(work, "reads.fa"), [f"r{i}" for i in range(len(reads))], reads)
elif scen == "c4":                     # BASELINE.json configs[3] shape: ultra-long reads split by qx2splitseq 50000 0 20000,
    # 35 % of the reads with a 20-30 % error rate, -kmer_size 19 -kmer_window 20 (README.md:164), SVR identity 80.0
    ref_len = int(float(os.environ.get("C4_REF_MBP", "8")) * 1_000_000)
    contigs = [synth.random_codes(ref_len // 4, rng) for _ in range(4)]
    samples = [synth.plant_events(c, rng, 6, sizes=(40, 100, 300, 1000, 3000), min_gap=100_000)[0] for c in contigs]
    r2 = np.random.default_rng(31)
    reads = []
    for i in range(int(os.environ.get("C4_READS", "10"))):
        smp = samples[i % 4]
        L = int(min(len(smp) - 1000, max(150_000, np.exp(r2.normal(np.log(200_000), 0.6)))))
        s0 = int(r2.integers(0, len(smp) - L))
        err = float(r2.uniform(0.20, 0.30)) if r2.random() < 0.35 else float(r2.uniform(0.08, 0.15))
        rd = synth.mutate(smp[s0:s0 + L], err, r2)
        reads.append(synth.revcomp_codes(rd) if i % 2 else rd)
    synth.write_fasta(os.path.join(work, "ref.fa"), [f"chr{i + 1}" for i in range(4)], contigs, width=0)
    synth.write_fasta(os.path.join(work, "long.fa"), [f"ul{i}" for i in range(len(reads))], reads, width=0)
    map_args += ["-kmer_size", "19", "-kmer_window", "20"]
else:
    raise SystemExit("unknown scenario")
env = dict(os.environ, OMP_NUM_THREADS="1")
t = str(os.cpu_count() or 4)
svr_args = ["3000", "80.0" if scen == "c4" else "70.0", "300", t] + (["trf.bed"] if scen == "trf" else [])
min_indel = "1" if scen == "minindel" else "40"
svsig_tail = [min_indel, t] + (["trf.bed"] if scen == "trf" else [])


def run(cmd, out=None, tag=""):
    t0 = time.time()
    with open(os.path.join(work, "log.txt"), "ab") as log:
        subprocess.run(cmd, cwd=work, check=True, env=env, stdout=out or log, stderr=log)
    print(f"  {tag or os.path.basename(cmd[0])}: {time.time() - t0:.1f} s", flush=True)


if scen == "c4":                       # the step before the path: both qx2splitseq builds cut the long reads
    with open(os.path.join(work, "reads.fa"), "wb") as f:
        run([os.path.join(REF, "qx2splitseq"), "50000", "0", "20000", "long.fa"], f, "qx2splitseq (reference)")
    with open(os.path.join(work, "reads_gpu.fa"), "wb") as f:
        run([os.path.join(OURS, "qx2splitseq"), "50000", "0", "20000", "long.fa"], f, "qx2splitseq (ours)")
    same = open(os.path.join(work, "reads.fa"), "rb").read() == open(os.path.join(work, "reads_gpu.fa"), "rb").read()
    print(f"splitseq: identical {same}")
if scen in ("m4x", "md"):              # mapper output formats only: the rest of the chain reads the default m4 text
    fmt = ["-outfmt", "m4x"] if scen == "m4x" else ["-cigar", "-md"]
    with open(os.path.join(work, "ref.out"), "wb") as f:
        run([os.path.join(REF, "qx2map")] + map_args + fmt + ["-db_dir", "db", "-keep_db", "-num_threads", "1", "reads.fa", "ref.fa"], f, "qx2map (reference)")
    with open(os.path.join(work, "gpu.out"), "wb") as f:
        run([os.path.join(OURS, "qx2map")] + map_args + fmt + ["-db_dir", "db_gpu", "-keep_db", "reads.fa", "ref.fa"], f, "qx2map (ours)")
    if scen == "m4x":
        def recs(p):
            raw = np.fromfile(p, dtype=np.uint8)
            assert len(raw) % 56 == 0, p
            return sorted(bytes(r) for r in raw.reshape(-1, 56)[:, :52])     # 4 trailing pad bytes are uninitialised in the reference
        bk = "backup_results/stagebackup_results_Q00000000_D00000000"
        a, b = recs(os.path.join(work, "ref.out")), recs(os.path.join(work, "gpu.out"))
        a2, b2 = recs(os.path.join(work, "db", bk)), recs(os.path.join(work, "db_gpu", bk))
        print(f"m4x: reference {len(a)} records, ours {len(b)}, identical {a == b}; backup files identical {a2 == b2 and a2 == a}")
    else:
        a, b = sorted(open(os.path.join(work, "ref.out")).readlines()), sorted(open(os.path.join(work, "gpu.out")).readlines())
        n_md = sum("\tMD:Z:" in x for x in a)
        print(f"md: reference {len(a)} lines ({n_md} with MD:Z:), ours {len(b)}, identical {a == b}")
    sys.exit(0)
print("reference:")
with open(os.path.join(work, "ref.m4"), "wb") as f:
    run([os.path.join(REF, "qx2map")] + map_args + ["-db_dir", "db", "-keep_db", "-num_threads", os.environ.get("REF_MAP_THREADS", "1"),
         "reads.fa", "ref.fa"], f, "qx2map")
run([os.path.join(REF, "qx2m4x"), "db"])
run([os.path.join(REF, "qx2svr"), "svr", "db"] + svr_args)
run([os.path.join(REF, "qx2svsig"), "svsig", "svr", "db"] + svsig_tail)
if os.environ.get("PROBE_REF_ONLY"):   # scenario check without a GPU: how the reference behaves on this input
    for d, pat in (("svr", ".sv_read"), ("svsig", ".sv_signature")):
        print(d, sum(len(open(os.path.join(work, d, f)).readlines()) for f in os.listdir(os.path.join(work, d)) if f.endswith(pat)), "records")
    print("m4 lines", len(open(os.path.join(work, "ref.m4")).readlines()))
    sys.exit(0)
print("ours:")
env["LESV_TIMING"] = "1"
with open(os.path.join(work, "gpu.m4"), "wb") as f:
    run([os.path.join(OURS, "qx2map")] + map_args + ["-db_dir", "db_gpu", "-keep_db", "reads.fa", "ref.fa"], f)
for l in open(os.path.join(work, "log.txt"), errors="replace"):
    if "[lesv-b200] '" in l: print("     " + l.strip())
run([os.path.join(OURS, "qx2m4x"), "db_gpu"])
run([os.path.join(OURS, "qx2svr"), "svr_gpu", "db_gpu"] + svr_args)
run([os.path.join(OURS, "qx2svsig"), "svsig_gpu", "svr_gpu", "db_gpu"] + svsig_tail)


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
