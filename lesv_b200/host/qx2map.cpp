// qx2map (GPU) -- the reference mapper's command line, database directory and output formats
// (src/app/map/main.c:10-76, cmdline_args.cpp:15-68, 633-713), with the per-batch work done on the GPU
// through lesvgpu_map_batch (include/lesvgpu.h):
//     qx2map [options] <reads.fasta> <reference.fasta>  > map.m4
// Options kept: -query_batch_size -max_query_vol_res -max_subject_vol_res -min_query_size -min_subject_size
//   -db_dir -keep_db -kmer_size -kmer_window -max_kmer_occ -min_ddfs -memsc_kmer_size -memsc_kmer_window
//   -memsc_mem_score -qcov_hsp_res -perc_identity -skip_overhang -max_hsps -max_target_seqs -outfmt {m4,m4x,subseq}
//   -cigar -md -num_threads (accepted; one feeder thread per GPU) -grid "node_id num_nodes" -out -task rm
// Files: <db_dir>/{query,subject}.* (lesv_dbbuild.hpp), <db_dir>/backup_results/stagebackup_results_Q%08d_D%08d
// [+ .mapped], exactly as hbn_job_control.c:4-92 names them, so qx2m4x / qx2svr run on the result unchanged.
// Output order is deterministic (volume, batch, read), where the reference's depends on thread scheduling.
// Several GPUs (lesv_gpus.hpp; LESV_GPUS / LESV_GPU select devices): the index is built on every GPU, the query batches
// of a volume -- fixed by -query_batch_size alone, so the query-side occurrence filter sees the same reads whatever the
// number of GPUs -- are pulled from one queue and written in batch order.  -grid "node_id
// num_nodes" still shards whole query volumes over processes / nodes.
#include "../../include/lesvgpu.h"
#include "lesv_dbbuild.hpp"
#include "lesv_gpus.hpp"

#include <mutex>
#include <string>
#include <time.h>

using namespace lesvdb;

namespace {

struct Options {
    size_t query_batch_size = 500000000, max_query_vol_res = 4000000000ull, max_subject_vol_res = 4000000000ull;
    int min_query_size = 0, min_subject_size = 0, max_query_vol_seqs = INT_MAX, max_subject_vol_seqs = INT_MAX;
    std::string db_dir = "hbndb", outfmt = "m4", output = "-", query, subject;
    bool keep_db = false, cigar = false, md = false, skip_overhang = false;
    int kmer_size = 15, kmer_window = 10, max_kmer_occ = 200, min_ddfs = 3;
    int memsc_kmer_size = 10, memsc_kmer_window = 10, memsc_score = 30;
    int qcov_hsp_res = 100, max_hsps = 5, hitlist = 3, num_threads = 1, node_id = 0, num_nodes = 1;
    double perc_identity = 0.0;
};

size_t parse_size(const std::string& s)              // NStr::StringToUInt8_DataSize: k/m/g decimal, ki/mi/gi binary
{
    size_t pos = 0;
    const double v = std::stod(s, &pos);
    std::string suf = s.substr(pos);
    for (auto& c : suf) c = (char)tolower((unsigned char)c);
    while (!suf.empty() && (suf.back() == 'b' || isspace((unsigned char)suf.back()))) suf.pop_back();
    double m = 1;
    if (suf == "k") m = 1e3; else if (suf == "m") m = 1e6; else if (suf == "g") m = 1e9; else if (suf == "t") m = 1e12;
    else if (suf == "ki") m = 1024.; else if (suf == "mi") m = 1048576.; else if (suf == "gi") m = 1073741824.;
    else if (!suf.empty()) die("bad data size: " + s);
    return (size_t)(v * m);
}

void usage()
{
    fprintf(stderr, "USAGE: qx2map [options] query subject\n"
                    "  (options of the reference's qx2map: -db_dir -keep_db -max_target_seqs -max_hsps -query_batch_size ...)\n");
    exit(1);
}

Options parse(int argc, char** argv)
{
    Options o;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) die("option " + a + " needs a value"); return argv[++i]; };
        if (a.size() < 2 || a[0] != '-' || isdigit((unsigned char)a[1])) { pos.push_back(a); continue; }
        const std::string k = a.substr(1);
        if (k == "query_batch_size") o.query_batch_size = parse_size(val());
        else if (k == "max_query_vol_res") o.max_query_vol_res = parse_size(val());
        else if (k == "max_subject_vol_res") o.max_subject_vol_res = parse_size(val());
        else if (k == "max_query_vol_seqs") o.max_query_vol_seqs = atoi(val().c_str());
        else if (k == "max_subject_vol_seqs") o.max_subject_vol_seqs = atoi(val().c_str());
        else if (k == "min_query_size") o.min_query_size = atoi(val().c_str());
        else if (k == "min_subject_size") o.min_subject_size = atoi(val().c_str());
        else if (k == "db_dir") o.db_dir = val();
        else if (k == "keep_db") o.keep_db = true;
        else if (k == "kmer_size") o.kmer_size = atoi(val().c_str());
        else if (k == "kmer_window") o.kmer_window = atoi(val().c_str());
        else if (k == "max_kmer_occ") o.max_kmer_occ = atoi(val().c_str());
        else if (k == "min_ddfs") o.min_ddfs = atoi(val().c_str());
        else if (k == "memsc_kmer_size") o.memsc_kmer_size = atoi(val().c_str());
        else if (k == "memsc_kmer_window") o.memsc_kmer_window = atoi(val().c_str());
        else if (k == "memsc_mem_score") o.memsc_score = atoi(val().c_str());
        else if (k == "memsc_mem_size" || k == "block_size" || k == "evalue" || k == "penalty" || k == "reward" ||
                 k == "gapopen" || k == "gapextend" || k == "strand") (void)val();       // parsed and unused by the reference too
        else if (k == "qcov_hsp_res") o.qcov_hsp_res = atoi(val().c_str());
        else if (k == "perc_identity") o.perc_identity = atof(val().c_str());
        else if (k == "skip_overhang") o.skip_overhang = true;
        else if (k == "max_hsps") o.max_hsps = atoi(val().c_str());
        else if (k == "max_target_seqs") o.hitlist = atoi(val().c_str());
        else if (k == "outfmt") o.outfmt = val();
        else if (k == "cigar") o.cigar = true;
        else if (k == "md") o.md = true;
        else if (k == "num_threads") o.num_threads = atoi(val().c_str());
        else if (k == "out") o.output = val();
        else if (k == "task") { if (val() != "rm") die("only -task rm is supported"); }
        else if (k == "grid") {
            const std::string g = val();
            if (sscanf(g.c_str(), "%d %d", &o.node_id, &o.num_nodes) != 2) die("bad -grid: " + g);
            if (o.node_id < 0 || o.num_nodes <= 0 || o.node_id >= o.num_nodes) die("bad -grid: " + g);
        } else if (k == "h" || k == "help") usage();
        else die("unknown option " + a);
    }
    if (pos.size() != 2) usage();
    o.query = pos[0]; o.subject = pos[1];
    if (o.query == o.subject) die("self-mapping (query == subject) is not supported by the GPU driver");
    if (o.outfmt != "m4" && o.outfmt != "m4x" && o.outfmt != "subseq") die("-outfmt " + o.outfmt + " is not supported by the GPU driver");
    return o;
}

std::string results_path(const Options& o, int qi, int sj)      // make_qi_vs_sj_results_path, hbn_job_control.c:4-19
{
    return o.db_dir + "/backup_results/stagebackup_results_Q" + fixed8(qi) + "_D" + fixed8(sj);
}
bool exists(const std::string& p) { return access(p.c_str(), F_OK) == 0; }
void copy_to(const std::string& path, FILE* out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) die("cannot open " + path);
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) fwrite(buf, 1, n, out);
    fclose(f);
}

struct M4Record {          // corelib/m4_record.h:11-24
    int qid, qdir, qoff, qend, qsize, sid, sdir, soff, send, ssize;
    double ident_perc;
    int score;
};

inline char base_char(const std::vector<uint8_t>& pac, size_t l) { return "ACGT"[(pac[l >> 2] >> ((~l & 3) << 1)) & 3]; }

// print_one_m4_result (mecat_results.c:256-297) for one HSP
void format_hsp(const Options& o, const lesvgpu_hsp& h, const uint8_t* ops, const Volume& qv, const Volume& sv, std::string& out)
{
    int qoff = h.qoff, qend = h.qend;
    if (h.qdir == 1) { qoff = h.qsize - h.qend; qend = h.qsize - h.qoff; }          // dump_m4_hits :325-330
    const double ident = 100.0 * h.n_match / h.aln_len;
    const int score = 8 * h.aln_len - h.n_match;
    const int qid_l = h.qid, sid_l = h.sid;
    if (o.outfmt == "m4x") {
        M4Record r;
        memset(&r, 0, sizeof r);
        r.qid = qv.info.seq_start_id + qid_l; r.qdir = h.qdir; r.qoff = qoff; r.qend = qend; r.qsize = h.qsize;
        r.sid = sv.info.seq_start_id + sid_l; r.sdir = 0; r.soff = h.soff; r.send = h.send; r.ssize = (int)h.ssize;
        r.ident_perc = ident; r.score = score;
        out.append((const char*)&r, sizeof r);
        return;
    }
    char buf[512];
    out += qv.name(qid_l);                              // names are not bounded: appended, not formatted into the fixed buffer
    out.push_back('\t');
    out += sv.name(sid_l);
    snprintf(buf, sizeof buf, "\t%g\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%lld", ident, score, h.qdir,
             qoff, qend, h.qsize, 0, h.soff, h.send, (long long)h.ssize);
    out += buf;
    if (o.cigar) {                                      // print_paf_cigar / print_cigar_core :27-62 (roles swapped there)
        out += "\tcg:Z:";
        for (int a = 0; a < h.aln_len;) {
            int b = a;
            while (b < h.aln_len && ops[b] == ops[a]) ++b;
            snprintf(buf, sizeof buf, "%d%c", b - a, "MDI"[ops[a]]);
            out += buf;
            a = b;
        }
    }
    if (o.md) {                                         // print_md :84-123: q := reference string, s := read string
        out += "\tMD:Z:";
        const size_t qbase = qv.seqs[qid_l].seq_offset, sbase = sv.seqs[sid_l].seq_offset;
        auto qchar = [&](int i) {                       // base i of the read on strand qdir
            if (!h.qdir) return base_char(qv.pac, qbase + (size_t)i);
            const char c = base_char(qv.pac, qbase + (size_t)(h.qsize - 1 - i));
            return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
        };
        int qi = h.qoff, si = h.soff, l_md = 0;
        for (int a = 0; a < h.aln_len;) {
            int b = a;
            while (b < h.aln_len && ops[b] == ops[a]) ++b;
            if (ops[a] == 2) si += b - a;               // gap in the read string: skipped
            else if (ops[a] == 1) {                     // gap in the reference string: "%d^" + read bases
                snprintf(buf, sizeof buf, "%d^", l_md);
                out += buf;
                for (int x = a; x < b; ++x) out.push_back(qchar(qi++));
                l_md = 0;
            } else {
                for (int x = a; x < b; ++x, ++qi, ++si) {
                    const char qc = qchar(qi), sc = base_char(sv.pac, sbase + (size_t)si);
                    if (qc != sc) { snprintf(buf, sizeof buf, "%d%c", l_md, qc); out += buf; l_md = 0; }
                    else ++l_md;
                }
            }
            a = b;
        }
        if (l_md > 0) { snprintf(buf, sizeof buf, "%d", l_md); out += buf; }
    }
    out.push_back('\n');
}

}  // namespace

static double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
// the reference logs "'<title>' takes N secs." around its phases (hbn_aux.h:95-106); LESV_TIMING=1 prints the same here
static void phase(const char* title, double& t0)
{
    static const bool on = getenv("LESV_TIMING") != nullptr;
    const double t = now_s();
    if (on) fprintf(stderr, "[lesv-b200] '%s' takes %.2f secs.\n", title, t - t0);
    t0 = t;
}

int main(int argc, char** argv)
{
    const Options o = parse(argc, argv);
    double t0 = now_s();
    if (!exists(o.db_dir) && mkdir(o.db_dir.c_str(), S_IRWXU) != 0) die("Failed to create directory " + o.db_dir);
    if (!db_is_built(o.query, o.db_dir, "query", o.min_query_size, o.max_query_vol_seqs, o.max_query_vol_res)) {
        build_db(o.query, o.db_dir, "query", o.min_query_size, o.max_query_vol_seqs, o.max_query_vol_res);
        db_make_built(o.query, o.db_dir, "query", o.min_query_size, o.max_query_vol_seqs, o.max_query_vol_res);
    }
    if (!db_is_built(o.subject, o.db_dir, "subject", o.min_subject_size, o.max_subject_vol_seqs, o.max_subject_vol_res)) {
        build_db(o.subject, o.db_dir, "subject", o.min_subject_size, o.max_subject_vol_seqs, o.max_subject_vol_res);
        db_make_built(o.subject, o.db_dir, "subject", o.min_subject_size, o.max_subject_vol_seqs, o.max_subject_vol_res);
    }
    phase("build databases", t0);
    const std::string bdir = o.db_dir + "/backup_results";
    if (!exists(bdir) && mkdir(bdir.c_str(), S_IRWXU) != 0) die("Failed to create directory " + bdir);
    FILE* out = o.output == "-" ? stdout : fopen(o.output.c_str(), "w");
    if (!out) die("cannot open " + o.output);

    lesvgpus::Pool gpus;
    gpus.open();
    phase("create GPU context", t0);
    lesvgpu_map_opts mo;
    memset(&mo, 0, sizeof mo);
    mo.min_ddfs = o.min_ddfs; mo.hitlist_size = o.hitlist; mo.max_hsps = o.max_hsps;
    mo.memsc_kmer_size = o.memsc_kmer_size; mo.memsc_kmer_window = o.memsc_kmer_window; mo.memsc_score = o.memsc_score;
    mo.qcov_hsp_res = o.qcov_hsp_res; mo.perc_identity = o.perc_identity; mo.skip_overhang = o.skip_overhang;
    mo.want_ops = (o.cigar || o.md) && o.outfmt == "m4";

    const int nqv = num_volumes(o.db_dir, "query"), nsv = num_volumes(o.db_dir, "subject");
    for (int svid = 0; svid < nsv; ++svid) {
        bool all_done = true;
        for (int q = o.node_id; q < nqv; q += o.num_nodes) all_done = all_done && exists(results_path(o, q, svid) + ".mapped");
        if (all_done) {
            for (int q = o.node_id; q < nqv; q += o.num_nodes) copy_to(results_path(o, q, svid), out);
            continue;
        }
        const Volume sv = load_volume(o.db_dir, "subject", svid);
        std::vector<lesvgpu_seq> sseqs(sv.seqs.size());
        for (size_t i = 0; i < sv.seqs.size(); ++i) { sseqs[i].seq_offset = sv.seqs[i].seq_offset; sseqs[i].seq_size = sv.seqs[i].seq_size; }
        gpus.each([&](int g) {
            if (lesvgpu_subject_load(gpus.ctx[g], sv.pac.data(), sv.pac.size()) != LESVGPU_OK ||
                lesvgpu_index_build(gpus.ctx[g], sseqs.data(), (int)sseqs.size(), o.kmer_size, o.kmer_window, o.max_kmer_occ) != LESVGPU_OK)
                die(std::string("index build failed: ") + lesvgpu_last_error(gpus.ctx[g]));
        });
        phase("load subject volume + build index", t0);
        for (int qvid = o.node_id; qvid < nqv; qvid += o.num_nodes) {
            const std::string rpath = results_path(o, qvid, svid);
            if (exists(rpath + ".mapped")) { copy_to(rpath, out); continue; }
            const Volume qv = load_volume(o.db_dir, "query", qvid);
            std::vector<lesvgpu_seq> qseqs(qv.seqs.size());
            size_t left = 0;
            for (size_t i = 0; i < qv.seqs.size(); ++i) {
                qseqs[i].seq_offset = qv.seqs[i].seq_offset; qseqs[i].seq_size = qv.seqs[i].seq_size; left += qv.seqs[i].seq_size;
            }
            FILE* backup = fopen(rpath.c_str(), "w");
            if (!backup) die("cannot open " + rpath);
            const int nq = (int)qseqs.size();
            std::vector<std::pair<int, int>> batches;
            for (int to = 0; to < nq;) {                // get_next_batch_query_qidx_range, hbn_align_one_volume.c:56-84
                const int from = to;
                size_t curr = qseqs[from].seq_size;
                to = from + 1;
                while (curr < o.query_batch_size && to < nq) { curr += qseqs[to].seq_size; ++to; }
                const size_t min_batch = (size_t)(o.query_batch_size * 0.2);
                left -= curr;
                if (left <= min_batch) { to = nq; left = 0; }
                batches.emplace_back(from, to);
            }
            // results are written in batch order as soon as every earlier batch has been written (bounded memory with -cigar)
            std::vector<std::string> texts(batches.size());
            std::vector<char> ready(batches.size(), 0);
            std::mutex write_mutex;
            size_t next_write = 0;
            gpus.queue((int)batches.size(), [&](int g, int bi) {
                lesvgpu_ctx* ctx = gpus.ctx[g];
                const int from = batches[bi].first, to = batches[bi].second;
                std::string& text = texts[bi];
                if (o.outfmt == "subseq") {
                    const lesvgpu_subseq_hit* w = nullptr;
                    const int* first = nullptr;
                    if (lesvgpu_map_candidates(ctx, qv.pac.data(), qseqs.data(), from, to, &mo, &w, &first) != LESVGPU_OK)
                        die(std::string("lesvgpu_map_candidates: ") + lesvgpu_last_error(ctx));
                    char buf[256];
                    for (int x = 0; x < first[to - from]; ++x) {    // dump_subseq_hit, app/map/hbn_subseq_hit.h:23-33
                        snprintf(buf, sizeof buf, "qid = %d, qdir = %d, qoff = %d, sid = %d, soff = %d, sfrom = %zu, sto = %zu, score = %d\n",
                                 w[x].qid + qv.info.seq_start_id, w[x].qdir, w[x].qoff, w[x].sid + sv.info.seq_start_id, w[x].soff,
                                 w[x].sfrom, w[x].sto, w[x].score);
                        text += buf;
                    }
                } else {
                    const lesvgpu_hsp* hsps = nullptr;
                    const int* first = nullptr;
                    const uint8_t* pool = nullptr;
                    if (lesvgpu_map_batch(ctx, qv.pac.data(), qseqs.data(), from, to, &mo, &hsps, &first, &pool) != LESVGPU_OK)
                        die(std::string("lesvgpu_map_batch: ") + lesvgpu_last_error(ctx));
                    for (int x = 0; x < first[to - from]; ++x)
                        format_hsp(o, hsps[x], pool ? pool + hsps[x].ops_off : nullptr, qv, sv, text);
                }
                std::lock_guard<std::mutex> lock(write_mutex);
                ready[bi] = 1;
                while (next_write < texts.size() && ready[next_write]) {
                    std::string& t = texts[next_write++];
                    fwrite(t.data(), 1, t.size(), out);
                    fwrite(t.data(), 1, t.size(), backup);
                    std::string().swap(t);
                }
            });
            phase("map the query batches of one volume", t0);
            fclose(backup);
            FILE* m = fopen((rpath + ".mapped").c_str(), "w");   // qi_vs_sj_make_mapped
            if (m) fclose(m);
        }
    }
    gpus.close();
    if (out != stdout) fclose(out);
    if (!o.keep_db) {                                   // hbn_task_struct_free, hbn_task_struct.c:50-54
        const std::string cmd = "rm -r " + o.db_dir;
        if (system(cmd.c_str()) != 0) fprintf(stderr, "[lesv-b200] could not remove %s\n", o.db_dir.c_str());
    }
    return 0;
}
