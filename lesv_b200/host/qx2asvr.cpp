// qx2asvr (GPU) -- same command line, same input and output files as the reference's qx2asvr
// (src/app/necat2sv/map_cns_sv_read.c:396-447):
//     qx2asvr subject_pdb_dir svr_group_dir num_threads
// For every contig it packs <svr_group_dir>/subject_%08d.normal.cns.fasta (consensus reads of the SV read groups) into
// <svr_group_dir>/s%08d_svr_pdb, aligns every consensus read -- on the strand its header names -- globally against the
// reference window its header names with nw_ksw2_extd2 (band 0.2 x the longer sequence, then the full band; :205-258),
// cuts the alignment back to the consensus sub-range (s_dump_sv_read_info, :59-163), drops reads below 85 %
// effective identity and writes one SAM record each (print_one_sam_result, map_results.c:249-323) to
// subject_%08d.normal.cns.fasta.sam, then touches the .mapped sentinel.  The alignments run on the GPU through
// lesvgpu_ksw_batch (include/lesvgpu.h), sub-batches of LESV_ASVR_BATCH reads (default 256) pulled by one feeder thread
// per GPU (lesv_gpus.hpp); records are written in read order (the reference's order depends on thread scheduling).
// num_threads is accepted and ignored.
#include "../../include/lesvgpu.h"
#include "lesv_dbbuild.hpp"
#include "lesv_gpus.hpp"

#include <algorithm>
#include <string>

using namespace lesvdb;

namespace {

const char* kSamVersion = "1.6";                 // map_results.c:3
const char* kPackageVersion = "2.0.0";           // HBN_PACKAGE_VERSION, corelib/hbn_package_version.h:12-25
const char* kDefaultSample = "mmsv";             // DEFAULT_SAMPLE, map_results.h:20

// the ints behind the last "<tag>:" of a name (extract_info_from_sv_read_header, sv_read_group_file_name.cpp:132-196;
// extract_info_from_cns_read_header, corelib/cns_read_header.c:24-95): both scan BACKWARDS for the three tag letters
std::vector<int> header_ints(const std::string& name, const char* tag, int n_ints)
{
    const int n = (int)name.size();
    int x = n;
    bool found = false;
    while (x > 2) {
        --x;
        if (name[x] == tag[2] && name[x - 1] == tag[1] && name[x - 2] == tag[0]) { x -= 2; found = true; break; }
    }
    if (!found || x + 3 >= n || name[x + 3] != ':') die("read name without a " + std::string(tag) + ": field: " + name);
    std::vector<int> v;
    x += 4;
    for (int k = 0; k < n_ints; ++k) {
        if (x >= n) die("short " + std::string(tag) + ": field in " + name);
        v.push_back(atoi(name.c_str() + x));
        if (k + 1 < n_ints) {
            while (x < n && name[x] != ':') ++x;
            if (x >= n) die("short " + std::string(tag) + ": field in " + name);
            ++x;
        }
    }
    return v;
}

inline uint8_t pac_code(const std::vector<uint8_t>& pac, size_t l) { return (pac[l >> 2] >> ((~l & 3) << 1)) & 3; }

// calc_effective_ident_perc, algo/hbn_traceback_aux.c:22-95: gap runs of >= 20 columns do not count
double effective_ident(const std::string& q, const std::string& s)
{
    const int n = (int)q.size(), E = 20;
    int eff_len = 0, eff_mat = 0, i = 0;
    while (i < n) {
        if (q[i] != '-' && s[i] != '-') { eff_mat += q[i] == s[i]; ++eff_len; ++i; continue; }
        const std::string& g = q[i] == '-' ? q : s;           // the string the run of gaps is in
        int j = i + 1;
        while (j < n && g[j] == '-') ++j;
        if (j - i < E) eff_len += j - i;
        i = j;
    }
    return eff_len == 0 ? 0.0 : 100.0 * eff_mat / eff_len;
}

// print_cigar_core + print_md, map_results.c:63-157 (q = read string, s = reference string)
void sam_cigar_md(const std::string& q, const std::string& s, std::string& cigar, std::string& md)
{
    const int n = (int)q.size();
    char buf[32];
    int l_md = 0;
    for (int i = 0; i < n;) {
        int j = i;
        if (q[i] == '-') {
            while (j < n && q[j] == '-') ++j;
            snprintf(buf, sizeof buf, "%dD", j - i); cigar += buf;
            snprintf(buf, sizeof buf, "%d^", l_md); md += buf;
            md.append(s, i, j - i);
            l_md = 0;
        } else if (s[i] == '-') {
            while (j < n && s[j] == '-') ++j;
            snprintf(buf, sizeof buf, "%dI", j - i); cigar += buf;
        } else {
            while (j < n && q[j] != '-' && s[j] != '-') ++j;
            snprintf(buf, sizeof buf, "%dM", j - i); cigar += buf;
            for (int k = i; k < j; ++k) {
                if (q[k] != s[k]) { snprintf(buf, sizeof buf, "%d%c", l_md, s[k]); md += buf; l_md = 0; }
                else ++l_md;
            }
        }
        i = j;
    }
    if (l_md > 0) { snprintf(buf, sizeof buf, "%d", l_md); md += buf; }
}

struct CnsRead {
    int read_dir, sid, sfrom, sto, cns_from, cns_to, length;
    std::vector<uint8_t> read, window;         // residues 0..3: the read on its strand, subject[sfrom, sto)
};

// s_dump_sv_read_info (:59-163) from the CIGAR of the global alignment: one SAM line, or nothing
void sam_record(const CnsRead& R, const char* read_name, const char* subject_name, int subject_id, int subject_size,
                const uint32_t* cigar, int n_cigar, std::string& out)
{
    std::string qaln, saln;
    int qi = 0, si = 0;
    for (int k = 0; k < n_cigar; ++k) {                         // nw_ksw2_extd2, ksw2_wrapper.c:139-181
        const int len = (int)(cigar[k] >> 4), op = (int)(cigar[k] & 15);
        for (int t = 0; t < len; ++t) {
            qaln.push_back(op == 2 ? '-' : "ACGT"[R.read[qi++]]);
            saln.push_back(op == 1 ? '-' : "ACGT"[R.window[si++]]);
        }
    }
    const int aln_size = (int)qaln.size(), qb = 0, qe = R.length, sb = 0;
    int cns_qb = R.cns_from, cns_qe = R.cns_to;
    if (R.read_dir != 0) { cns_qb = R.length - R.cns_to; cns_qe = R.length - R.cns_from; }
    int aln_from = 0;
    qi = qb;
    while (qi < cns_qb) { if (qaln[aln_from] != '-') ++qi; ++aln_from; }
    int aln_to = aln_size;
    qi = qe;
    while (qi > cns_qe) { if (qaln[aln_to - 1] != '-') --qi; --aln_to; }
    if (aln_from >= aln_to) return;
    int qif = qb, sif = sb;
    for (int a = 0; a < aln_from; ++a) { qif += qaln[a] != '-'; sif += saln[a] != '-'; }
    int qie = qif, sie = sif;
    for (int a = aln_from; a < aln_to; ++a) { qie += qaln[a] != '-'; sie += saln[a] != '-'; }
    const std::string q = qaln.substr(aln_from, aln_to - aln_from), s = saln.substr(aln_from, aln_to - aln_from);
    const int as = (int)q.size();
    int n_ident = 0;
    for (int a = 0; a < as; ++a) n_ident += q[a] == s[a];
    const int score = 8 * as - n_ident;                          // calc_ident_perc: MATCH_REWARD 8, MISMATCH_PENALTY -1
    const double ident = 100.0 * n_ident / as, eff = effective_ident(q, s);
    if (eff < 85.0) return;
    (void)sie;
    std::string cigar_s, md;
    sam_cigar_md(q, s, cigar_s, md);
    char buf[256];
    out += read_name;
    snprintf(buf, sizeof buf, "\t%d\t", R.read_dir != 0 ? 0x10 : 0);
    out += buf;
    out += subject_name;                                        // names are not bounded: no %s into the fixed buffer
    snprintf(buf, sizeof buf, "\t%d\t%d\t", R.sfrom + sif + 1, 60);
    out += buf;
    out += cigar_s;                                             // qoff 0, qend == qsize: no soft clips (:151)
    out += "\t*\t*\t0\t";
    for (char c : q) if (c != '-') out.push_back(c);
    snprintf(buf, sizeof buf, "\t*\ts1:i:%d\tNM:i:%d\tAS:i:%d\tMD:Z:", score, as - n_ident, score);
    out += buf;
    out += md;
    snprintf(buf, sizeof buf, "\tRG:Z:subject%s\tqs:i:%d\tqe:i:%d\tql:i:%d\tmc:f:%g\tec:f:%g\n", fixed8(subject_id).c_str(), 0, qie - qif,
             qie - qif, ident, eff);
    out += buf;
    (void)subject_size;
}

}  // namespace

int main(int argc, char* argv[])
{
    if (argc != 4) {
        fprintf(stderr, "USAGE:\n%s subject_pdb_dir svr_group_dir num_threads\n", argv[0]);
        return 1;
    }
    const std::string subject_dir = argv[1], svr_dir = argv[2];
    const int batch = std::max(1, getenv("LESV_ASVR_BATCH") ? atoi(getenv("LESV_ASVR_BATCH")) : 256);
    Reader subj;
    subj.open(subject_dir, "subject");
    lesvgpus::Pool gpus;
    bool opened = false;

    for (int sid = 0; sid < subj.info.num_seqs; ++sid) {
        const std::string fasta = svr_dir + "/subject_" + fixed8(sid) + ".normal.cns.fasta";
        if (access((fasta + ".mapped").c_str(), F_OK) == 0) continue;                       // :372-379
        FILE* sam = fopen((fasta + ".sam").c_str(), "w");
        if (!sam) die("cannot write " + fasta + ".sam");
        // print_sam_prolog, map_results.c:13-61
        fprintf(sam, "@HD\tVN:%s\tSO:unknown\tGO:query\n", kSamVersion);
        for (int i = 0; i < subj.info.num_seqs; ++i) fprintf(sam, "@SQ\tSN:%s\tLN:%zu\n", subj.name(i), (size_t)subj.seqs[i].seq_size);
        fprintf(sam, "@PG\tID:%s\tVN:%s\tCL:", argv[0], kPackageVersion);
        for (int i = 0; i < argc - 1; ++i) fprintf(sam, "%s ", argv[i]);
        fprintf(sam, "%s\tPN:%s\n", argv[argc - 1], argv[0]);
        fprintf(sam, "@RG\tID:subject%s\tSM:%s\n", fixed8(sid).c_str(), kDefaultSample);

        // s_make_subject_raw_reads, :287-298
        const std::string pdb = svr_dir + "/s" + fixed8(sid) + "_svr_pdb";
        if (access(pdb.c_str(), F_OK) != 0 && mkdir(pdb.c_str(), S_IRWXU) != 0) die("cannot create " + pdb);
        build_db(fasta, pdb, "query", 0, INT32_MAX, 4000000000ull);
        Reader reads;
        reads.open(pdb, "query");
        const int n_reads = reads.info.num_seqs;
        std::vector<std::string> lines((size_t)std::max(1, n_reads));
        if (n_reads > 0 && !opened) { gpus.open(); opened = true; }
        const int n_items = (n_reads + batch - 1) / batch;
        if (n_items > 0)
            gpus.queue(n_items, [&](int g, int item) {
                const int from = item * batch, to = std::min(n_reads, from + batch), n = to - from;
                std::vector<CnsRead> R((size_t)n);
                std::vector<lesvgpu_ksw_pair> pairs((size_t)n);
                for (int k = 0; k < n; ++k) {
                    const int id = from + k;
                    const std::string name = reads.name(id);
                    const std::vector<int> svr = header_ints(name, "svr", 5);             // dir : sid : gid : sfrom : sto
                    const std::vector<int> cns = header_ints(name, "cns", 4);             // . : . : from : to
                    CnsRead& r = R[k];
                    r.read_dir = svr[0]; r.sid = svr[1]; r.sfrom = svr[3]; r.sto = svr[4];
                    r.cns_from = cns[2]; r.cns_to = cns[3];
                    r.length = reads.size(id);
                    if (r.sid != sid) die("consensus read of another contig in " + fasta + ": " + name);
                    if (r.sfrom < 0 || r.sto > subj.size(sid) || r.sfrom >= r.sto) die("reference window outside the contig: " + name);
                    if (r.cns_from < 0 || r.cns_to > r.length || r.cns_from >= r.cns_to) die("consensus range outside the read: " + name);
                    r.read.resize((size_t)r.length);
                    const size_t qoff = reads.seqs[id].seq_offset, soff = subj.seqs[sid].seq_offset;
                    for (int x = 0; x < r.length; ++x)                                     // RawReadReader_ExtractRead on strand read_dir
                        r.read[x] = r.read_dir == 0 ? pac_code(reads.pac, qoff + (size_t)x)
                                                    : (uint8_t)(3 - pac_code(reads.pac, qoff + (size_t)(r.length - 1 - x)));
                    r.window.resize((size_t)(r.sto - r.sfrom));
                    for (int x = 0; x < r.sto - r.sfrom; ++x) r.window[x] = pac_code(subj.pac, soff + (size_t)(r.sfrom + x));
                    const int longest = std::max(r.length, r.sto - r.sfrom);
                    int distance = longest;
                    distance = (int)(distance * 0.2);                                      // :205-206
                    pairs[k].query = r.read.data(); pairs[k].qlen = r.length;
                    pairs[k].target = r.window.data(); pairs[k].tlen = r.sto - r.sfrom;
                    pairs[k].max_distance = distance;
                }
                std::vector<lesvgpu_ksw_result> res((size_t)n);
                const uint32_t* pool = nullptr;
                if (lesvgpu_ksw_batch(gpus.ctx[g], pairs.data(), n, res.data(), &pool) != LESVGPU_OK)
                    die(std::string("lesvgpu_ksw_batch: ") + lesvgpu_last_error(gpus.ctx[g]));
                std::vector<int> again;
                for (int k = 0; k < n; ++k) {
                    if (!res[k].ok) { again.push_back(k); continue; }
                    sam_record(R[k], reads.name(from + k), subj.name(sid), sid, subj.size(sid), pool + res[k].cigar_off, res[k].n_cigar, lines[from + k]);
                }
                if (!again.empty()) {                                                      // :232-258: once more without a band
                    std::vector<lesvgpu_ksw_pair> p2;
                    for (int k : again) { p2.push_back(pairs[k]); p2.back().max_distance = std::max(pairs[k].qlen, pairs[k].tlen); }
                    std::vector<lesvgpu_ksw_result> r2(p2.size());
                    if (lesvgpu_ksw_batch(gpus.ctx[g], p2.data(), (int)p2.size(), r2.data(), &pool) != LESVGPU_OK)
                        die(std::string("lesvgpu_ksw_batch: ") + lesvgpu_last_error(gpus.ctx[g]));
                    for (size_t x = 0; x < again.size(); ++x) {
                        const int k = again[x];
                        if (!r2[x].ok) { fprintf(stderr, "[lesv-b200] ksw fail for %d\n", from + k); continue; }
                        sam_record(R[k], reads.name(from + k), subj.name(sid), sid, subj.size(sid), pool + r2[x].cigar_off, r2[x].n_cigar, lines[from + k]);
                    }
                }
            });
        size_t n_out = 0;
        for (int k = 0; k < n_reads; ++k)
            if (!lines[k].empty()) { fwrite(lines[k].data(), 1, lines[k].size(), sam); ++n_out; }
        fclose(sam);
        FILE* m = fopen((fasta + ".mapped").c_str(), "w");                                 // s_subject_sv_reads_make_mapped
        if (m) fclose(m);
        fprintf(stderr, "[lesv-b200] subject %s: %d consensus reads -> %zu SAM records\n", fixed8(sid).c_str(), n_reads, n_out);
    }
    if (opened) gpus.close();
    return 0;
}
