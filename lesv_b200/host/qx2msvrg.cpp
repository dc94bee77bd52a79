// qx2msvrg -- same command line, same input and output files as the reference's qx2msvrg
// (src/app/necat2sv/make_sv_read_groups.c:87-178), the step right after the hot path:
//     qx2msvrg wrk_dir svsig_dir db_dir
// Per contig it reads subject_%08d.sv_signature (what qx2svsig wrote: INS block then DEL block, each sorted by sfrom),
// sweeps the INS records and then the DEL records into groups -- a strict pass (densest 10-bp window, members within 20 bp
// of its median record, >= 4 records) and a relaxed pass over what is left (chains of records <= 50 bp apart whose
// lengths differ by <= 10 % and <= 50 bp, >= 4 links) -- and writes subject_%08d_sv_read_group: per group a header line
// "<group id>\t<count>" followed by the records, then touches the .created sentinel.
// Host-only: a sequential sweep over a few thousand records per contig; nothing here is GPU work.
#include "lesv_db.hpp"

#include <fstream>
#include <sstream>
#include <sys/stat.h>
#include <unistd.h>

using namespace lesvdb;

namespace {

struct SvSig {              // SvSignature, app/necat2sv/sv_signature.h:11-21
    int qid, qdir, qfrom, qto, fqfrom, fqto, qsize, sfrom, sto, fsfrom, fsto, ssize, type;
};
const int kDel = 0, kIns = 6, kInvalid = 7;       // eGapAlignDel / eGapAlignIns / eGapAlignInvalid, gapinfo.h:47-55

// find_one_sv_group.cpp:11-19
const int kSvsigWindow = 10, kMinSvsigCnt = 4, kMaxSvsigDist = 20;
const int kSvsigWindowRelax = 50, kMinSvsigCntRelax = 4, kMaxSvsigLenDiff = 50;
const double kMaxSvsigLenDiffRatio = 0.1;

std::vector<SvSig> load(const std::string& path, int type)
{   // load_sv_signature_array + sread_sv_signature, sv_signature.cpp:28-71: 13 or 14 tab-separated columns
    std::vector<SvSig> v;
    std::ifstream in(path);
    if (!in) die("cannot open " + path);
    std::string line;
    while (std::getline(in, line)) {
        std::vector<std::string> col(1);
        for (char c : line) { if (c == '\t') col.emplace_back(); else col.back().push_back(c); }
        if (col.size() != 13 && col.size() != 14) die("bad sv_signature line: " + line);
        int f[13];
        for (int k = 0; k < 13; ++k) f[k] = atoi(col[k].c_str());
        SvSig s{f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9], f[10], f[11], f[12]};
        if (s.type < kDel || s.type >= kInvalid) die("bad signature type in " + path);
        if (s.type == type) v.push_back(s);
    }
    return v;
}

void dump_group(SvSig* a, int n, FILE* out, int& group_id)
{   // s_dump_svsig_group, find_one_sv_group.cpp:21-33
    fprintf(out, "%d\t%d\n", group_id, n);
    ++group_id;
    for (int i = 0; i < n; ++i) {
        const SvSig& s = a[i];
        fprintf(out, "%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", s.qid, s.qdir, s.qfrom, s.qto, s.fqfrom, s.fqto, s.qsize,
                s.sfrom, s.sto, s.fsfrom, s.fsto, s.ssize, s.type);
        a[i].type = kInvalid;
    }
}

// find_next_ins_group / find_next_del_group, find_one_sv_group.cpp:106-169, 237-291 (the two are the same sweep)
void strict_group(std::vector<SvSig>& v, int& next_i, int& group_id, FILE* out)
{
    SvSig* sga = v.data();
    const int sgc = (int)v.size();
    int i = next_i, j = i + 1;
    int soff = sga[i].sfrom, send = soff + kSvsigWindow;
    while (j < sgc && sga[j].sfrom < send) ++j;
    int cnt = j - i, max_cnt = cnt, max_i = i + cnt / 2;
    while (j < sgc) {                                  // slide the window one base at a time while records keep entering it
        int reduced = 0;
        for (int k = i; k < j; ++k) reduced += sga[k].sfrom == soff;
        int k = j, added = 0;
        while (k < sgc && sga[k].sfrom == send + 1) { ++added; ++k; }
        if (added == 0) break;
        cnt += added - reduced;
        ++soff; ++send;
        i += reduced;
        j = k;
        if (cnt > max_cnt) { max_cnt = cnt; max_i = i + cnt / 2; }
    }
    int gi_from = max_i;
    while (gi_from > next_i && sga[max_i].sfrom - sga[gi_from - 1].sfrom <= kMaxSvsigDist) --gi_from;
    int gi_to = max_i + 1;
    while (gi_to < sgc && sga[gi_to].sfrom - sga[max_i].sfrom <= kMaxSvsigDist) ++gi_to;
    if (gi_to - gi_from >= kMinSvsigCnt) dump_group(sga + gi_from, gi_to - gi_from, out, group_id);
    next_i = gi_to;
}

// find_next_ins_group_relax / find_next_del_group_relax, :37-102, 173-233; the length is qto - qfrom for INS, sto - sfrom for DEL
void relaxed_group(std::vector<SvSig>& v, int& next_i, int& group_id, FILE* out, bool ins)
{
    auto len = [&](const SvSig& s) { return ins ? s.qto - s.qfrom : s.sto - s.sfrom; };
    auto same = [&](const SvSig& a, const SvSig& b) {
        const int la = len(a), lb = len(b), mx = std::max(la, lb), mn = std::min(la, lb);
        return (mx - mn) <= (mx * kMaxSvsigLenDiffRatio) && (mx - mn <= kMaxSvsigLenDiff);
    };
    const int sgc = (int)v.size(), i = next_i;
    int last_i = i, last_send = v[last_i].sfrom + kSvsigWindowRelax, cnt = 0;
    std::vector<SvSig> list;
    for (int j = last_i + 1; j < sgc; ++j) {
        if (v[j].sfrom > last_send) break;
        if (same(v[last_i], v[j])) {
            last_i = j;
            last_send = v[j].sfrom + kSvsigWindowRelax;
            ++cnt;
            list.push_back(v[j]);
        }
    }
    if (cnt < kMinSvsigCntRelax) { next_i = i + 1; return; }
    list.push_back(v[i]);
    dump_group(list.data(), (int)list.size(), out, group_id);       // marks the COPIES invalid, as the reference does
    next_i = last_i + 1;
}

}  // namespace

int main(int argc, char* argv[])
{
    if (argc != 4) {
        fprintf(stderr, "USAGE:\n%s wrk_dir svsig_dir db_dir\n", argv[0]);
        return 1;
    }
    const std::string wrk_dir = argv[1], svsig_dir = argv[2], db_dir = argv[3];
    if (access(wrk_dir.c_str(), F_OK) != 0 && mkdir(wrk_dir.c_str(), S_IRWXU) != 0) die("cannot create " + wrk_dir);
    Reader subj;
    subj.open_names(db_dir, "subject");                            // seqdb_load_num_reads
    for (int sid = 0; sid < subj.info.num_seqs; ++sid) {
        const std::string out_path = wrk_dir + "/subject_" + fixed8(sid) + "_sv_read_group";
        if (access((out_path + ".created").c_str(), F_OK) == 0) continue;
        const std::string in_path = svsig_dir + "/subject_" + fixed8(sid) + ".sv_signature";
        FILE* out = fopen(out_path.c_str(), "w");
        if (!out) die("cannot write " + out_path);
        int group_id = 0;
        for (int pass = 0; pass < 2; ++pass) {                     // INS first, then DEL (make_sv_read_groups.c:107-152)
            const bool ins = pass == 0;
            std::vector<SvSig> v = load(in_path, ins ? kIns : kDel);
            int next_i = 0;
            while (next_i < (int)v.size()) strict_group(v, next_i, group_id, out);
            std::vector<SvSig> rest;
            for (const SvSig& s : v) if (s.type != kInvalid) rest.push_back(s);
            next_i = 0;
            while (next_i < (int)rest.size()) relaxed_group(rest, next_i, group_id, out, ins);
        }
        fclose(out);
        FILE* c = fopen((out_path + ".created").c_str(), "w");
        if (c) fclose(c);
    }
    return 0;
}
