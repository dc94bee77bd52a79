// qx2m4x -- text m4 results of every (query volume, subject volume) pair -> binary M4Record files, the
// step between qx2map and qx2svr (reference src/app/necat2sv/m4x.cpp:13-87):
//     qx2m4x db_dir
// reads  <db_dir>/backup_results/stagebackup_results_Q%08d_D%08d   (12 tab-separated columns, :18)
// writes <db_dir>/backup_results/stagebackup_results_Q%08d_D%08d.m4x (M4Record[], corelib/m4_record.h:11-24)
// Names are turned into database-wide sequence ids (NameToIdMapSet, corelib/name2id_map.c:27-49); the
// identity column is re-parsed from its 6-significant-digit text.  Pure host
// I/O -- no GPU work in this step (the GPU qx2map can also emit the records directly: -outfmt m4x).
#include "lesv_dbbuild.hpp"

#include <string>
#include <unordered_map>

using namespace lesvdb;

namespace {

struct M4Record {          // corelib/m4_record.h:11-24
    int qid, qdir, qoff, qend, qsize, sid, sdir, soff, send, ssize;
    double ident_perc;
    int score;
};

std::unordered_map<std::string, int> name_map(const Reader& r, const char* what)
{
    std::unordered_map<std::string, int> m;
    m.reserve((size_t)r.info.num_seqs * 2);
    for (int i = 0; i < r.info.num_seqs; ++i)
        if (!m.emplace(r.name(i), i).second) die(std::string(what) + " sequence name '" + r.name(i) + "' already exists");
    return m;
}

int to_int(const std::string& s, const std::string& line)
{
    char* e = nullptr;
    const long v = strtol(s.c_str(), &e, 10);
    if (s.empty() || *e) die("bad integer '" + s + "' in m4 line: " + line);
    return (int)v;
}

}  // namespace

int main(int argc, char** argv)
{
    if (argc != 2) {
        fprintf(stderr, "USAGE:\n%s db_dir\n", argv[0]);
        return 1;
    }
    const std::string db = argv[1];
    Reader q, s;
    q.open_names(db, "query");
    s.open_names(db, "subject");
    const auto qmap = name_map(q, "query"), smap = name_map(s, "subject");
    const int nqv = num_volumes(db, "query"), nsv = num_volumes(db, "subject");
    for (int i = 0; i < nqv; ++i)
        for (int j = 0; j < nsv; ++j) {
            const std::string path = db + "/backup_results/stagebackup_results_Q" + fixed8(i) + "_D" + fixed8(j);
            FILE* in = fopen(path.c_str(), "r");
            if (!in) die("cannot open " + path);
            FILE* out = fopen((path + ".m4x").c_str(), "wb");
            if (!out) die("cannot open " + path + ".m4x");
            char* buf = nullptr;
            size_t cap = 0;
            ssize_t len;
            int n = 0;
            std::vector<std::string> c;
            while ((len = getline(&buf, &cap, in)) >= 0) {
                while (len > 0 && (buf[len - 1] == '\n' || buf[len - 1] == '\r')) buf[--len] = 0;
                const std::string line(buf, (size_t)len);
                c.clear();
                size_t a = 0;
                for (;;) {
                    const size_t b = line.find('\t', a);
                    c.push_back(line.substr(a, b == std::string::npos ? b : b - a));
                    if (b == std::string::npos) break;
                    a = b + 1;
                }
                if (c.size() != 12) die("m4 line without exactly 12 columns (written with -cigar/-md?): " + line);   // :18
                const auto qi = qmap.find(c[0]);
                const auto si = smap.find(c[1]);
                if (qi == qmap.end() || si == smap.end()) die("unknown sequence name in m4 line: " + line);
                M4Record r;
                memset(&r, 0, sizeof r);
                r.qid = qi->second; r.sid = si->second;
                r.ident_perc = strtod(c[2].c_str(), nullptr);
                r.score = to_int(c[3], line); r.qdir = to_int(c[4], line); r.qoff = to_int(c[5], line);
                r.qend = to_int(c[6], line); r.qsize = to_int(c[7], line); r.sdir = to_int(c[8], line);
                r.soff = to_int(c[9], line); r.send = to_int(c[10], line); r.ssize = to_int(c[11], line);
                if (fwrite(&r, sizeof r, 1, out) != 1) die("write error on " + path + ".m4x");
                ++n;
            }
            free(buf);
            fclose(in);
            fclose(out);
            fprintf(stderr, "[lesv-b200] %s: number of m4: %d\n", path.c_str(), n);
        }
    return 0;
}
