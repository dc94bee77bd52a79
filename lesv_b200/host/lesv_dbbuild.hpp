// lesv_dbbuild.hpp -- FASTA/FASTQ(.gz) -> lesv's packed sequence database, byte for byte what the
// reference's build_db writes (src/corelib/build_db.c:17-211, 304-389; fingerprint :214-301):
//   <title>.pac          2 bits per base, 4 per byte (base l at byte l>>2, bits (~l&3)<<1), every sequence
//                        padded to a 4-base boundary; ambiguous residues stored as A
//   <title>.hdr          first whitespace-delimited token of each header, NUL terminated
//   <title>.seq_info     CSeqInfo[] (48 B)
//   <title>.ambig        CAmbigSubseq[] (offset, residue, count) runs of equal ambiguous residues
//   <title>.volume_info.{bin,txt}   CSeqDBInfo: entry 0 = whole database, then one per volume
//   <title>.hbndb_is_built          parameters + input size, to skip rebuilding
// Host-side I/O of the drivers ("next" row f2 of SURVEY.md §8f); the path's inputs start here.
#pragma once
#include "lesv_db.hpp"

#include <cctype>
#include <climits>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

namespace lesvdb {

struct AmbigSubseq { size_t offset; char residue; int count; };   // CAmbigSubseq, corelib/seqdb.h:14-18 (16 B)

inline int nt_code(int c)                                   // nst_nt16_table, corelib/hbn_aux.c:199-216
{
    static const signed char t[26] = {0, 10, 1, 11, 16, 16, 2, 12, 16, 16, 7, 16, 6, 14, 16, 16, 16, 4, 9, 3, 16, 13, 8, 16, 5, 16};
    if (c >= 'A' && c <= 'Z') return t[c - 'A'];
    if (c >= 'a' && c <= 'z') return t[c - 'a'];
    if (c == '-') return 15;
    return 16;
}

struct LineReader {
    gzFile f = nullptr;
    std::vector<char> buf;
    size_t pos = 0, len = 0;
    bool eof = false;
    explicit LineReader(const std::string& path) : buf(1 << 20)
    {
        f = gzopen(path.c_str(), "rb");
        if (!f) die("cannot open " + path);
    }
    ~LineReader() { if (f) gzclose(f); }
    bool getline(std::string& line)
    {
        line.clear();
        for (;;) {
            if (pos == len) {
                if (eof) return !line.empty();
                const int n = gzread(f, buf.data(), (unsigned)buf.size());
                if (n <= 0) { eof = true; return !line.empty(); }
                pos = 0; len = (size_t)n;
            }
            const char* p = (const char*)memchr(buf.data() + pos, '\n', len - pos);
            if (p) {
                line.append(buf.data() + pos, p - (buf.data() + pos));
                pos = (size_t)(p - buf.data()) + 1;
                if (!line.empty() && line.back() == '\r') line.pop_back();
                return true;
            }
            line.append(buf.data() + pos, len - pos);
            pos = len;
        }
    }
};

inline std::string db_base(const std::string& dir, const std::string& title) { return dir + "/" + title + "."; }

inline size_t file_size(const std::string& path)
{
    struct stat st;
    if (stat(path.c_str(), &st) != 0) die("cannot stat " + path);
    return (size_t)st.st_size;
}

// hbndb_is_built / hbndb_make_built
inline bool db_is_built(const std::string& input, const std::string& dir, const std::string& title, int min_size, int max_seqs,
                        size_t max_res)
{
    const std::string path = db_base(dir, title) + "hbndb_is_built";
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    int len = 0, ms = 0, mq = 0, rn = 0;
    size_t mr = 0, bytes = 0;
    bool ok = fread(&len, sizeof(int), 1, f) == 1 && len >= 0 && len < (1 << 20);
    std::string old(ok ? len : 0, '\0');
    ok = ok && (len == 0 || fread(&old[0], 1, (size_t)len, f) == (size_t)len) && fread(&ms, sizeof(int), 1, f) == 1 &&
         fread(&mq, sizeof(int), 1, f) == 1 && fread(&mr, sizeof(size_t), 1, f) == 1 && fread(&rn, sizeof(int), 1, f) == 1 &&
         fread(&bytes, sizeof(size_t), 1, f) == 1;
    fclose(f);
    return ok && old == input && ms == min_size && mq == max_seqs && mr == max_res && rn == 0 && bytes == file_size(input);
}
inline void db_make_built(const std::string& input, const std::string& dir, const std::string& title, int min_size, int max_seqs,
                          size_t max_res)
{
    FILE* f = fopen((db_base(dir, title) + "hbndb_is_built").c_str(), "wb");
    if (!f) die("cannot write fingerprint in " + dir);
    const int len = (int)input.size(), rn = 0;
    const size_t bytes = file_size(input);
    fwrite(&len, sizeof(int), 1, f); fwrite(input.data(), 1, input.size(), f);
    fwrite(&min_size, sizeof(int), 1, f); fwrite(&max_seqs, sizeof(int), 1, f); fwrite(&max_res, sizeof(size_t), 1, f);
    fwrite(&rn, sizeof(int), 1, f); fwrite(&bytes, sizeof(size_t), 1, f);
    fclose(f);
}

inline void dump_info_txt(FILE* f, const DbInfo& d)
{
    fprintf(f, "%d\t%d\t%zu\t%zu\t%zu\t%zu\t%zu\t%zu\t%zu\n", d.seq_start_id, d.num_seqs, d.db_size, d.seq_offset_from,
            d.seq_offset_to, d.hdr_offset_from, d.hdr_offset_to, d.ambig_offset_from, d.ambig_offset_to);
}

// build_db for one FASTA/FASTQ(.gz) file
inline void build_db(const std::string& input, const std::string& dir, const std::string& title, int min_size, int max_seqs,
                     size_t max_res)
{
    const std::string base = db_base(dir, title);
    FILE* f_pac = fopen((base + "pac").c_str(), "wb");
    FILE* f_hdr = fopen((base + "hdr").c_str(), "wb");
    FILE* f_inf = fopen((base + "seq_info").c_str(), "wb");
    FILE* f_amb = fopen((base + "ambig").c_str(), "wb");
    if (!f_pac || !f_hdr || !f_inf || !f_amb) die("cannot create database files in " + dir);
    std::vector<SeqInfo> infos;
    size_t seq_off = 0, hdr_off = 0, amb_off = 0, res = 0;
    std::vector<uint8_t> es;
    auto flush_seq = [&](const std::string& name, const std::string& seq) {
        if ((long long)seq.size() < (long long)min_size) return;
        SeqInfo si;
        si.seq_offset = seq_off; si.seq_size = seq.size(); si.hdr_offset = hdr_off; si.ambig_offset = amb_off; si.ambig_size = 0;
        size_t n = 0;
        while (n < name.size() && !isspace((unsigned char)name[n])) ++n;
        if (n == 0) die("sequence without a name in " + input);
        si.hdr_size = n;
        fwrite(name.data(), 1, n, f_hdr); fputc('\0', f_hdr);
        hdr_off += n + 1;
        const size_t ubytes = (seq.size() + 3) >> 2;
        es.assign(ubytes, 0);
        size_t i = 0;
        while (i < seq.size()) {
            const int c = (unsigned char)seq[i], ec = nt_code(c);
            if (ec > 3) {                              // ambiguous run: stored as A (code 0) + one record
                AmbigSubseq a;
                memset(&a, 0, sizeof a);
                a.offset = i; a.residue = (char)c; a.count = 1;
                ++i;
                while (i < seq.size() && nt_code((unsigned char)seq[i]) == ec) { ++a.count; ++i; }
                ++si.ambig_size; ++amb_off;
                fwrite(&a, sizeof a, 1, f_amb);
            } else {
                es[i >> 2] |= (uint8_t)(ec << ((~i & 3) << 1));
                ++i;
            }
        }
        fwrite(es.data(), 1, ubytes, f_pac);
        fwrite(&si, sizeof si, 1, f_inf);
        infos.push_back(si);
        seq_off += ubytes << 2;
        res += seq.size();
    };
    {
        LineReader in(input);
        std::string line, name, seq;
        bool have = false;
        while (in.getline(line)) {
            if (line.empty()) continue;
            if (line[0] == '>') {
                if (have) flush_seq(name, seq);
                name = line.substr(1); seq.clear(); have = true;
            } else if (line[0] == '@' && !have) {      // FASTQ: 4-line records
                name = line.substr(1);
                if (!in.getline(seq)) break;
                std::string plus, qual;
                in.getline(plus); in.getline(qual);
                flush_seq(name, seq);
                seq.clear();
            } else if (have) {
                for (char c : line) if (!isspace((unsigned char)c)) seq.push_back(c);
            }
        }
        if (have) flush_seq(name, seq);
    }
    fclose(f_pac); fclose(f_hdr); fclose(f_inf); fclose(f_amb);
    // build_volume_info (:150-211)
    DbInfo all{};
    all.seq_start_id = 0; all.num_seqs = (int)infos.size(); all.db_size = res;
    all.seq_offset_from = 0; all.seq_offset_to = seq_off; all.hdr_offset_from = 0; all.hdr_offset_to = hdr_off;
    all.ambig_offset_from = 0; all.ambig_offset_to = amb_off;
    FILE* f_txt = fopen((base + "volume_info.txt").c_str(), "w");
    FILE* f_bin = fopen((base + "volume_info.bin").c_str(), "wb");
    if (!f_txt || !f_bin) die("cannot create volume info in " + dir);
    dump_info_txt(f_txt, all); fwrite(&all, sizeof all, 1, f_bin);
    size_t s = 0, last = 0;
    int cnt = 0;
    auto emit = [&](size_t i) {
        DbInfo d = all;
        d.db_size = s; d.num_seqs = (int)(i + 1 - last); d.seq_start_id = (int)last;
        d.seq_offset_from = infos[last].seq_offset; d.seq_offset_to = infos[i].seq_offset + infos[i].seq_size;
        d.hdr_offset_from = infos[last].hdr_offset; d.hdr_offset_to = infos[i].hdr_offset + infos[i].hdr_size + 1;
        d.ambig_offset_from = infos[last].ambig_offset; d.ambig_offset_to = infos[i].ambig_offset + infos[i].ambig_size;
        dump_info_txt(f_txt, d); fwrite(&d, sizeof d, 1, f_bin);
    };
    for (size_t i = 0; i < infos.size(); ++i) {
        s += infos[i].seq_size; ++cnt;
        if (s >= max_res || cnt >= max_seqs) { emit(i); last = i + 1; s = 0; cnt = 0; }
    }
    if (s > 0) emit(infos.size() - 1);
    fclose(f_txt); fclose(f_bin);
}

// One volume of a database (seqdb_load, corelib/seqdb.c:306-326): offsets rebased to the volume.
struct Volume {
    DbInfo info{};
    std::vector<SeqInfo> seqs;
    std::vector<char> names;
    std::vector<uint8_t> pac;
    const char* name(int id) const { return names.data() + seqs[id].hdr_offset; }
};
inline int num_volumes(const std::string& dir, const std::string& title)
{
    return (int)(file_size(db_base(dir, title) + "volume_info.bin") / sizeof(DbInfo)) - 1;
}
inline int num_seqs_total(const std::string& dir, const std::string& title)
{
    DbInfo d;
    auto b = read_file(db_base(dir, title) + "volume_info.bin", 0, sizeof d);
    memcpy(&d, b.data(), sizeof d);
    return d.num_seqs;
}
inline Volume load_volume(const std::string& dir, const std::string& title, int vol)
{
    Volume v;
    const std::string base = db_base(dir, title);
    auto b = read_file(base + "volume_info.bin", sizeof(DbInfo) * (size_t)(vol + 1), sizeof(DbInfo));
    if (b.size() != sizeof(DbInfo)) die("no such volume in " + base);
    memcpy(&v.info, b.data(), sizeof(DbInfo));
    auto si = read_file(base + "seq_info", sizeof(SeqInfo) * (size_t)v.info.seq_start_id, sizeof(SeqInfo) * (size_t)v.info.num_seqs);
    v.seqs.resize(v.info.num_seqs);
    memcpy(v.seqs.data(), si.data(), si.size());
    auto h = read_file(base + "hdr", v.info.hdr_offset_from, v.info.hdr_offset_to - v.info.hdr_offset_from);
    v.names.assign(h.begin(), h.end());
    v.pac = read_file(base + "pac", v.info.seq_offset_from >> 2, ((v.info.seq_offset_to + 3) >> 2) - (v.info.seq_offset_from >> 2));
    for (auto& s : v.seqs) { s.hdr_offset -= v.info.hdr_offset_from; s.seq_offset -= v.info.seq_offset_from; }
    return v;
}

}  // namespace lesvdb
