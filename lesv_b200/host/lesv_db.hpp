// lesv_db.hpp -- read-only view of lesv's packed sequence database (what hbn_build_seqdb writes,
// reference src/corelib/build_db.c:17-211, loaded by corelib/seqdb.c:90-191 and
// corelib/raw_reads_reader.c:7-36).  Host-side plumbing shared by the GPU drivers; file formats
// are the reference's, byte for byte.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace lesvdb {

struct SeqInfo {            // CSeqInfo, corelib/seqdb.h:20-27
    size_t seq_offset, seq_size, hdr_offset, hdr_size, ambig_offset, ambig_size;
};
struct DbInfo {             // CSeqDBInfo, corelib/seqdb.h:29-39
    int seq_start_id, num_seqs;
    size_t db_size, seq_offset_from, seq_offset_to, hdr_offset_from, hdr_offset_to, ambig_offset_from, ambig_offset_to;
};

[[noreturn]] inline void die(const std::string& msg)
{
    fprintf(stderr, "[lesv-b200] FATAL: %s\n", msg.c_str());   // the reference's HBN_ERR aborts too
    abort();
}

inline std::vector<uint8_t> read_file(const std::string& path, size_t from = 0, size_t n = (size_t)-1)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) die("cannot open " + path);
    fseek(f, 0, SEEK_END);
    size_t sz = (size_t)ftell(f);
    if (from > sz) from = sz;
    if (n > sz - from) n = sz - from;
    std::vector<uint8_t> buf(n);
    fseek(f, (long)from, SEEK_SET);
    if (n && fread(buf.data(), 1, n, f) != n) die("short read on " + path);
    fclose(f);
    return buf;
}

// Whole-database reader (RawReadReaderNew(..., use_batch_mode = FALSE)): all volumes, .pac resident.
struct Reader {
    DbInfo info{};
    std::vector<SeqInfo> seqs;
    std::vector<char> names;
    std::vector<uint8_t> pac;
    std::vector<size_t> read_offset;     // residue offset + 1 (raw_read_offset_array)

    void open_names(const std::string& db_dir, const std::string& title)      // headers only (qx2m4x)
    {
        const std::string base = db_dir + "/" + title + ".";
        auto vi = read_file(base + "volume_info.bin", 0, sizeof(DbInfo));
        if (vi.size() != sizeof(DbInfo)) die("bad volume_info.bin in " + db_dir);
        memcpy(&info, vi.data(), sizeof(DbInfo));
        auto si = read_file(base + "seq_info", 0, sizeof(SeqInfo) * (size_t)info.num_seqs);
        seqs.resize(info.num_seqs);
        memcpy(seqs.data(), si.data(), si.size());
        auto h = read_file(base + "hdr", 0, info.hdr_offset_to);
        names.assign(h.begin(), h.end());
    }
    void open(const std::string& db_dir, const std::string& title)
    {
        const std::string base = db_dir + "/" + title + ".";
        auto vi = read_file(base + "volume_info.bin", 0, sizeof(DbInfo));
        if (vi.size() != sizeof(DbInfo)) die("bad volume_info.bin in " + db_dir);
        memcpy(&info, vi.data(), sizeof(DbInfo));         // entry 0 = whole database
        auto si = read_file(base + "seq_info", 0, sizeof(SeqInfo) * (size_t)info.num_seqs);
        seqs.resize(info.num_seqs);
        memcpy(seqs.data(), si.data(), si.size());
        auto h = read_file(base + "hdr", 0, info.hdr_offset_to);
        names.assign(h.begin(), h.end());
        pac = read_file(base + "pac", 0, (info.seq_offset_to + 3) / 4);
        read_offset.resize(info.num_seqs);
        for (int i = 0; i < info.num_seqs; ++i) read_offset[i] = seqs[i].seq_offset + 1;
    }
    const char* name(int id) const { return names.data() + seqs[id].hdr_offset; }
    int size(int id) const { return (int)seqs[id].seq_size; }
};

// Tandem-repeat mask, one byte per subject residue (TrfArrayBuild, app/necat2sv/trf_array.cpp:10-63): every line is
// exactly `name \t from \t to`; names the database does not know are skipped, everything else the reference asserts
// (three columns, from < to <= contig size) is fatal here too.
inline std::vector<char> load_trf_mask(const std::string& path, const Reader& subj)
{
    std::vector<char> mask(subj.info.num_seqs ? subj.seqs.back().seq_offset + subj.seqs.back().seq_size : 0, 0);
    FILE* f = fopen(path.c_str(), "r");
    if (!f) die("cannot open " + path);
    char* line = nullptr;
    size_t cap = 0, lineno = 0;
    ssize_t len;
    while ((len = getline(&line, &cap, f)) >= 0) {
        ++lineno;
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        std::vector<std::string> col(1);
        for (ssize_t i = 0; i < len; ++i) {
            if (line[i] == '\t') col.emplace_back();
            else col.back().push_back(line[i]);
        }
        const std::string where = path + ":" + std::to_string(lineno);
        if (col.size() != 3) die(where + ": a TRF line has exactly 3 tab-separated columns");
        int sid = -1;
        for (int i = 0; i < subj.info.num_seqs && sid < 0; ++i)
            if (col[0] == subj.name(i)) sid = i;
        if (sid < 0) continue;
        char *e1 = nullptr, *e2 = nullptr;
        const unsigned long long from = strtoull(col[1].c_str(), &e1, 10), to = strtoull(col[2].c_str(), &e2, 10);
        if (col[1].empty() || col[2].empty() || *e1 || *e2) die(where + ": bad interval");
        if (!(from < to)) die(where + ": from >= to");
        if (to > subj.seqs[sid].seq_size) die(where + ": interval ends behind the contig");
        memset(mask.data() + subj.seqs[sid].seq_offset + from, 1, (size_t)(to - from));
    }
    free(line);
    fclose(f);
    return mask;
}

inline std::string fixed8(int x)
{
    char b[32];
    snprintf(b, sizeof b, "%08d", x);      // u64_to_fixed_width_string_r(x, ., HBN_DIGIT_WIDTH = 8)
    return b;
}

}  // namespace lesvdb
