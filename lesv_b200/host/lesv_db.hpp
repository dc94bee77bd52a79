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

inline std::string fixed8(int x)
{
    char b[32];
    snprintf(b, sizeof b, "%08d", x);      // u64_to_fixed_width_string_r(x, ., HBN_DIGIT_WIDTH = 8)
    return b;
}

}  // namespace lesvdb
