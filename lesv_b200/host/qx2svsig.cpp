// qx2svsig (GPU) -- same command line, same input and output files as the reference's qx2svsig
// (src/app/necat2sv/find_sv_signature.c:41-57, 383-440):
//     qx2svsig wrk_dir svr_dir db_dir min_indel_size num_threads [trf_bed]
// For every contig: load subject_%08d.sv_read, run the per-read body on the GPU through
// lesvgpu_svsig_batch (include/lesvgpu.h), sort each list by sfrom and write INS then DEL records
// to subject_%08d.sv_signature, then touch the .created sentinel.  num_threads is accepted and
// ignored (one feeder thread per GPU).  Environment: LESV_GPU=<device> (default 0),
// LESV_SVSIG_BATCH=<reads per call> (default 1024).
#include "../../include/lesvgpu.h"
#include "lesv_db.hpp"

#include <algorithm>
#include <fstream>
#include <sstream>
#include <sys/stat.h>
#include <unistd.h>

using namespace lesvdb;

static std::vector<lesvgpu_svread> load_sv_reads(const std::string& path)
{
    // sread_sv_read, sv_reads.cpp:11-41: 9 ints [+ 6 dual ints] [+ name]
    std::vector<lesvgpu_svread> v;
    std::ifstream in(path);
    if (!in) die("cannot open " + path);
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        lesvgpu_svread r;
        if (!(ss >> r.query_id >> r.qdir >> r.qoff >> r.qend >> r.qsize >> r.subject_id >> r.soff >> r.send >> r.dist))
            die("bad sv_read line: " + line);
        v.push_back(r);
    }
    return v;
}

int main(int argc, char* argv[])
{
    if (argc != 6 && argc != 7) {
        fprintf(stderr, "USAGE:\n%s wrk_dir svr_dir db_dir min_indel_size num_threads [trf_array_path]\n", argv[0]);
        return 1;
    }
    const std::string wrk_dir = argv[1], svr_dir = argv[2], db_dir = argv[3];
    const int min_indel = atoi(argv[4]);
    const char* trf_path = argc > 6 ? argv[6] : nullptr;
    const int device = getenv("LESV_GPU") ? atoi(getenv("LESV_GPU")) : 0;
    const int batch = getenv("LESV_SVSIG_BATCH") ? atoi(getenv("LESV_SVSIG_BATCH")) : 1024;
    if (access(wrk_dir.c_str(), F_OK) != 0 && mkdir(wrk_dir.c_str(), S_IRWXU) != 0) die("cannot create " + wrk_dir);

    Reader subj, query;
    subj.open(db_dir, "subject");
    query.open(db_dir, "query");
    std::vector<char> trf;
    if (trf_path) trf = load_trf_mask(trf_path, subj);

    lesvgpu_ctx* ctx = nullptr;
    if (lesvgpu_ctx_create(device, 0, &ctx) != LESVGPU_OK) die("no usable B200 (lesvgpu_ctx_create failed); there is no CPU path");
    if (lesvgpu_subject_load(ctx, subj.pac.data(), subj.pac.size()) != LESVGPU_OK) die(lesvgpu_last_error(ctx));

    for (int sid = 0; sid < subj.info.num_seqs; ++sid) {
        const std::string out_path = wrk_dir + "/subject_" + fixed8(sid) + ".sv_signature";
        if (access((out_path + ".created").c_str(), F_OK) == 0) continue;       // resume, :386-389
        const std::string in_path = svr_dir + "/subject_" + fixed8(sid) + ".sv_read";
        std::vector<lesvgpu_svread> reads;
        if (access(in_path.c_str(), F_OK) == 0) reads = load_sv_reads(in_path);
        std::vector<lesvgpu_svsig> ins, del;
        std::vector<lesvgpu_svsig_read> per(std::max<size_t>(1, reads.size()));
        for (size_t from = 0; from < reads.size(); from += (size_t)batch) {
            const int n = (int)std::min<size_t>((size_t)batch, reads.size() - from);
            const lesvgpu_svsig* sigs = nullptr;
            int n_sigs = 0;
            const int rc = lesvgpu_svsig_batch(ctx, query.pac.data(), query.read_offset.data(),
                                               subj.seqs[sid].seq_offset, (int)subj.seqs[sid].seq_size,
                                               reads.data() + from, n, min_indel,
                                               trf.empty() ? nullptr : trf.data(), per.data() + from, &sigs, &n_sigs);
            if (rc != LESVGPU_OK) die(std::string("lesvgpu_svsig_batch: ") + lesvgpu_last_error(ctx));
            for (int k = 0; k < n_sigs; ++k) (sigs[k].type == 6 ? ins : del).push_back(sigs[k]);
        }
        // sort_svsig_sfrom_lt (sv_signature.cpp:23-26) is a key-only unstable sort; a stable sort on
        // the same key yields one of its admissible orders
        auto by_sfrom = [](const lesvgpu_svsig& a, const lesvgpu_svsig& b) { return a.sfrom < b.sfrom; };
        std::stable_sort(ins.begin(), ins.end(), by_sfrom);
        std::stable_sort(del.begin(), del.end(), by_sfrom);
        FILE* out = fopen(out_path.c_str(), "w");
        if (!out) die("cannot write " + out_path);
        for (const auto* lst : {&ins, &del})
            for (const auto& s : *lst)                                          // dump_svsig, sv_signature.h:27-45
                fprintf(out, "%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%s\n", s.qid, s.qdir, s.qfrom, s.qto,
                        s.fqfrom, s.fqto, s.qsize, s.sfrom, s.sto, s.fsfrom, s.fsto, s.ssize, s.type, query.name(s.qid));
        fclose(out);
        FILE* c = fopen((out_path + ".created").c_str(), "w");
        if (c) fclose(c);
        fprintf(stderr, "[lesv-b200] subject %s: %zu sv reads -> %zu INS + %zu DEL signatures\n",
                fixed8(sid).c_str(), reads.size(), ins.size(), del.size());
    }
    lesvgpu_ctx_destroy(ctx);
    return 0;
}
