// qx2svsig (GPU) -- same command line, same input and output files as the reference's qx2svsig
// (src/app/necat2sv/find_sv_signature.c:41-57, 383-440):
//     qx2svsig wrk_dir svr_dir db_dir min_indel_size num_threads [trf_bed]
// For every contig: load subject_%08d.sv_read, run the per-read body on the GPU through
// lesvgpu_svsig_batch (include/lesvgpu.h), sort each list by sfrom and write INS then DEL records
// to subject_%08d.sv_signature, then touch the .created sentinel.  num_threads is accepted and
// ignored: the SvRead arrays of all contigs still to do are cut into sub-batches of LESV_SVSIG_BATCH reads
// (default 2048), every visible GPU pulls sub-batches from one queue (lesv_gpus.hpp; LESV_GPUS / LESV_GPU select
// devices) and the records are put back together in sub-batch order, so the files are the same for 1, 2, 4 or 8 GPUs.
#include "../../include/lesvgpu.h"
#include "lesv_db.hpp"
#include "lesv_gpus.hpp"

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
    const int batch = std::max(1, getenv("LESV_SVSIG_BATCH") ? atoi(getenv("LESV_SVSIG_BATCH")) : 2048);
    if (access(wrk_dir.c_str(), F_OK) != 0 && mkdir(wrk_dir.c_str(), S_IRWXU) != 0) die("cannot create " + wrk_dir);

    Reader subj, query;
    subj.open(db_dir, "subject");
    query.open(db_dir, "query");
    std::vector<char> trf;
    if (trf_path) trf = load_trf_mask(trf_path, subj);
    std::vector<lesvgpu_seq> sseqs(subj.info.num_seqs);
    for (int i = 0; i < subj.info.num_seqs; ++i) { sseqs[i].seq_offset = subj.seqs[i].seq_offset; sseqs[i].seq_size = subj.seqs[i].seq_size; }

    // every contig that is not done yet (resume, :386-389), its SvReads appended to one list in contig order
    std::vector<int> todo;
    std::vector<size_t> n_reads(subj.info.num_seqs, 0);
    std::vector<lesvgpu_svread> reads;
    for (int sid = 0; sid < subj.info.num_seqs; ++sid) {
        const std::string out_path = wrk_dir + "/subject_" + fixed8(sid) + ".sv_signature";
        if (access((out_path + ".created").c_str(), F_OK) == 0) continue;
        todo.push_back(sid);
        const std::string in_path = svr_dir + "/subject_" + fixed8(sid) + ".sv_read";
        if (access(in_path.c_str(), F_OK) != 0) continue;
        for (lesvgpu_svread r : load_sv_reads(in_path)) { r.subject_id = sid; reads.push_back(r); ++n_reads[sid]; }
    }
    if (todo.empty()) return 0;

    const int n_items = (int)((reads.size() + (size_t)batch - 1) / (size_t)batch);
    struct Part { std::vector<lesvgpu_svsig> sigs; std::vector<int> sid; };
    std::vector<Part> parts(n_items);
    if (n_items > 0) {
        lesvgpus::Pool gpus;
        gpus.open();
        gpus.each([&](int g) {
            if (lesvgpu_subject_load(gpus.ctx[g], subj.pac.data(), subj.pac.size()) != LESVGPU_OK) die(lesvgpu_last_error(gpus.ctx[g]));
        });
        gpus.queue(n_items, [&](int g, int item) {
            const size_t from = (size_t)item * (size_t)batch;
            const int n = (int)std::min<size_t>((size_t)batch, reads.size() - from);
            std::vector<lesvgpu_svsig_read> per((size_t)n);
            const lesvgpu_svsig* sigs = nullptr;
            const int* sig_sid = nullptr;
            int n_sigs = 0;
            const int rc = lesvgpu_svsig_batch_contigs(gpus.ctx[g], query.pac.data(), query.read_offset.data(), sseqs.data(), (int)sseqs.size(),
                                                       reads.data() + from, n, min_indel, trf.empty() ? nullptr : trf.data(), per.data(),
                                                       &sigs, &sig_sid, &n_sigs);
            if (rc != LESVGPU_OK) die(std::string("lesvgpu_svsig_batch_contigs: ") + lesvgpu_last_error(gpus.ctx[g]));
            parts[item].sigs.assign(sigs, sigs + n_sigs);
            parts[item].sid.assign(sig_sid, sig_sid + n_sigs);
        });
        fprintf(stderr, "[lesv-b200] %zu sv reads of %zu contigs in %d sub-batches on %d GPU(s)\n", reads.size(), todo.size(), n_items, gpus.size());
        gpus.close();
    }
    std::vector<std::vector<lesvgpu_svsig>> ins(subj.info.num_seqs), del(subj.info.num_seqs);
    for (const Part& p : parts)
        for (size_t k = 0; k < p.sigs.size(); ++k) (p.sigs[k].type == 6 ? ins : del)[p.sid[k]].push_back(p.sigs[k]);
    for (int sid : todo) {
        const std::string out_path = wrk_dir + "/subject_" + fixed8(sid) + ".sv_signature";
        // sort_svsig_sfrom_lt (sv_signature.cpp:23-26) is a key-only unstable sort; a stable sort on
        // the same key yields one of its admissible orders
        auto by_sfrom = [](const lesvgpu_svsig& a, const lesvgpu_svsig& b) { return a.sfrom < b.sfrom; };
        std::stable_sort(ins[sid].begin(), ins[sid].end(), by_sfrom);
        std::stable_sort(del[sid].begin(), del[sid].end(), by_sfrom);
        FILE* out = fopen(out_path.c_str(), "w");
        if (!out) die("cannot write " + out_path);
        for (const auto* lst : {&ins[sid], &del[sid]})
            for (const auto& s : *lst)                                          // dump_svsig, sv_signature.h:27-45
                fprintf(out, "%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%s\n", s.qid, s.qdir, s.qfrom, s.qto,
                        s.fqfrom, s.fqto, s.qsize, s.sfrom, s.sto, s.fsfrom, s.fsto, s.ssize, s.type, query.name(s.qid));
        fclose(out);
        FILE* c = fopen((out_path + ".created").c_str(), "w");
        if (c) fclose(c);
        fprintf(stderr, "[lesv-b200] subject %s: %zu sv reads -> %zu INS + %zu DEL signatures\n",
                fixed8(sid).c_str(), n_reads[sid], ins[sid].size(), del[sid].size());
    }
    return 0;
}
