// qx2svr (GPU) -- same command line, same input and output files as the reference's qx2svr
// (src/app/necat2sv/find_sv_reads.c:909-951):
//     qx2svr wrk_dir db_dir min_seq_size min_sve_perc_identity max_overhang num_threads [trf_bed]
// Per query volume it reads backup_results/stagebackup_results_Q%08d_D00000000.m4x, decides per read
// which placement is trusted (process_one_m4_list, :705-725) and writes Q%08d.sv_read; at the end it
// writes subject_%08d.sv_read (+ read names) and subject_sv_read_file_count.
//
// The per-volume work is one lesvgpu_svr_batch call (include/lesvgpu.h; lesv_b200/csrc/svr_batch.cu): the
// branchy selection logic runs on the host inside the library, restated from find_sv_reads.c, and every
// align_and_refine_subseq_with_ksw call it needs goes to the GPU.  This file keeps the files and the argv.
//
// Several GPUs: the m4 records of a volume are cut at read boundaries into chunks, every visible GPU pulls chunks from
// one queue (lesv_gpus.hpp; LESV_GPUS / LESV_GPU select devices) and the selections are merged in chunk order and
// ordered by query id, as the single call does -- the files do not depend on the number of GPUs.  One GPU = one chunk.
#include "../../include/lesvgpu.h"
#include "lesv_db.hpp"
#include "lesv_gpus.hpp"

#include <algorithm>
#include <fstream>
#include <sys/stat.h>
#include <unistd.h>

using namespace lesvdb;

typedef lesvgpu_svread SvRead;     // SvRead, app/necat2sv/sv_reads.h:11-20

static const std::vector<char>* g_trf = nullptr;

int main(int argc, char* argv[])
{
    if (argc != 7 && argc != 8) {
        fprintf(stderr, "USAGE:\n%s wrk_dir db_dir min_seq_size min_sve_perc_identity max_overhang num_threads [trf_file_path]\n", argv[0]);
        return 1;
    }
    const std::string wrk_dir = argv[1], db_dir = argv[2];
    lesvgpu_svr_opts sopts;
    sopts.min_seq_size = atoi(argv[3]);
    sopts.min_ident_perc = atof(argv[4]);
    sopts.max_overhang = atoi(argv[5]);
    const char* trf_path = argc == 8 ? argv[7] : nullptr;
    if (access(wrk_dir.c_str(), F_OK) != 0 && mkdir(wrk_dir.c_str(), S_IRWXU) != 0) die("cannot create " + wrk_dir);

    Reader subj, query;
    subj.open(db_dir, "subject");
    query.open(db_dir, "query");
    std::vector<lesvgpu_seq> sseqs(subj.info.num_seqs);
    for (int i = 0; i < subj.info.num_seqs; ++i) { sseqs[i].seq_offset = subj.seqs[i].seq_offset; sseqs[i].seq_size = subj.seqs[i].seq_size; }
    std::vector<char> trf;
    if (trf_path) { trf = load_trf_mask(trf_path, subj); g_trf = &trf; }
    lesvgpus::Pool gpus;
    gpus.open();
    gpus.each([&](int g) {
        if (lesvgpu_subject_load(gpus.ctx[g], subj.pac.data(), subj.pac.size()) != LESVGPU_OK) die(lesvgpu_last_error(gpus.ctx[g]));
    });

    // seqdb_load_num_volumes: volume_info.bin holds 1 + nvol records
    const int num_vols = (int)(read_file(db_dir + "/query.volume_info.bin").size() / sizeof(DbInfo)) - 1;
    std::vector<SvRead> merged(query.info.num_seqs);
    for (auto& r : merged) r.query_id = r.subject_id = -1;

    for (int qi = 0; qi < num_vols; ++qi) {
        const std::string out_path = wrk_dir + "/Q" + fixed8(qi) + ".sv_read";
        std::vector<SvRead> out;
        if (access((out_path + ".created").c_str(), F_OK) != 0) {
            auto raw = read_file(db_dir + "/backup_results/stagebackup_results_Q" + fixed8(qi) + "_D" + fixed8(0) + ".m4x");
            const lesvgpu_m4* m4a = reinterpret_cast<const lesvgpu_m4*>(raw.data());
            const size_t m4c = raw.size() / sizeof(lesvgpu_m4);
            for (size_t i = 0; i < m4c; ++i)            // a .m4x of another database must not index past the read table
                if (m4a[i].qid < 0 || m4a[i].qid >= query.info.num_seqs) die("m4x record names a read the query database does not have");
            // chunks of whole reads (records of a read are consecutive, get_next_m4a :728-747): about two per GPU, so that
            // the full-matrix alignments of a chunk still come many at a time, and at least LESV_SVR_CHUNK records each
            const size_t min_chunk = (size_t)std::max(1, getenv("LESV_SVR_CHUNK") ? atoi(getenv("LESV_SVR_CHUNK")) : 4096);
            const size_t want = gpus.size() == 1 ? m4c : std::max(min_chunk, (m4c + 2 * (size_t)gpus.size() - 1) / (2 * (size_t)gpus.size()));
            std::vector<size_t> cut{0};
            while (cut.back() < m4c) {
                size_t e = std::min(m4c, cut.back() + std::max<size_t>(1, want));
                while (e < m4c && m4a[e].qid == m4a[e - 1].qid) ++e;
                cut.push_back(e);
            }
            const int n_chunks = (int)cut.size() - 1;
            std::vector<std::vector<SvRead>> part(std::max(1, n_chunks));
            std::vector<lesvgpu_svr_stats> stats(std::max(1, n_chunks));
            gpus.queue(n_chunks, [&](int g, int c) {
                const lesvgpu_svread* sel = nullptr;
                int n_sel = 0;
                if (lesvgpu_svr_batch(gpus.ctx[g], query.pac.data(), query.read_offset.data(), sseqs.data(), (int)sseqs.size(),
                                      m4a + cut[c], cut[c + 1] - cut[c], &sopts, g_trf ? g_trf->data() : nullptr, &sel, &n_sel,
                                      &stats[c]) != LESVGPU_OK)
                    die(std::string("lesvgpu_svr_batch: ") + lesvgpu_last_error(gpus.ctx[g]));
                part[c].assign(sel, sel + n_sel);
            });
            size_t n_jobs = 0, n_chain = 0;
            for (int c = 0; c < n_chunks; ++c) {
                out.insert(out.end(), part[c].begin(), part[c].end());
                n_jobs += stats[c].n_refine_jobs; n_chain += stats[c].n_chain_jobs;
            }
            std::stable_sort(out.begin(), out.end(), [](const SvRead& a, const SvRead& b) { return a.query_id < b.query_id; });
            FILE* f = fopen(out_path.c_str(), "w");
            if (!f) die("cannot write " + out_path);
            for (const SvRead& r : out)                                    // dump_sv_read, sv_reads.h:26-45
                fprintf(f, "%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", r.query_id, r.qdir, r.qoff, r.qend, r.qsize, r.subject_id, r.soff, r.send, r.dist);
            fclose(f);
            FILE* c = fopen((out_path + ".created").c_str(), "w");
            if (c) fclose(c);
            fprintf(stderr, "[lesv-b200] query volume %d: %zu m4 -> %zu sv reads (%zu + %zu alignments in %d chunk(s) on %d GPU(s))\n",
                    qi, m4c, out.size(), n_jobs, n_chain, n_chunks, gpus.size());
        } else {
            std::ifstream in(out_path);                                     // resume: reload what was written
            SvRead r;
            while (in >> r.query_id >> r.qdir >> r.qoff >> r.qend >> r.qsize >> r.subject_id >> r.soff >> r.send >> r.dist) out.push_back(r);
        }
        for (const SvRead& r : out) merged[r.query_id] = r;                // merge_ref_pos_results, :835-851
    }
    // dump_sv_read_array, :861-897
    for (int sid = 0; sid < subj.info.num_seqs; ++sid) {
        FILE* f = fopen((wrk_dir + "/subject_" + fixed8(sid) + ".sv_read").c_str(), "w");
        if (!f) die("cannot write subject sv_read file");
        for (int k = 0; k < query.info.num_seqs; ++k) {
            const SvRead& r = merged[k];
            if (r.subject_id != sid) continue;
            fprintf(f, "%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%s\n", r.query_id, r.qdir, r.qoff, r.qend, r.qsize, r.subject_id,
                    r.soff, r.send, r.dist, query.name(r.query_id));
        }
        fclose(f);
    }
    FILE* f = fopen((wrk_dir + "/subject_sv_read_file_count").c_str(), "w");
    if (f) { fprintf(f, "%d\n", subj.info.num_seqs); fclose(f); }
    gpus.close();
    return 0;
}
