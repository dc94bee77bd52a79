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
#include "../../include/lesvgpu.h"
#include "lesv_db.hpp"

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
    const int device = getenv("LESV_GPU") ? atoi(getenv("LESV_GPU")) : 0;
    if (access(wrk_dir.c_str(), F_OK) != 0 && mkdir(wrk_dir.c_str(), S_IRWXU) != 0) die("cannot create " + wrk_dir);

    Reader subj, query;
    subj.open(db_dir, "subject");
    query.open(db_dir, "query");
    std::vector<lesvgpu_seq> sseqs(subj.info.num_seqs);
    for (int i = 0; i < subj.info.num_seqs; ++i) { sseqs[i].seq_offset = subj.seqs[i].seq_offset; sseqs[i].seq_size = subj.seqs[i].seq_size; }
    std::vector<char> trf;
    if (trf_path) { trf = load_trf_mask(trf_path, subj); g_trf = &trf; }
    lesvgpu_ctx* ctx = nullptr;
    if (lesvgpu_ctx_create(device, 0, &ctx) != LESVGPU_OK) die("no usable B200 (lesvgpu_ctx_create failed); there is no CPU path");
    if (lesvgpu_subject_load(ctx, subj.pac.data(), subj.pac.size()) != LESVGPU_OK) die(lesvgpu_last_error(ctx));

    // seqdb_load_num_volumes: volume_info.bin holds 1 + nvol records
    const int num_vols = (int)(read_file(db_dir + "/query.volume_info.bin").size() / sizeof(DbInfo)) - 1;
    std::vector<SvRead> merged(query.info.num_seqs);
    for (auto& r : merged) r.query_id = r.subject_id = -1;

    for (int qi = 0; qi < num_vols; ++qi) {
        const std::string out_path = wrk_dir + "/Q" + fixed8(qi) + ".sv_read";
        std::vector<SvRead> out;
        if (access((out_path + ".created").c_str(), F_OK) != 0) {
            auto raw = read_file(db_dir + "/backup_results/stagebackup_results_Q" + fixed8(qi) + "_D" + fixed8(0) + ".m4x");
            const lesvgpu_svread* sel = nullptr;
            int n_sel = 0;
            lesvgpu_svr_stats st = {};
            if (lesvgpu_svr_batch(ctx, query.pac.data(), query.read_offset.data(), sseqs.data(), (int)sseqs.size(),
                                  reinterpret_cast<const lesvgpu_m4*>(raw.data()), raw.size() / sizeof(lesvgpu_m4), &sopts,
                                  g_trf ? g_trf->data() : nullptr, &sel, &n_sel, &st) != LESVGPU_OK)
                die(std::string("lesvgpu_svr_batch: ") + lesvgpu_last_error(ctx));
            out.assign(sel, sel + n_sel);
            const size_t m4c = st.n_m4;
            const size_t n_jobs = st.n_refine_jobs, n_chain = st.n_chain_jobs;
            FILE* f = fopen(out_path.c_str(), "w");
            if (!f) die("cannot write " + out_path);
            for (const SvRead& r : out)                                    // dump_sv_read, sv_reads.h:26-45
                fprintf(f, "%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", r.query_id, r.qdir, r.qoff, r.qend, r.qsize, r.subject_id, r.soff, r.send, r.dist);
            fclose(f);
            FILE* c = fopen((out_path + ".created").c_str(), "w");
            if (c) fclose(c);
            fprintf(stderr, "[lesv-b200] query volume %d: %zu m4 -> %zu sv reads (%zu + %zu alignments on the GPU)\n",
                    qi, m4c, out.size(), n_jobs, n_chain);
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
    lesvgpu_ctx_destroy(ctx);
    return 0;
}
