/* lesvgpu.h -- C ABI of liblesvgpu.so, the B200 (sm_100a) implementation of lesv's mapping /
 * SV-signature hot path.  Plain pointers and sizes only; every entry point names the reference
 * interface it stands in for (paths relative to the reference's src/).
 *
 * Conventions kept from the reference:
 *   - the caller owns all input buffers and all FILE*s; results the library allocates stay valid
 *     until the next call on the same context (or lesvgpu_ctx_destroy);
 *   - "no alignment" is a normal outcome (ok = 0), not an error; fatal conditions return a
 *     non-zero status and the driver decides (the reference's drivers HBN_ERR -> abort());
 *   - one feeder thread per context; a context is bound to one GPU.  There is NO CPU fallback:
 *     without a usable GPU lesvgpu_ctx_create fails.
 */
#ifndef LESVGPU_H
#define LESVGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lesvgpu_ctx lesvgpu_ctx;

enum {
    LESVGPU_OK = 0,
    LESVGPU_E_CUDA = 1,      /* CUDA runtime error, see lesvgpu_last_error */
    LESVGPU_E_ARG = 2,       /* invalid argument */
    LESVGPU_E_NOMEM = 3,     /* a single alignment does not fit the traceback arena */
    LESVGPU_E_NOGPU = 4      /* no CUDA device / wrong architecture */
};

/* ---- context -------------------------------------------------------------------------------- */

int lesvgpu_device_count(void);
/* arena_bytes = HBM reserved for traceback matrices (0 = up to 88 % of the free memory, grown lazily and never past 90 % of what is free at the time). */
int lesvgpu_ctx_create(int device, size_t arena_bytes, lesvgpu_ctx** ctx);
void lesvgpu_ctx_destroy(lesvgpu_ctx* ctx);
const char* lesvgpu_last_error(const lesvgpu_ctx* ctx);

/* Device-side timings of the last batch call, CUDA events on the launching stream. */
typedef struct {
    float h2d_ms;            /* host -> device copies of the inputs                    */
    float pack_ms;           /* 2-bit packing / extraction kernels                      */
    float dp_ms;             /* ksw_dp_kernel launches (DP + traceback walk)            */
    float scan_ms;           /* alignment-scan kernels (trim, identity, gap runs)       */
    float d2h_ms;            /* device -> host copies of the results                    */
    int   dp_launches;       /* ksw_dp_kernel launches inside dp_ms                     */
    int   kernel_launches;   /* all kernels launched by the call                        */
    size_t h2d_bytes, d2h_bytes;
    uint64_t dp_cells;       /* DP cells of all ksw_extd2 invocations of the call (lesvgpu_ksw_cells), failed-then-retried ones included */
} lesvgpu_timing;
void lesvgpu_last_timing(const lesvgpu_ctx* ctx, lesvgpu_timing* t);
/* Launch classes of the last K-ksw batch (lesvgpu_ksw_batch, lesvgpu_align_refine_batch, lesvgpu_svsig_batch; for
 * lesvgpu_svr_batch: of its two alignment rounds, one after the other).  The classes of a batch run concurrently on one GPU, so
 * done_ms is the time from the start of the first launch to the moment this class had finished. */
typedef struct {
    int ctas_per_alignment;  /* 1 = single CTA, 2..16 = thread-block cluster, 0 = global-memory fallback, -1 = column strips */
    int threads;             /* CTA size class (DP threads)                                          */
    int n_alignments;
    int n_ctas;              /* persistent CTAs (clusters) launched                                  */
    uint64_t cells;
    float done_ms;
} lesvgpu_class_timing;
int lesvgpu_last_class_timing(const lesvgpu_ctx* ctx, lesvgpu_class_timing* out, int cap);   /* returns the number of classes */
/* Kernels this process has launched through the library so far (all contexts, all stages). */
uint64_t lesvgpu_launch_count(void);

/* ---- pairwise dual-affine banded global alignment ------------------------------------------
 * Stands in for
 *   int nw_ksw2_extd2(Ksw2Data*, int qid, const u8* query, int qfrom, int qto, int qsize,
 *                     int sid, const u8* subject, int sfrom, int sto, int ssize,
 *                     int min_align_size, double min_ident_perc, int max_distance,
 *                     int* qoff, int* qend, int* soff, int* send, double* ident_perc,
 *                     kstring_t* qaln, kstring_t* saln)           algo/ksw2_wrapper.h:65-85
 * (algo/ksw2_wrapper.c:97-211 -> ksw_extd2_sse, algo/ksw2_extd2_sse.c:26-393, + ksw_backtrack,
 * algo/ksw2.h:121-153) in batch form, with the scoring of ksw2_extd2_set_params
 * (algo/ksw2_wrapper.c:72-95).  query/target point at the already sliced sub-ranges
 * (query + qfrom, qto - qfrom, ...), residues coded 0..3.
 */
typedef struct {
    const uint8_t* query;  int qlen;
    const uint8_t* target; int tlen;
    int max_distance;        /* band; 0 -> (int)(max(qlen,tlen) * 0.1) as ksw2_wrapper.c:121 */
} lesvgpu_ksw_pair;

typedef struct {
    int ok;                  /* what nw_ksw2_extd2 returns with min_ident_perc = 0           */
    int n_cigar;             /* BAM-coded ops (len<<4 | op; 0 = M, 1 = I, 2 = D), first op first */
    size_t cigar_off;        /* offset (in uint32) of this pair's ops in *cigar_pool          */
    int aln_size;            /* alignment columns = strlen(qaln)                             */
    int n_match;             /* identical columns; ident_perc = 100.0 * n_match / aln_size   */
} lesvgpu_ksw_result;

int lesvgpu_ksw_batch(lesvgpu_ctx* ctx, const lesvgpu_ksw_pair* pairs, int n,
                      lesvgpu_ksw_result* results, const uint32_t** cigar_pool);

/* ---- qx2svsig: per-read SV-signature extraction ---------------------------------------------
 * Stands in for the body of discover_func (app/necat2sv/find_sv_signature.c:234-380) over one
 * contig's SvRead array: extract read[qoff,qend) on its strand and contig[soff,send)
 * (RawReadReader_ExtractSubRead, corelib/raw_reads_reader.c:83-124), nw_ksw2_extd2 with band
 * (int)(dist*1.2) and, if that band cannot reach the last cell, max(qlen,slen) (:272-316),
 * truncate_align_bad_ends (algo/hbn_traceback.c:547-606), the optional TRF filter
 * (trf_array.cpp:74-83), calc_effective_ident_perc >= 70 (algo/hbn_traceback_aux.c:22-95) and
 * find_sv_signature (find_sv_signature.c:126-220).  The caller keeps the file formats, the read
 * names and the final sort_svsig_sfrom_lt / INS-then-DEL dump (:222-232, :406-407).
 */
typedef struct {            /* SvRead, app/necat2sv/sv_reads.h:11-20 (fields discover_func reads) */
    int query_id, subject_id;
    int qdir, qoff, qend, qsize;
    int soff, send;
    int dist;
} lesvgpu_svread;

typedef struct {            /* SvSignature, app/necat2sv/sv_signature.h:11-21 */
    int qid, qdir;
    int qfrom, qto, fqfrom, fqto, qsize;
    int sfrom, sto, fsfrom, fsto, ssize;
    int type;               /* 0 = DEL (eGapAlignDel), 6 = INS (eGapAlignIns), gapinfo.h:47-53 */
} lesvgpu_svsig;

typedef struct {
    int status;             /* 1 signatures scanned; 0 ksw fail; -1 truncate fail; -2 TRF; -3 eff. identity < 70 */
    int n_sig;              /* signatures this read contributed */
    int qb, qe, sb, se;     /* trimmed alignment range inside the aligned sub-sequences */
    int eff_mat, eff_len;   /* calc_effective_ident_perc = 100.0 * eff_mat / eff_len */
    int band;               /* band the alignment ran with */
} lesvgpu_svsig_read;

/* The subject volume's 2-bit .pac (CSeqDB.packed_seq / RawReadReader.packed_raw_seq_array) stays
 * resident in HBM; load it once per context.  A 3.1 Gbp genome is 0.78 GB. */
int lesvgpu_subject_load(lesvgpu_ctx* ctx, const uint8_t* subject_pac, size_t pac_bytes);

/* query_pac / query_read_offset are RawReadReader.packed_raw_seq_array / .raw_read_offset_array
 * (residue offset of the read + 1, raw_reads_reader.c:25,160); subject_seq_offset / subject_size
 * are CSeqInfo.seq_offset / .seq_size of the contig (corelib/seqdb.h:20-27).  trf_mask may be NULL
 * (TrfArray.trf_array, one byte per subject residue).  *sigs is library-owned (valid until the
 * next call), grouped by read in input order, each read's records in alignment order. */
int lesvgpu_svsig_batch(lesvgpu_ctx* ctx,
                        const uint8_t* query_pac, const size_t* query_read_offset,
                        size_t subject_seq_offset, int subject_size,
                        const lesvgpu_svread* reads, int n, int min_indel_size,
                        const char* trf_mask,
                        lesvgpu_svsig_read* per_read, const lesvgpu_svsig** sigs, int* n_sigs);

/* ---- qx2svr: align_and_refine_subseq_with_ksw in batch form -----------------------------------
 * Stands in for
 *   BOOL align_and_refine_subseq_with_ksw(RawReadReader* query_reader, RawReadReader* subject_reader,
 *            HbnTracebackData*, const int max_dist, int query_gid, int query_dir, int query_from,
 *            int query_to, int subject_gid, int subject_from, int subject_to, SubseqAlignResult*)
 *                                                       app/necat2sv/align_subseqs.c:192-342
 * i.e. extract both sub-sequences, nw_ksw2_extd2 with band max_dist (< 0: the longer length), one
 * retry with the longer length if the band was cut and max_dist < KSW_MAX_BW (:233-277),
 * truncate_align_bad_ends, the 0.8-coverage test (:307) and the two identity figures.  Called by
 * calc_m4_eff_perc_identity (find_sv_reads.c:130-164) and s_chain_dual_m4s (:341-430).
 */
typedef struct {
    int query_id, query_dir;          /* query_from/query_to are coordinates on strand query_dir */
    int query_from, query_to, query_size;
    size_t subject_seq_offset;        /* CSeqInfo.seq_offset of the contig */
    int subject_from, subject_to;
    int max_dist;
} lesvgpu_refine_job;

typedef struct {
    int ok;                           /* the function's BOOL */
    int qb, qe, sb, se;               /* SubseqAlignResult.qb.. (offsets already added back) */
    int mat, len;                     /* perc_identity = 100.0*mat/len; dist = len-mat; score = 8*len-mat */
    int eff_mat, eff_len;             /* eff_perc_identity = 100.0*eff_mat/eff_len */
    int band;
} lesvgpu_refine_result;

int lesvgpu_align_refine_batch(lesvgpu_ctx* ctx, const uint8_t* query_pac, const size_t* query_read_offset,
                               const lesvgpu_refine_job* jobs, int n, lesvgpu_refine_result* results);

typedef struct {            /* CSeqInfo.seq_offset / .seq_size, corelib/seqdb.h:20-27 */
    uint64_t seq_offset;
    uint64_t seq_size;
} lesvgpu_seq;

/* lesvgpu_svsig_batch over SvReads of SEVERAL contigs in one call: reads[i].subject_id indexes subject_seqs[]
 * (the reference loops over contigs, find_sv_signature.c:383-411; a contig with few placed reads is a batch too
 * small to fill a GPU, and a multi-GPU driver wants sub-batches of even cost, not contigs).  Each signature carries
 * its read's contig in the extra array *sig_subject (library-owned, parallel to *sigs). */
int lesvgpu_svsig_batch_contigs(lesvgpu_ctx* ctx,
                                const uint8_t* query_pac, const size_t* query_read_offset,
                                const lesvgpu_seq* subject_seqs, int n_subject,
                                const lesvgpu_svread* reads, int n, int min_indel_size,
                                const char* trf_mask,
                                lesvgpu_svsig_read* per_read, const lesvgpu_svsig** sigs, const int** sig_subject, int* n_sigs);

/* ---- qx2svr: one trusted placement per read ------------------------------------------------------
 * lesvgpu_svr_batch stands in, over the M4Record array of one query volume, for the worker loop
 *   thread_func -> get_next_m4a -> process_one_m4_list(M4Record* m4a, int m4c, vec_sv_read*, ...)
 *                                                       app/necat2sv/find_sv_reads.c:750-760, 728-747, 705-725
 * with min_seq_size / min_ident_perc / max_overhang = the tool's SVR_MIN_SEQ_SIZE, SVR_MIN_SVE_PERC_IDENTITY,
 * SVR_MAX_OVERHANG arguments (:909-922).  Records of one read must be consecutive (qx2map writes them so).
 * The selection logic runs on the host, its alignments (calc_m4_eff_perc_identity :130-164, s_chain_dual_m4s
 * :341-430) on the GPU through lesvgpu_align_refine_batch.  *reads is library-owned, ordered by query id.
 * lesvgpu_last_timing afterwards holds the sum over the alignment batches. */
typedef struct {            /* M4Record, corelib/m4_record.h:11-24 (56 bytes with tail padding) */
    int qid, qdir, qoff, qend, qsize;
    int sid, sdir, soff, send, ssize;
    double ident_perc;
    int score;
    int pad_;
} lesvgpu_m4;

typedef struct {
    int min_seq_size;
    double min_ident_perc;
    int max_overhang;
} lesvgpu_svr_opts;

typedef struct {
    size_t n_m4, n_refine_jobs, n_chain_jobs;
} lesvgpu_svr_stats;

int lesvgpu_svr_batch(lesvgpu_ctx* ctx, const uint8_t* query_pac, const size_t* query_read_offset,
                      const lesvgpu_seq* subject_seqs, int n_subject, const lesvgpu_m4* m4, size_t n_m4,
                      const lesvgpu_svr_opts* opts, const char* trf_mask,
                      const lesvgpu_svread** reads, int* n_reads, lesvgpu_svr_stats* stats);

/* ---- qx2map, candidate stage: index, seeding, chaining, candidate windows ----------------------
 * lesvgpu_index_build stands in for
 *   LookupTable* build_lookup_table(const text_t* db, int kmer_size, int window_size,
 *                                   int max_kmer_occ, int num_threads)     algo/lookup_table.h:29-33
 * over the subject volume already resident through lesvgpu_subject_load: k-mers at offsets 0, w, 2w ..
 * of every contig, sorted by (hash, offset) -- the reference's stable LSD radix sort,
 * algo/hash_list_bucket_sort.c:138-235 -- runs longer than max_kmer_occ removed (lookup_table.c:95-138).
 * The index stays in HBM; lesvgpu_index_export copies it out in the reference's layout
 * (LookupTable.offset_list, LktblKmerInfo{hash, offset_list_idx, cnt}) for parity checks.
 */
int lesvgpu_index_build(lesvgpu_ctx* ctx, const lesvgpu_seq* subject_seqs, int nseq,
                        int kmer_size, int kmer_window, int max_kmer_occ);
int lesvgpu_index_stats(const lesvgpu_ctx* ctx, uint64_t* n_offsets, uint64_t* n_distinct);
/* offset_list[n_offsets], hashes[n_distinct], starts[n_distinct + 1] (cnt = starts[i+1] - starts[i]) */
int lesvgpu_index_export(lesvgpu_ctx* ctx, uint64_t* offset_list, uint64_t* hashes, uint32_t* starts);

/* lesvgpu_map_candidates stands in, for the reads [qfrom, qto) of a query volume, for
 *   hbn_find_kmer_matches_mt(HbnWordFindData*, qfrom, qto, soff_max)       algo/hbn_word_finder.h:91-95
 * followed, per read, by the first half of process_one_query
 * (app/map/chain_and_extend_kmer_matches.c:188-232): find_candidates_for_one_context (:144-186, with
 * chaining_find_candidates, algo/chain_dp.c:273) and find_candidate_subject_subseqs
 * (app/map/hbn_find_subseq_hit.c:213).  The result is what `qx2map -outfmt subseq` prints:
 * HbnSubseqHit records (app/map/hbn_subseq_hit.h:10-21) in the reference's order, read by read.
 * query_pac / query_seqs describe the whole query volume (CSeqDB.packed_seq, seq_info_list).
 */
typedef struct {
    int min_ddfs;           /* -min_ddfs, default 3  (chain min_cnt; min_score = (int)(min_ddfs*k*0.8)) */
    int hitlist_size;       /* -max_target_seqs                                                        */
    /* extension stage (lesvgpu_map_batch); 0 = the reference's default */
    int max_hsps;           /* -max_hsps, default 5 for task rm (cmdline_args.cpp:692-698)             */
    int memsc_kmer_size;    /* -memsc_kmer_size 10, -memsc_kmer_window 10, -memsc_score 30              */
    int memsc_kmer_window;
    int memsc_score;
    int qcov_hsp_res;       /* -qcov_hsp_res 100: shortest alignment kept                               */
    double perc_identity;   /* -perc_identity 0.0                                                       */
    int skip_overhang;      /* -skip_overhang                                                           */
    int want_ops;           /* copy the alignment columns back (needed for -cigar / -md / sam / paf)    */
} lesvgpu_map_opts;

typedef struct {            /* HbnSubseqHit */
    int qid, qdir, qoff;
    int sid, soff;
    size_t sfrom, sto;
    int score;
} lesvgpu_subseq_hit;

typedef struct {            /* HbnInitHit as fill_init_hit_info leaves it (fields read downstream) */
    int qid, qdir, qoff, qbeg, qend;
    int sid;
    int64_t soff, sbeg, send;   /* contig-relative */
    int score;
} lesvgpu_init_hit;

typedef struct {            /* DDFKmerMatch, algo/hbn_word_finder.h:21-25 */
    int context, qoff;
    int64_t soff;
} lesvgpu_kmer_match;

typedef struct {
    float h2d_ms, lookup_ms, emit_ms, sort_ms, chain_ms, d2h_ms;
    uint64_t n_kmers, n_matches, n_segments, n_hits;
    size_t h2d_bytes, d2h_bytes;
} lesvgpu_map_timing;

/* One HSP as s_hbn_extend_subject_subseq_hit_list fills BlastHSP (app/map/hbn_extend_subseq_hit.c:279-345):
 * perc_identity = 100.0 * n_match / aln_len, raw_score = 8 * aln_len - n_match
 * (calc_ident_perc, algo/hbn_traceback_aux.c:4-19).  Query coordinates are on strand qdir, as in
 * BlastHSP before dump_m4_hits flips them (chain_and_extend_kmer_matches.c:325-330). */
typedef struct {
    int qid, qdir, qoff, qend, qsize;
    int sid, soff, send;
    int64_t ssize;
    int aln_len, n_match;
    int chain_score, ddf_score;
    size_t ops_off;         /* first alignment column in *ops_pool: 0 = both bases, 1 = read base over '-', 2 = '-' over reference base */
    int flags;              /* bit 0 / 1: the left / right overhang extension (hbn_traceback.c:214-311) ran */
} lesvgpu_hsp;

typedef struct {
    float ihf_ms, sort_ms, chain_ms, trace_ms, d2h_ms;
    uint64_t n_windows, n_seeds, n_candidates;
    uint64_t n_unsupported;        /* candidates dropped because a step is not available on the GPU (0 today) */
    uint64_t n_overhang_skipped;   /* reserved (0)                                                             */
    size_t d2h_bytes;
    float prep_ms, gap_ms, finish_ms;   /* inside trace_ms: MEMs + trace seeds, the edlib_nw gap alignments, assembly + end extension */
    uint64_t n_gap_jobs;                /* gaps between trace seeds (one edlib_nw each) */
    uint64_t n_gaps_class[4];           /* by size class: 1 block x <=128 columns, <=256, <=512 block-columns, larger / Hirschberg */
} lesvgpu_extend_timing;

/* lesvgpu_map_batch stands in, for the reads [qfrom, qto) of a query volume, for
 *   hbn_find_kmer_matches_mt(...) + chain_and_extend_km_mt(HbnWordFindData*, const HbnProgramOptions*, CSeqDB* qvol,
 *       CSeqDB* svol, int qfrom, int qto, FILE* out, FILE* backup_out)   app/map/chain_and_extend_kmer_matches.h:11-19
 * up to, but not including, the text formatting of dump_m4_hits: HSPs read by read, inside a read contig
 * group by contig group, inside a group by raw_score (remove_contained_hsp).  The caller prints them
 * (print_one_m4_result, app/map/mecat_results.c:256). */
int lesvgpu_map_batch(lesvgpu_ctx* ctx, const uint8_t* query_pac, const lesvgpu_seq* query_seqs,
                      int qfrom, int qto, const lesvgpu_map_opts* opts,
                      const lesvgpu_hsp** hsps, const int** first, const uint8_t** ops_pool);
void lesvgpu_map_last_extend_timing(const lesvgpu_ctx* ctx, lesvgpu_extend_timing* t);

/* *windows / *first are library-owned: windows of read qfrom+i are (*windows)[(*first)[i] .. (*first)[i+1]). */
int lesvgpu_map_candidates(lesvgpu_ctx* ctx, const uint8_t* query_pac, const lesvgpu_seq* query_seqs,
                           int qfrom, int qto, const lesvgpu_map_opts* opts,
                           const lesvgpu_subseq_hit** windows, const int** first);
/* intermediate results of the last lesvgpu_map_candidates call (parity checks, later stages) */
int lesvgpu_map_last_init_hits(const lesvgpu_ctx* ctx, const lesvgpu_init_hit** hits, const int** first);
int lesvgpu_map_last_matches(lesvgpu_ctx* ctx, int qfrom, const lesvgpu_kmer_match** matches, size_t* n);
void lesvgpu_map_last_timing(const lesvgpu_ctx* ctx, lesvgpu_map_timing* t);

/* Test hook: the warp-cooperative edlib_nw of the mapper's extension stage on one pair (residues 0..3):
 * alignment columns as in lesvgpu_hsp.ops_off's pool (0 = both, 1 = query base over '-', 2 = '-' over target base),
 * the output of edlib_nw (algo/edlib_wrapper.c:32-111).  ops_out: n + m bytes. */
int lesvgpu_debug_edlib_nw(lesvgpu_ctx* ctx, const uint8_t* query, int n, const uint8_t* target, int m,
                           uint8_t* ops_out, int* n_ops);
/* the same through the thread-block variant used for giant gaps (strips pipelined over the block's warps) */
int lesvgpu_debug_edlib_nw_block(lesvgpu_ctx* ctx, const uint8_t* query, int n, const uint8_t* target, int m,
                                 uint8_t* ops_out, int* n_ops);

/* DP cells ksw_extd2_sse visits for one pair (sum over anti-diagonals of the band width,
 * algo/ksw2_extd2_sse.c:124-138); *ok = 0 if the band cuts the matrix. Host arithmetic only. */
int64_t lesvgpu_ksw_cells(int qlen, int tlen, int max_distance, int* ok);

#ifdef __cplusplus
}
#endif
#endif
