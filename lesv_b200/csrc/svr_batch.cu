// svr_batch.cu -- qx2svr's per-volume selection of one trusted placement per read, as one batch call.
//
// Replaces (paths relative to the reference's src/): thread_func / process_one_m4_list and its helpers,
// app/necat2sv/find_sv_reads.c:705-760 + :39-584 (set_sv_read_from_m4, m4_is_complete, find_complete_m4,
// remove_contained_m4s, remove_repeat_m4s, find_dual_m4s, s_chain_dual_m4s).  The branchy selection logic is
// host code restated from that file; every align_and_refine_subseq_with_ksw call it needs goes to the GPU
// in two batches (lesvgpu_align_refine_batch): round 1 = all calc_m4_eff_perc_identity calls (:130-164),
// round 2 = the dual-HSP chainings that survive round 1 (s_chain_dual_m4s, :341-430).
#include "ctx_internal.cuh"

#include <algorithm>
#include <cstdlib>
#include <utility>

namespace {

typedef lesvgpu_m4 M4;
static_assert(sizeof(M4) == 56, "M4Record layout");
typedef lesvgpu_svread SvRead;

static thread_local int g_min_seq_size = 3000;
static thread_local double g_min_ident_perc = 70.0;
static thread_local int g_max_overhang = 300;
static thread_local const char* g_trf = nullptr;
static thread_local const lesvgpu_seq* g_subj = nullptr;

static bool in_trf(int sid, size_t from, size_t to)
{   // fall_in_trf_subseq, trf_array.cpp:74-83
    if (!g_trf) return false;
    const size_t off = g_subj[sid].seq_offset;
    size_t cnt = 0;
    for (size_t i = from + off; i < to + off; ++i) cnt += g_trf[i];
    return to - from - cnt <= 2000;
}

static SvRead sv_read_from_m4(const M4& m)
{   // set_sv_read_from_m4, :39-53
    SvRead r;
    r.query_id = m.qid; r.qdir = m.qdir; r.qsize = m.qsize; r.qoff = m.qoff; r.qend = m.qend;
    r.subject_id = m.sid; r.soff = m.soff; r.send = m.send;
    r.dist = (int)((m.qend - m.qoff + m.send - m.soff) * (100.0 - m.ident_perc) / 200.0);
    return r;
}

static bool m4_is_complete(const M4& m)
{   // :121-128
    return m.qoff <= g_max_overhang && m.qsize - m.qend <= g_max_overhang;
}

static void remove_contained_one_subject(M4* a, int n)
{   // s_remove_contained_m4s_one_subject, :490-522
    const int E = 200;
    for (int i = 0; i < n; ++i) {
        M4* mi = a + i;
        if (mi->qid == -1) continue;
        for (int j = i + 1; j < n; ++j) {
            M4* mj = a + j;
            if (mj->qid == -1) continue;
            if (mi->qdir != mj->qdir) continue;
            bool r = (mj->qoff + E >= mi->qoff && mj->qend <= mi->qend + E) && (mj->soff + E >= mi->soff && mj->send <= mi->send + E);
            if (r) { mj->qid = -1; continue; }
            r = (mi->qoff + E >= mj->qoff && mi->qend <= mj->qend + E) && (mi->soff + E >= mj->soff && mi->send <= mj->send + E);
            if (r) { mi->qid = -1; continue; }
        }
    }
}

static void remove_contained(std::vector<M4>& v)
{   // remove_contained_m4s, :524-543
    size_t i = 0;
    while (i < v.size()) {
        size_t j = i + 1;
        while (j < v.size() && v[j].sid == v[i].sid) ++j;
        remove_contained_one_subject(v.data() + i, (int)(j - i));
        i = j;
    }
    v.erase(std::remove_if(v.begin(), v.end(), [](const M4& m) { return m.qid == -1; }), v.end());
}

static void remove_repeat(std::vector<M4>& v)
{   // remove_repeat_m4s, :545-584
    const int E = 300;
    for (size_t i = 0; i < v.size(); ++i) {
        M4& mi = v[i];
        if (mi.qid == -1) continue;
        for (size_t j = i + 1; j < v.size(); ++j) {
            M4& mj = v[j];
            if (mj.qid == -1) continue;
            const int max_qoff = std::max(mi.qoff, mj.qoff), min_qend = std::min(mi.qend, mj.qend);
            if (max_qoff < min_qend) {
                const int x = std::abs(max_qoff - mi.qoff), y = std::abs(min_qend - mi.qend);
                const int u = std::abs(max_qoff - mj.qoff), w = std::abs(min_qend - mj.qend);
                if (x + y <= E && u + w <= E) { mi.qid = -1; mj.qid = -1; }
            }
        }
    }
    v.erase(std::remove_if(v.begin(), v.end(), [](const M4& m) { return m.qid == -1; }), v.end());
}

static lesvgpu_refine_job eff_job(const M4& m)
{   // calc_m4_eff_perc_identity, :130-147
    int qb = m.qoff, qe = m.qend;
    if (m.qdir == 1) { qb = m.qsize - m.qend; qe = m.qsize - m.qoff; }
    int max_dist = std::max(m.qend - m.qoff, m.send - m.soff);
    max_dist = (int)(max_dist * (100.0 - m.ident_perc) / 100.0 * 1.2);
    lesvgpu_refine_job j;
    j.query_id = m.qid; j.query_dir = m.qdir; j.query_from = qb; j.query_to = qe; j.query_size = m.qsize;
    j.subject_seq_offset = g_subj[m.sid].seq_offset;
    j.subject_from = m.soff; j.subject_to = m.send; j.max_dist = max_dist;
    return j;
}

static double eff_ident(const lesvgpu_refine_result& r) { return r.eff_len ? 100.0 * r.eff_mat / r.eff_len : 0.0; }

struct Group {
    std::vector<M4> m4;                 // after the containment filter(s)
    int state = 0;                      // 1 = waiting for "complete" evaluations, 2 = waiting for dual evaluations
    std::vector<M4> cmp;                // complete HSPs (state 1)
    std::vector<int> cmp_job;           // job index per complete HSP
    // state 2: per subject run, candidate left / right ends and the job (or -1) that re-scores them
    struct Side { M4 m; int job; };
    std::vector<std::pair<std::vector<Side>, std::vector<Side>>> runs;
    int chain_job = -1;
    M4 c1, c2;
};


}  // namespace

extern "C" int lesvgpu_svr_batch(lesvgpu_ctx* ctx, const uint8_t* query_pac, const size_t* query_read_offset,
                                 const lesvgpu_seq* subject_seqs, int n_subject, const lesvgpu_m4* m4a, size_t m4c,
                                 const lesvgpu_svr_opts* opts, const char* trf_mask,
                                 const lesvgpu_svread** out_reads, int* n_out, lesvgpu_svr_stats* stats)
{
    if (!ctx) return LESVGPU_E_ARG;
    if (!query_pac || !query_read_offset || !subject_seqs || n_subject <= 0 || (!m4a && m4c) || !opts || !out_reads || !n_out) {
        ctx->err = "lesvgpu_svr_batch: null argument";
        return LESVGPU_E_ARG;
    }
    g_min_seq_size = opts->min_seq_size; g_min_ident_perc = opts->min_ident_perc; g_max_overhang = opts->max_overhang;
    g_trf = trf_mask; g_subj = subject_seqs;
    std::vector<SvRead>& out = ctx->svreads;
    out.clear();
    lesvgpu_timing acc = {};
    std::vector<lesvgpu_class_timing> classes;          // launch classes of both alignment rounds
    auto add_timing = [&]() {
        classes.insert(classes.end(), ctx->classes.begin(), ctx->classes.end());
        const lesvgpu_timing& t = ctx->timing;
        acc.h2d_ms += t.h2d_ms; acc.pack_ms += t.pack_ms; acc.dp_ms += t.dp_ms; acc.scan_ms += t.scan_ms; acc.d2h_ms += t.d2h_ms;
        acc.dp_launches += t.dp_launches; acc.kernel_launches += t.kernel_launches;
        acc.h2d_bytes += t.h2d_bytes; acc.d2h_bytes += t.d2h_bytes; acc.dp_cells += t.dp_cells;
    };
    // The records usually come straight out of a .m4x file: a stale or foreign file (another db_dir) must end in a
    // diagnosed error, not in an out-of-bounds read of subject_seqs[] / the read table.
    for (size_t i = 0; i < m4c; ++i) {
        const M4& m = m4a[i];
        const bool ok = m.qid >= 0 && m.sid >= 0 && m.sid < n_subject && m.qsize > 0 && m.qoff >= 0 && m.qoff <= m.qend &&
                        m.qend <= m.qsize && m.soff >= 0 && m.soff <= m.send && (size_t)m.send <= subject_seqs[m.sid].seq_size &&
                        (m.qdir == 0 || m.qdir == 1);
        if (!ok) {
            char b[160];
            snprintf(b, sizeof b, "lesvgpu_svr_batch: m4 record %zu does not fit the database (qid %d, sid %d of %d)", i, m.qid, m.sid, n_subject);
            ctx->err = b;
            return LESVGPU_E_ARG;
        }
    }
    std::vector<Group> groups;
    std::vector<lesvgpu_refine_job> jobs;
    auto emit = [&](const M4& m) { if (!in_trf(m.sid, m.soff, m.send)) out.push_back(sv_read_from_m4(m)); };
    // ---- phase 1: host decisions that need no alignment; queue the rest (:705-725) ---------
    for (size_t i = 0; i < m4c;) {
        size_t j = i;
        while (j < m4c && m4a[j].qid == m4a[i].qid) ++j;          // get_next_m4a, :728-747
        Group g;
        g.m4.assign(m4a + i, m4a + j);
        i = j;
        if (g.m4[0].qsize < g_min_seq_size) continue;
        remove_contained(g.m4);
        // find_complete_m4, :169-271
        for (const M4& m : g.m4) if (m4_is_complete(m)) g.cmp.push_back(m);
        if (!g.cmp.empty()) {
            if (g.cmp.size() == 1 && g.cmp[0].ident_perc >= g_min_ident_perc) { emit(g.cmp[0]); continue; }
            for (size_t k = 0; k < g.cmp.size(); ++k) {           // :193 evaluates m4a + k, not cmp_m4a + k
                g.cmp_job.push_back((int)jobs.size());
                jobs.push_back(eff_job(g.m4[k]));
            }
            g.state = 1;
            groups.push_back(std::move(g));
            continue;
        }
        remove_repeat(g.m4);
        if (g.m4.empty()) continue;
        // find_dual_m4s, :433-456 -> find_dual_m4s_on_one_subject, :293-339
        for (size_t a = 0; a < g.m4.size();) {
            size_t b = a + 1;
            while (b < g.m4.size() && g.m4[b].sid == g.m4[a].sid) ++b;
            std::vector<Group::Side> lm, rm;
            for (size_t k = a; k < b; ++k) {
                const M4& m = g.m4[k];
                if (m4_is_complete(m)) continue;
                if (in_trf(m.sid, m.soff, m.send)) continue;
                if (m.qoff <= g_max_overhang) lm.push_back({m, -1});
                if (m.qsize - m.qend <= g_max_overhang) rm.push_back({m, -1});
            }
            a = b;
            if (lm.empty() || rm.empty()) continue;
            // the right side is only re-scored if some left candidate survives; queue both and
            // let phase 2 apply the same short-circuit
            for (auto* side : {&lm, &rm})
                for (auto& s : *side)
                    if (s.m.ident_perc < g_min_ident_perc) { s.job = (int)jobs.size(); jobs.push_back(eff_job(s.m)); }
            g.runs.emplace_back(std::move(lm), std::move(rm));
        }
        if (g.runs.empty()) continue;
        g.state = 2;
        groups.push_back(std::move(g));
    }
    std::vector<lesvgpu_refine_result> res(std::max<size_t>(1, jobs.size()));
    if (!jobs.empty()) {
        const int rc = lesvgpu_align_refine_batch(ctx, query_pac, query_read_offset, jobs.data(), (int)jobs.size(), res.data());
        if (rc != LESVGPU_OK) return rc;
    }
    if (!jobs.empty()) add_timing();
    // ---- phase 2 -----------------------------------------------------------------------------
    std::vector<lesvgpu_refine_job> chain_jobs;
    for (Group& g : groups) {
        if (g.state == 1) {                                        // :191-269
            std::vector<M4> keep;
            for (size_t k = 0; k < g.cmp.size(); ++k) {
                const lesvgpu_refine_result& r = res[g.cmp_job[k]];
                const double eff = r.ok ? eff_ident(r) : 0.0;
                M4 m = g.cmp[k];
                m.ident_perc = eff;
                if (r.ok && !(eff < g_min_ident_perc)) keep.push_back(m);
            }
            if (keep.empty()) continue;
            if (keep.size() == 1) { emit(keep[0]); continue; }
            // ks_introsort_m4_ident_gt: only the two largest identities matter
            std::stable_sort(keep.begin(), keep.end(), [](const M4& a, const M4& b) { return a.ident_perc > b.ident_perc; });
            if (keep[0].ident_perc - keep[1].ident_perc > 10.0) emit(keep[0]);
            continue;
        }
        // state 2: :311-339 per subject run, then :444-453
        std::vector<std::pair<M4, M4>> duals;
        for (auto& run : g.runs) {
            auto filter = [&](std::vector<Group::Side>& side) {
                std::vector<M4> o;
                for (auto& s : side) {
                    if (s.job >= 0) {
                        const lesvgpu_refine_result& r = res[s.job];
                        const double eff = r.ok ? eff_ident(r) : 0.0;     // eff_perc_identity is read even when r == 0
                        if (!r.ok || eff < g_min_ident_perc) continue;
                        s.m.ident_perc = eff;
                    }
                    o.push_back(s.m);
                }
                return o;
            };
            std::vector<M4> lm = filter(run.first);
            if (lm.empty()) continue;
            std::vector<M4> rm = filter(run.second);
            if (rm.empty()) continue;
            for (const M4& l : lm)
                for (const M4& r : rm) {                           // two_m4s_are_dual, :273-291
                    int lsb, lse, rsb;
                    if (l.soff < r.soff) { lsb = l.soff; lse = l.send; rsb = r.soff; }
                    else { lsb = r.soff; lse = r.send; rsb = l.soff; }
                    (void)lsb;
                    if (rsb <= lse || rsb - lse <= 30000) duals.emplace_back(l, r);
                }
        }
        if (duals.size() != 1) continue;
        // s_chain_dual_m4s, :341-389
        const M4 &m1 = duals[0].first, &m2 = duals[0].second;
        if (m1.qdir != m2.qdir) continue;
        int lqb, lqe, lsb, lse, rqb, rqe, rsb, rse;
        if (m1.qdir == 0) {
            lqb = m1.qoff; lqe = m1.qend; lsb = m1.soff; lse = m1.send;
            rqb = m2.qoff; rqe = m2.qend; rsb = m2.soff; rse = m2.send;
        } else {
            lqb = m2.qsize - m2.qend; lqe = m2.qsize - m2.qoff; lsb = m2.soff; lse = m2.send;
            rqb = m1.qsize - m1.qend; rqe = m1.qsize - m1.qoff; rsb = m1.soff; rse = m1.send;
        }
        if (lsb > rsb) continue;
        lesvgpu_refine_job j;
        j.query_id = m1.qid; j.query_dir = m1.qdir; j.query_size = m1.qsize;
        j.query_from = std::min(lqb, rqb); j.query_to = std::max(lqe, rqe);
        j.subject_seq_offset = subject_seqs[m1.sid].seq_offset;
        j.subject_from = std::min(lsb, rsb); j.subject_to = std::max(lse, rse);
        j.max_dist = -1;
        g.chain_job = (int)chain_jobs.size();
        g.c1 = m1; g.c2 = m2;
        chain_jobs.push_back(j);
    }
    // ---- phase 3: chained duals (:391-429) ---------------------------------------------------
    std::vector<lesvgpu_refine_result> cres(std::max<size_t>(1, chain_jobs.size()));
    if (!chain_jobs.empty()) {
        const int rc = lesvgpu_align_refine_batch(ctx, query_pac, query_read_offset, chain_jobs.data(), (int)chain_jobs.size(), cres.data());
        if (rc != LESVGPU_OK) return rc;
    }
    if (!chain_jobs.empty()) add_timing();
    for (const Group& g : groups) {
        if (g.chain_job < 0) continue;
        const lesvgpu_refine_result& r = cres[g.chain_job];
        if (!r.ok) continue;
        const double eff = eff_ident(r);
        if (eff > g.c1.ident_perc - 4.0 || eff > g.c2.ident_perc - 4.0) {
            SvRead rp;
            rp.query_id = g.c1.qid; rp.qdir = g.c1.qdir; rp.qsize = g.c1.qsize;
            if (rp.qdir == 0) { rp.qoff = r.qb; rp.qend = r.qe; }
            else { rp.qoff = g.c1.qsize - r.qe; rp.qend = g.c1.qsize - r.qb; }
            rp.subject_id = g.c1.sid; rp.soff = r.sb; rp.send = r.se;
            rp.dist = r.len - r.mat;
            out.push_back(rp);
        }
    }
    // the reference appends in thread order; any order is one of its admissible outputs. Keep
    // input (query id) order so the result is reproducible.
    std::stable_sort(out.begin(), out.end(), [](const SvRead& a, const SvRead& b) { return a.query_id < b.query_id; });
    ctx->timing = acc;
    ctx->classes = classes;
    if (stats) { stats->n_m4 = m4c; stats->n_refine_jobs = jobs.size(); stats->n_chain_jobs = chain_jobs.size(); }
    *out_reads = out.data();
    *n_out = (int)out.size();
    return LESVGPU_OK;
}
