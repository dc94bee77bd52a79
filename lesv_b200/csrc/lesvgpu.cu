// lesvgpu.cu -- host side of the C ABI (include/lesvgpu.h): contexts, batching, scheduling.
//
// The reference calls its aligner from T pthreads, one pair at a time.  Here one
// feeder thread hands a whole batch to one GPU: pairs are binned by band width into the thread
// counts the DP kernel is instantiated for, each bin is sorted by decreasing traceback size and
// pulled by persistent CTAs, and every CTA reuses one traceback arena in HBM.
#include "ctx_internal.cuh"
#include "gpu_prims.cuh"
#include "ksw_core.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace lesvksw;

extern "C" int lesvgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int lesvgpu_ctx_create(int device, size_t arena_bytes, lesvgpu_ctx** out)
{
    if (!out) return LESVGPU_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return LESVGPU_E_NOGPU;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return LESVGPU_E_NOGPU;
    if (prop.major != 10) return LESVGPU_E_NOGPU;          // sm_100a code only
    if (cudaSetDevice(device) != cudaSuccess) return LESVGPU_E_NOGPU;
    lesvgpu_ctx* ctx = new lesvgpu_ctx;
    ctx->device = device;
    ctx->n_sm = prop.multiProcessorCount;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    ctx->arena_limit = arena_bytes ? arena_bytes : (size_t)(free_b * 0.88);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LESVGPU_E_CUDA; }
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    // launch classes are issued largest shape first on side[0], side[1], ...: descending priority lets the
    // cluster launches take their SMs before the bulk of single-CTA alignments fills the machine
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    for (int k = 0; k < kSideStreams; ++k) {
        const int pr = std::min(prio_least, prio_greatest + k);
        cudaStreamCreateWithPriority(&ctx->side[k], cudaStreamNonBlocking, pr);
    }
    for (auto& e : ctx->side_done) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    *out = ctx;
    return LESVGPU_OK;
}

extern "C" void lesvgpu_ctx_destroy(lesvgpu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->d_raw.release(); ctx->d_arena.release(); ctx->d_subject.release(); ctx->d_pool.release();
    ctx->d_cigar.release(); ctx->d_scratch.release(); ctx->d_jobs.release(); ctx->d_outs.release();
    ctx->d_runs.release(); ctx->d_order.release(); ctx->d_counter.release(); ctx->d_aoff.release();
    ctx->d_sjobs.release(); ctx->d_edges.release(); ctx->d_sprog.release();
    ctx->h_raw.release(); ctx->h_cigar.release(); ctx->h_outs.release(); ctx->h_runs.release();
    for (auto& e : ctx->ev) cudaEventDestroy(e);
    for (auto& e : ctx->side_done) cudaEventDestroy(e);
    for (auto& st : ctx->side) cudaStreamDestroy(st);
    cudaStreamDestroy(ctx->stream);
    if (ctx->map) lesvmap_state_free(ctx->map);
    delete ctx;
}

extern "C" uint64_t lesvgpu_launch_count(void) { return g_lesv_launches; }
extern "C" const char* lesvgpu_last_error(const lesvgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void lesvgpu_last_timing(const lesvgpu_ctx* ctx, lesvgpu_timing* t) { if (ctx && t) *t = ctx->timing; }
extern "C" int lesvgpu_last_class_timing(const lesvgpu_ctx* ctx, lesvgpu_class_timing* out, int cap)
{
    if (!ctx) return 0;
    const int n = (int)ctx->classes.size();
    for (int i = 0; i < n && i < cap && out; ++i) out[i] = ctx->classes[i];
    return n;
}

static inline int effective_band(int qlen, int tlen, int max_distance)
{
    // ksw2_wrapper.c:121: a zero band means 10 % of the longer sequence (double product, truncated)
    if (max_distance == 0) max_distance = (int)((qlen > tlen ? qlen : tlen) * 0.1);
    // ksw2_extd2_sse.c:82: negative = no band
    if (max_distance < 0) max_distance = qlen > tlen ? qlen : tlen;
    return max_distance;
}

static inline bool band_reaches_end(int qlen, int tlen, int w)
{
    // st0 <= en0 on every anti-diagonal (ksw2_extd2_sse.c:130-137)  <=>  |qlen - tlen| <= w, and a
    // zero band only survives a 1x1 matrix.
    const int d = qlen > tlen ? qlen - tlen : tlen - qlen;
    if (d > w) return false;
    if (w == 0 && qlen + tlen - 1 > 1) return false;
    return true;
}

// Sum of en0 - st0 + 1 over all anti-diagonals when none is empty, in closed form: both bounds are piecewise
// {constant, r + c, (r + c) >> 1} with at most six breakpoints; two rows on either side of every breakpoint are
// evaluated directly, the interior of each segment with the piece that is active there.
static inline int64_t sum_lin(int64_t a, int64_t b) { return a > b ? 0 : (a + b) * (b - a + 1) / 2; }            // sum r, r=a..b
static inline int64_t half_prefix(int64_t x) { return x < 0 ? 0 : (x / 2) * ((x + 1) / 2); }                     // sum_{s=0..x} floor(s/2)
static inline int64_t sum_half(int64_t a, int64_t b, int64_t c) { return a > b ? 0 : half_prefix(b + c) - half_prefix(a + c - 1); } // needs a + c >= 0
static int64_t cells_fast(int ql, int tl, int w)
{
    const int nr = ql + tl - 1;
    long long cand[] = { (long long)w, 2LL * tl - 2 - w, (long long)ql - 1, (long long)w - 1, 2LL * ql - w - 1, (long long)tl - 1 };
    int64_t total = 0;
    // collect breakpoints
    int pts[16]; int np = 0; pts[np++] = 0; pts[np++] = nr;
    for (long long x : cand) if (x > 0 && x < nr) pts[np++] = (int)x;
    std::sort(pts, pts + np);
    auto st0 = [&](int r) { int st = r - ql + 1; st = st > 0 ? st : 0; const int s2 = (r - w + 1) >> 1; return st > s2 ? st : s2; };
    auto en0 = [&](int r) { int en = r < tl - 1 ? r : tl - 1; const int e2 = (r + w) >> 1; return en < e2 ? en : e2; };
    for (int i = 0; i + 1 < np; ++i) {
        int a = pts[i], b = pts[i + 1] - 1;          // rows a..b
        if (a > b) continue;
        // margins by brute force
        while (a <= b && (a < pts[i] + 2)) { total += en0(a) - st0(a) + 1; ++a; }
        int cnt = 0;
        while (a <= b && cnt < 2) { total += en0(b) - st0(b) + 1; --b; ++cnt; }
        if (a > b) continue;
        const int mid = a + (b - a) / 2;
        // which piece is active (constant over the interior)
        int64_t se, ss;
        { const int e = en0(mid);
          if (e == tl - 1 && en0(a) == tl - 1 && en0(b) == tl - 1) se = (int64_t)(tl - 1) * (b - a + 1);
          else if (e == mid && en0(a) == a && en0(b) == b) se = sum_lin(a, b);
          else se = sum_half(a, b, w); }
        { const int s = st0(mid);
          if (s == 0 && st0(a) == 0 && st0(b) == 0) ss = 0;
          else if (s == mid - ql + 1 && st0(a) == a - ql + 1 && st0(b) == b - ql + 1) ss = sum_lin(a, b) - (int64_t)(ql - 1) * (b - a + 1);
          else ss = sum_half(a, b, -(int64_t)w + 1); }
        total += se - ss + (b - a + 1);
    }
    return total;
}

extern "C" int64_t lesvgpu_ksw_cells(int qlen, int tlen, int max_distance, int* ok)
{
    const int w = effective_band(qlen, tlen, max_distance);
    int64_t cells = 0;
    if (ok) *ok = 1;
    if (qlen > 0 && tlen > 0 && band_reaches_end(qlen, tlen, w)) return cells_fast(qlen, tlen, w);
    for (int r = 0; r < qlen + tlen - 1; ++r) {
        const RowGeom g = row_geom(r, qlen, tlen, w);
        if (g.empty()) { if (ok) *ok = 0; break; }
        cells += g.en0 - g.st0 + 1;
    }
    return cells;
}

// ---------------------------------------------------------------------------------------------
// shared runner: jobs[] have qlen/tlen/w/src/seq offsets/min_indel filled in and cigar_cap set to
// 0 or 1 (1 = wants a CIGAR); pool/cigar/run offsets are assigned here.
// ---------------------------------------------------------------------------------------------
static int run_jobs(lesvgpu_ctx* ctx, std::vector<KswJob>& jobs, size_t raw_bytes, bool want_cigar, bool want_runs)
{
    // A launch class = (CTAs per alignment, threads per CTA).  One 16-column block per thread: the
    // smallest shape that holds the band window wastes the fewest lanes.  Bands wider than 640
    // blocks run on a thread-block cluster (2/4/8/16 CTAs passing carries through DSMEM).
    // `tier` is the CTA size of the launch (it fixes how many CTAs share an SM); `nt` what this
    // alignment actually uses of it.
    struct Shape { int ncta, tier, nt; };
    auto shape_of = [](int wb) -> Shape {
        static const int tiers[] = {32, 64, 96, 128, 160, kTierLow, kTierMid, kTierTop, kTierXL};
        static const bool groups = getenv("LESV_KSW_NO_GROUPS") == nullptr;
        if (groups && wb <= 8) return {1, 8, 8};             // several alignments per warp (ksw_dp_group_kernel)
        if (groups && wb <= 16) return {1, 16, 16};
        if (wb <= kTierXL) {
            const int nt = (wb + 31) / 32 * 32;
            for (int t : tiers) if (nt <= t) return {1, t, nt};
        }
        for (int ncta = 2; ncta <= 16; ncta *= 2)
            if (wb <= ncta * 640) return {ncta, 640, ((wb + ncta - 1) / ncta + 31) / 32 * 32};
        return {0, 0, 0};                                // wider than any cluster: global-memory fallback
    };
    const int nj = (int)jobs.size();
    const auto t_enter = std::chrono::steady_clock::now();
    struct Launch {
        int ncta, nt, first, count, n_clusters, max_slots, fair;
        size_t stride, smem, arena_off, arena_bytes;
        double work;
        std::vector<int> ids;
    };
    std::vector<Launch> launches;
    std::vector<size_t> tb_bytes(nj);
    size_t pool_words = 0, cigar_words = 0;
    double total_work = 0;
    // Full-matrix alignments (band >= both lengths, wider than one CTA) run as column strips with an arena per alignment
    // (ksw_dp_strip_kernel) in groups whose tracebacks fit into three quarters of the arena budget, one launch per group.
    static const bool use_strips = getenv("LESV_KSW_NO_STRIPS") == nullptr;
    std::vector<int> strip_ids;
    std::vector<char> is_strip(nj, 0);
    size_t strip_arena = 0;
    // what the arena may grow to right now: the configured limit, but never more than 90 % of what is free
    // (the subject, the index and the mapper's buffers live on the same device; other processes may, too)
    size_t arena_limit = ctx->arena_limit;
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
            arena_limit = std::min(arena_limit, (size_t)((free_b + ctx->d_arena.cap) * 0.90));
    }
    std::vector<std::pair<size_t, size_t>> strip_groups;            // [first, last) into strip_ids, one launch each
    if (use_strips) {
        std::vector<int> cand;
        for (int k = 0; k < nj; ++k) {
            const KswJob& J = jobs[k];
            if (J.w >= J.qlen && J.w >= J.tlen && window_blocks(J.tlen, J.w) > kTierXL) cand.push_back(k);
        }
        // A full-matrix alignment of 50 kb x 60 kb is 3 GB of traceback, and every alignment of a strip launch needs its
        // own arena: they are processed in groups that fit the budget -- all of it when the batch holds nothing else (the
        // chaining round of qx2svr), three quarters next to other launch classes -- one launch per group, the groups about
        // equally large (a launch costs the critical path of one alignment however few alignments it holds).
        const size_t cap = (int)cand.size() == nj ? arena_limit / 20 * 19 : arena_limit / 4 * 3;
        auto tb_of = [&](int k) {                        // rows folded modulo qlen (ksw_kernels.cuh: fold_tb_bytes)
            const KswJob& J = jobs[k];
            return (fold_tb_bytes(J.qlen, J.tlen) + 255) & ~(size_t)255;
        };
        size_t total = 0;
        for (int k : cand) if (tb_of(k) <= cap) total += tb_of(k);
        const size_t n_groups = std::max<size_t>(1, (total + cap - 1) / cap);
        const size_t target = (total + n_groups - 1) / n_groups;
        size_t group = 0, first = 0;
        for (int k : cand) {
            const size_t tb = tb_of(k);
            if (tb > cap) continue;                         // a single alignment that large takes the cluster / wide route
            if (group > 0 && (group + tb > cap || group >= target)) {
                strip_groups.emplace_back(first, strip_ids.size());
                first = strip_ids.size(); group = 0;
            }
            group += tb;
            strip_arena = std::max(strip_arena, group);
            strip_ids.push_back(k); is_strip[k] = 1;
        }
        if (strip_ids.size() > first) strip_groups.emplace_back(first, strip_ids.size());
    }
    for (int k = 0; k < nj; ++k) {
        KswJob& J = jobs[k];
        const Shape sh = shape_of(window_blocks(J.tlen, J.w));
        J.tb_blocks = tb_row_blocks(J.qlen, J.tlen, J.w);
        J.tp_off = (uint32_t)pool_words; pool_words += tp_words(J.tlen);
        J.qrp_off = (uint32_t)pool_words; pool_words += qrp_words(J.qlen);
        J.cigar_off = cigar_words;
        J.cigar_cap = want_cigar ? (uint32_t)(J.qlen + J.tlen + 2) : 0u;
        cigar_words += J.cigar_cap;
        J.sig_off = (uint32_t)k * kSigCap;
        if (pool_words > 0xFFFFFFF0ull) { ctx->err = "batch too large (packed pool offsets are 32-bit)"; return LESVGPU_E_ARG; }
        tb_bytes[k] = is_strip[k] ? fold_tb_bytes(J.qlen, J.tlen) : (size_t)(J.qlen + J.tlen - 1) * (size_t)J.tb_blocks * 16;
        if (tb_bytes[k] > arena_limit) { ctx->err = "one alignment's traceback does not fit the arena"; return LESVGPU_E_NOMEM; }
        if (is_strip[k]) { J.nt = kStripBlocks; continue; }
        size_t li = 0;
        for (; li < launches.size(); ++li)
            if (launches[li].ncta == sh.ncta && launches[li].nt == sh.tier) break;
        if (li == launches.size()) { Launch L{}; L.ncta = sh.ncta; L.nt = sh.tier; launches.push_back(L); }
        J.nt = sh.nt;
        launches[li].ids.push_back(k);
        launches[li].work += (double)tb_bytes[k];
        total_work += (double)tb_bytes[k];
    }
    // biggest alignments first, both across launches and inside each launch's queue
    std::sort(launches.begin(), launches.end(), [](const Launch& a, const Launch& b) {
        return a.ncta * a.nt != b.ncta * b.nt ? a.ncta * a.nt > b.ncta * b.nt : a.nt > b.nt;
    });
    {   // ... except that the class with the most work goes last: it is launched machine-wide and absorbs the SMs
        // the others release (the launches are chained, see below)
        size_t mi = 0;
        for (size_t i = 1; i < launches.size(); ++i) if (launches[i].work > launches[mi].work) mi = i;
        if (!launches.empty()) std::rotate(launches.begin() + mi, launches.begin() + mi + 1, launches.end());
    }
    std::vector<int> order;
    order.reserve(nj);
    size_t arena_need = 0, scratch_need = 0;
    // SM budget: every launch gets SMs in proportion to its share of the work (traceback bytes ~ cells)
    const double budget = ctx->n_sm * 1.0;
    auto shape_weight = [](int ncta) { return ncta <= 1 ? 1.0 : ncta == 2 ? 1.75 : ncta == 4 ? 2.0 : 2.4; };
    double weighted_work = 0;
    const Launch* major = launches.empty() ? nullptr : &launches[0];
    for (const Launch& L : launches) { weighted_work += L.work * shape_weight(L.ncta); if (L.work > major->work) major = &L; }
    for (Launch& L : launches) {
        std::sort(L.ids.begin(), L.ids.end(), [&](int a, int b) {
            return tb_bytes[a] != tb_bytes[b] ? tb_bytes[a] > tb_bytes[b] : a < b;
        });
        L.first = (int)order.size(); L.count = (int)L.ids.size();
        L.stride = (tb_bytes[L.ids[0]] + 255) & ~(size_t)255;
        L.max_slots = 0;
        int mq = 0, mt = 0;
        for (int k : L.ids) {
            mq = std::max(mq, jobs[k].qlen); mt = std::max(mt, jobs[k].tlen); order.push_back(k);
            if (L.ncta == 0) L.max_slots = std::max(L.max_slots, ksw_wide_slots(jobs[k].tlen, jobs[k].w));
        }
        L.smem = ksw_smem_bytes(mq, mt);
        if (L.smem > 200 * 1024) { ctx->err = "sequence too long for the shared-memory staging buffers"; return LESVGPU_E_ARG; }
        const int sms_per_job = L.ncta > 0 ? L.ncta : 1;
        const int per_sm = (L.ncta == 1 || L.nt <= 320) ? ksw_ctas_per_sm(L.nt) : 1;
        // `fair` = the SMs the class needs to finish together with the others: its share of the work, weighted by
        // what an SM achieves in that shape (measured: 2-CTA clusters run at ~0.57x the single-CTA rate per SM,
        // full-band 8-CTA clusters at ~0.45x because an anti-diagonal of the whole matrix is a triangle).
        // The class with the most work is additionally launched wide enough to fill the machine on its own
        // (persistent CTAs pulling from the class queue), the others 25 % above fair: when a class runs dry its SMs
        // go to the pending CTAs, so no SM idles while alignments are queued.  The arena fit below takes memory
        // away from the surplus first, so under memory pressure this degrades to the weighted static split.
        const double share = budget * L.work * shape_weight(L.ncta) / weighted_work;
        int cap = 1 << 20;
        if (L.ncta > 1) {
            cap = ksw_max_clusters(L.nt, L.ncta, L.smem);
            if (cap <= 0) { ctx->err = "cluster shape not launchable on this device"; return LESVGPU_E_CUDA; }
        }
        const int whole = std::max(1, std::min(std::min(ctx->n_sm * per_sm / sms_per_job, L.count), cap));
        L.fair = std::max(1, std::min((int)(share * per_sm / sms_per_job + 0.5), whole));
        int n = &L == major ? whole : std::min(whole, (5 * L.fair + 3) / 4);
        if (L.ncta == 1 && L.nt <= kGroupTierMax) n = (n + 32 / L.nt - 1) / (32 / L.nt) * (32 / L.nt);   // whole warps of lane groups
        L.n_clusters = n;
    }
    // fit the arenas.  Cluster i of a launch starts with job i of its (size-sorted) queue and only
    // ever pulls smaller ones, so its arena is exactly job i's traceback; the wide fallback keeps
    // a uniform stride.  Shrink the launch with the largest footprint until everything fits.
    auto footprint = [&](const Launch& L) {
        if (L.ncta == 0) return L.stride * (size_t)L.n_clusters;
        size_t b = 0;
        for (int i = 0; i < L.n_clusters && i < L.count; ++i) b += (tb_bytes[L.ids[i]] + 255) & ~(size_t)255;
        return b;
    };
    const size_t arena_budget = arena_limit;             // what the arena may be right now, strips included
    if (strip_arena > arena_limit) { ctx->err = "traceback arena too small for the full-matrix alignments"; return LESVGPU_E_NOMEM; }
    arena_limit -= strip_arena;                          // the strips' per-alignment arenas sit behind the classes' ones
    for (;;) {
        arena_need = 0;
        for (Launch& L : launches) { L.arena_off = arena_need; L.arena_bytes = footprint(L); arena_need += L.arena_bytes; }
        if (arena_need <= arena_limit) break;
        // Take memory from the surplus first -- that of the smaller classes before that of the class with the most work:
        // their SMs go to its pending CTAs when they run dry, but only as many of those can ever run as have an arena --
        // then from below the fair share.
        Launch* big = nullptr;
        // (a small class keeps one CTA above its share: with a handful of CTAs and alignments of 150 ms each, the exact
        // share ends a whole alignment later than the rest)
        for (Launch& L : launches)
            if (&L != major && L.n_clusters > std::max(L.fair + 1, (11 * L.fair + 9) / 10) && (!big || L.arena_bytes > big->arena_bytes)) big = &L;
        if (!big)
            for (Launch& L : launches)
                if (L.n_clusters > L.fair && (!big || L.arena_bytes > big->arena_bytes)) big = &L;
        if (!big)
            for (Launch& L : launches)
                if (L.n_clusters > ((L.ncta == 1 && L.nt <= kGroupTierMax) ? 32 / L.nt : 1) && (!big || L.arena_bytes > big->arena_bytes)) big = &L;
        if (!big) { ctx->err = "traceback arena too small for this batch"; return LESVGPU_E_NOMEM; }
        big->n_clusters -= (big->ncta == 1 && big->nt <= kGroupTierMax) ? 32 / big->nt : 1;
        if (big->n_clusters < 1) { ctx->err = "traceback arena too small for this batch"; return LESVGPU_E_NOMEM; }
    }
    std::vector<size_t> arena_offs;                 // per launch: n_clusters offsets relative to its arena_off
    std::vector<KswLaunchSlot> counters(launches.size());
    std::vector<size_t> offs_first(launches.size());
    for (size_t li = 0; li < launches.size(); ++li) {
        const Launch& L = launches[li];
        offs_first[li] = arena_offs.size();
        size_t b = 0;
        for (int i = 0; i < L.n_clusters; ++i) {
            arena_offs.push_back(b);
            b += L.ncta == 0 ? L.stride : i < L.count ? ((tb_bytes[L.ids[i]] + 255) & ~(size_t)255) : 0;   // lane groups beyond the queue: no job
        }
        arena_offs.push_back(b);                         // end of the last arena (ksw_dp_overlap_kernel needs the sizes)
        counters[li] = KswLaunchSlot{L.ncta == 0 ? 0 : L.n_clusters, 0, ~0ull, 0ull};       // clusters start with jobs 0..n-1
    }
    for (const Launch& L : launches)
        if (L.ncta == 0) scratch_need = std::max(scratch_need, ksw_wide_scratch_words(L.max_slots) * (size_t)L.n_clusters);

    // ---- column strips of the full-matrix alignments -------------------------------------------------------------
    std::vector<KswStripJob> sjobs;                      // all groups, one after the other
    struct StripGroup { int first, count, n_ctas; };
    std::vector<StripGroup> sgroups;
    size_t edge_words = 0;
    int prog_ints = 0, strip_smem_q = 0, strip_smem_t = 0;
    if (!strip_ids.empty()) {
        struct Per { int n_strips, prog_off; size_t arena_off, edge_off; };
        for (const auto& range : strip_groups) {
            const size_t x0 = range.first, x1 = range.second;
            size_t aoff = arena_need;
            std::vector<Per> per;
            int max_strips = 0;
            for (size_t x = x0; x < x1; ++x) {
                const size_t tb = (tb_bytes[strip_ids[x]] + 255) & ~(size_t)255;
                const KswJob& J = jobs[strip_ids[x]];
                const int max_b = ((J.tlen - 1) >> 4) + 1;
                Per p;
                p.n_strips = (max_b + 1 + kStripBlocks - 1) / kStripBlocks;
                p.prog_off = prog_ints; prog_ints += p.n_strips + 1;
                p.arena_off = aoff; aoff += tb;
                p.edge_off = edge_words; edge_words += (size_t)p.n_strips * (size_t)(J.qlen + J.tlen - 1);
                max_strips = std::max(max_strips, p.n_strips);
                strip_smem_q = std::max(strip_smem_q, J.qlen); strip_smem_t = std::max(strip_smem_t, J.tlen);
                per.push_back(p);
            }
            StripGroup g{(int)sjobs.size(), 0, 0};
            for (int st = 0; st < max_strips; ++st)        // (strip, alignment) order: a strip's predecessor is always handed out first
                for (size_t x = x0; x < x1; ++x)
                    if (st < per[x - x0].n_strips) {
                        const Per& p = per[x - x0];
                        sjobs.push_back(KswStripJob{strip_ids[x], st, p.n_strips, p.prog_off, p.arena_off, p.edge_off});
                    }
            g.count = (int)sjobs.size() - g.first;
            g.n_ctas = (int)std::min<size_t>((size_t)g.count, (size_t)ctx->n_sm * kStripCtasPerSm);
            sgroups.push_back(g);
        }
        arena_need += strip_arena;
        CK(ctx->d_sjobs.reserve(sjobs.size()));
        CK(ctx->d_edges.reserve(edge_words));
        CK(ctx->d_sprog.reserve((size_t)prog_ints));
    }
    for (const StripGroup& g : sgroups) counters.push_back(KswLaunchSlot{g.n_ctas, 0, ~0ull, 0ull});   // work counters of the strip launches (slots launches.size() + group)

    CK(ctx->d_raw.reserve(raw_bytes + 16));
    CK(ctx->d_pool.reserve(pool_words));
    CK(ctx->d_cigar.reserve(cigar_words + 1));
    CK(ctx->d_jobs.reserve(nj));
    CK(ctx->d_outs.reserve(nj));
    CK(ctx->d_runs.reserve((size_t)nj * kSigCap));
    CK(ctx->d_order.reserve(nj));
    CK(ctx->d_counter.reserve((launches.size() + sgroups.size() + 1) * kLaunchSlotInts));
    CK(ctx->d_aoff.reserve(arena_offs.size()));
    // A batch that needs more than half of the budget is memory-bound and the next one will want a little more or less:
    // take the whole budget at once instead of freeing and allocating 100+ GB whenever a batch exceeds the last maximum
    // (cudaMalloc of that size costs 0.1 s and showed up as jitter of the end-to-end time).
    if (arena_need > ctx->d_arena.cap)
        CK(ctx->d_arena.reserve_exact(arena_need > arena_budget / 2 ? std::max(arena_need, arena_budget) : arena_need));
    CK(ctx->d_scratch.reserve(scratch_need));
    CK(ctx->h_outs.reserve(nj));
    if (want_cigar) CK(ctx->h_cigar.reserve(cigar_words));
    if (want_runs) CK(ctx->h_runs.reserve((size_t)nj * kSigCap));

    cudaStream_t s = ctx->stream;
    const auto t_planned = std::chrono::steady_clock::now();
    CK(cudaEventRecord(ctx->ev[0], s));
    CK(cudaMemcpyAsync(ctx->d_raw.p, ctx->h_raw.p, raw_bytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_jobs.p, jobs.data(), sizeof(KswJob) * nj, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_order.p, order.data(), sizeof(int) * nj, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_counter.p, counters.data(), sizeof(KswLaunchSlot) * counters.size(), cudaMemcpyHostToDevice, s));
    if (!sjobs.empty()) {
        CK(cudaMemcpyAsync(ctx->d_sjobs.p, sjobs.data(), sizeof(KswStripJob) * sjobs.size(), cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(ctx->d_sprog.p, 0, sizeof(int) * (size_t)prog_ints, s));
    }
    CK(cudaMemcpyAsync(ctx->d_aoff.p, arena_offs.data(), sizeof(size_t) * arena_offs.size(), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->d_outs.p, 0, sizeof(KswOut) * nj, s));
    CK(cudaEventRecord(ctx->ev[1], s));
    launch_pack(ctx->d_jobs.p, nj, ctx->d_raw.p, ctx->d_subject.p, ctx->d_pool.p, s);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[2], s));
    // the launch classes run concurrently: one side stream each, joined back into the main stream
    // All DP launches of the batch go into ONE side stream, largest shape first, each chained to its predecessor
    // by programmatic stream serialization: a class starts as soon as every CTA of the class before it is resident.
    // So the cluster classes get their SMs first, the class with the most work (launched last, machine-wide) takes
    // what is left and inherits the SMs of the others as they run dry -- independent of how the hardware would
    // order competing streams (two processes on one box were enough to turn that order around).
    static const bool dbg = getenv("LESV_KSW_DEBUG") != nullptr;
    {
        cudaStream_t ls = ctx->side[0];
        CK(cudaStreamWaitEvent(ls, ctx->ev[2], 0));
        bool chained = false, wide_used = false;
        for (size_t li = 0; li < launches.size(); ++li) {
            const Launch& L = launches[li];
            if (L.ncta > 0) {
                CK(launch_ksw(L.nt, L.ncta, L.n_clusters, L.smem, ctx->d_jobs.p, ctx->d_order.p + L.first, L.count,
                              ctx->d_pool.p, ctx->d_arena.p + L.arena_off, ctx->d_aoff.p + offs_first[li], ctx->d_cigar.p,
                              ctx->d_runs.p, ctx->d_outs.p, ctx->d_counter.p + li * kLaunchSlotInts, ls, chained));
                chained = true;
            } else {
                cudaStream_t ws = ctx->side[1];
                if (!wide_used) CK(cudaStreamWaitEvent(ws, ctx->ev[2], 0));
                wide_used = true;
                launch_ksw_wide(L.n_clusters, L.smem, ctx->d_jobs.p, ctx->d_order.p + L.first, L.count, ctx->d_pool.p,
                                ctx->d_arena.p + L.arena_off, L.stride, ctx->d_scratch.p, ksw_wide_scratch_words(L.max_slots),
                                L.max_slots, ctx->d_cigar.p, ctx->d_runs.p, ctx->d_outs.p, ctx->d_counter.p + li * kLaunchSlotInts, ws);
                CK(cudaGetLastError());
            }
        }
        if (!sjobs.empty()) {                           // one launch per group, in order, on their own stream
            cudaStream_t ss = ctx->side[2];
            CK(cudaStreamWaitEvent(ss, ctx->ev[2], 0));
            for (size_t gi = 0; gi < sgroups.size(); ++gi) {
                const StripGroup& g = sgroups[gi];
                CK(launch_ksw_strips(g.n_ctas, ksw_smem_bytes(strip_smem_q, strip_smem_t), ctx->d_jobs.p, ctx->d_sjobs.p + g.first, g.count,
                                     ctx->d_pool.p, ctx->d_arena.p, ctx->d_edges.p, ctx->d_sprog.p, ctx->d_cigar.p, ctx->d_runs.p,
                                     ctx->d_outs.p, ctx->d_counter.p + (launches.size() + gi) * kLaunchSlotInts, ss));
            }
            CK(cudaEventRecord(ctx->side_done[2], ss));
            CK(cudaStreamWaitEvent(s, ctx->side_done[2], 0));
        }
        CK(cudaEventRecord(ctx->side_done[0], ls));
        CK(cudaStreamWaitEvent(s, ctx->side_done[0], 0));
        if (wide_used) {
            CK(cudaEventRecord(ctx->side_done[1], ctx->side[1]));
            CK(cudaStreamWaitEvent(s, ctx->side_done[1], 0));
        }
    }
    CK(cudaEventRecord(ctx->ev[3], s));
    size_t d2h = sizeof(KswOut) * nj;
    CK(cudaMemcpyAsync(ctx->h_outs.p, ctx->d_outs.p, sizeof(KswOut) * nj, cudaMemcpyDeviceToHost, s));
    if (want_cigar && cigar_words) {
        CK(cudaMemcpyAsync(ctx->h_cigar.p, ctx->d_cigar.p, sizeof(uint32_t) * cigar_words, cudaMemcpyDeviceToHost, s));
        d2h += sizeof(uint32_t) * cigar_words;
    }
    if (want_runs) {
        CK(cudaMemcpyAsync(ctx->h_runs.p, ctx->d_runs.p, sizeof(KswGapRun) * (size_t)nj * kSigCap, cudaMemcpyDeviceToHost, s));
        d2h += sizeof(KswGapRun) * (size_t)nj * kSigCap;
    }
    CK(cudaMemcpyAsync(counters.data(), ctx->d_counter.p, sizeof(KswLaunchSlot) * counters.size(), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(ctx->ev[4], s));
    CK(cudaStreamSynchronize(s));

    if (dbg) {
        const auto t_done = std::chrono::steady_clock::now();
        fprintf(stderr, "[ksw debug] arena %.1f GB of limit %.1f GB\n", arena_need / 1e9, arena_limit / 1e9);
        fprintf(stderr, "[ksw debug] host: plan+alloc %.1f ms, launch+wait %.1f ms\n",
                std::chrono::duration<double, std::milli>(t_planned - t_enter).count(),
                std::chrono::duration<double, std::milli>(t_done - t_planned).count());
        float total = 0;
        cudaEventElapsedTime(&total, ctx->ev[2], ctx->ev[3]);
        for (size_t li = 0; li < launches.size(); ++li) {
            const Launch& L = launches[li];
            double cells = 0;
            for (int k : L.ids) cells += (double)cells_fast(jobs[k].qlen, jobs[k].tlen, jobs[k].w);
            fprintf(stderr, "[ksw debug] launch %zu: ncta=%d tier=%d jobs=%d ctas/clusters=%d (fair %d) cells=%.1fG\n",
                    li, L.ncta, L.nt, L.count, L.n_clusters, L.fair, cells / 1e9);
        }
        if (!sjobs.empty()) fprintf(stderr, "[ksw debug] strips: %zu alignments, %zu strips in %zu group(s)\n", strip_ids.size(), sjobs.size(), sgroups.size());
        fprintf(stderr, "[ksw debug] all launches: %.1f ms\n", total);
    }
    ctx->classes.clear();
    for (size_t li = 0; li < launches.size() + sgroups.size(); ++li) {
        lesvgpu_class_timing c = {};
        if (li < launches.size()) {
            const Launch& L = launches[li];
            c.ctas_per_alignment = L.ncta; c.threads = L.nt; c.n_alignments = L.count; c.n_ctas = L.n_clusters;
            for (int k : L.ids) c.cells += (uint64_t)cells_fast(jobs[k].qlen, jobs[k].tlen, jobs[k].w);
        } else {
            const StripGroup& g = sgroups[li - launches.size()];
            c.ctas_per_alignment = -1; c.threads = kStripBlocks; c.n_ctas = g.n_ctas;
            for (int x = g.first; x < g.first + g.count; ++x)
                if (sjobs[x].s == 0) {
                    ++c.n_alignments;
                    const KswJob& J = jobs[sjobs[x].jid];
                    c.cells += (uint64_t)cells_fast(J.qlen, J.tlen, J.w);
                }
        }
        unsigned long long t_first = ~0ull;
        for (const KswLaunchSlot& k : counters) t_first = std::min(t_first, k.t_start);
        c.done_ms = counters[li].t_end > t_first ? (float)((counters[li].t_end - t_first) * 1e-6) : 0.f;
        ctx->classes.push_back(c);
    }
    lesvgpu_timing& T = ctx->timing;
    cudaEventElapsedTime(&T.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&T.pack_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&T.dp_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&T.d2h_ms, ctx->ev[3], ctx->ev[4]);
    T.scan_ms = 0.f;                                  // the scan runs inside the traceback walk
    T.dp_launches = (int)launches.size() + (int)sgroups.size();
    T.kernel_launches = 1 + T.dp_launches;
    g_lesv_launches += (unsigned long long)T.kernel_launches;
    T.h2d_bytes = raw_bytes + sizeof(KswJob) * nj + sizeof(int) * nj;
    T.d2h_bytes = d2h;
    for (const KswJob& J : jobs) T.dp_cells += (uint64_t)cells_fast(J.qlen, J.tlen, J.w);
    return LESVGPU_OK;
}

extern "C" int lesvgpu_ksw_batch(lesvgpu_ctx* ctx, const lesvgpu_ksw_pair* pairs, int n,
                                 lesvgpu_ksw_result* results, const uint32_t** cigar_pool)
{
    if (!ctx || n < 0 || (n && (!pairs || !results))) return LESVGPU_E_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LESVGPU_E_CUDA;
    ctx->timing = lesvgpu_timing{};
    if (cigar_pool) *cigar_pool = nullptr;
    std::vector<KswJob> jobs;
    std::vector<int> job_pair;
    jobs.reserve(n); job_pair.reserve(n);
    size_t raw_bytes = 0;
    for (int i = 0; i < n; ++i) {
        lesvgpu_ksw_result& R = results[i];
        R.ok = 0; R.n_cigar = 0; R.cigar_off = 0; R.aln_size = 0; R.n_match = 0;
        const lesvgpu_ksw_pair& P = pairs[i];
        if (P.qlen <= 0 || P.tlen <= 0) continue;              // ksw2_extd2_sse.c:68 -> n_cigar 0
        if (!P.query || !P.target) { ctx->err = "null sequence pointer"; return LESVGPU_E_ARG; }
        const int w = effective_band(P.qlen, P.tlen, P.max_distance);
        if (!band_reaches_end(P.qlen, P.tlen, w)) continue;    // partial CIGAR -> the wrapper returns 0
        KswJob J;
        memset(&J, 0, sizeof J);
        J.qlen = P.qlen; J.tlen = P.tlen; J.w = w;
        J.seq_off_q = raw_bytes; raw_bytes += (size_t)P.qlen;
        J.seq_off_t = raw_bytes; raw_bytes += (size_t)P.tlen;
        jobs.push_back(J); job_pair.push_back(i);
    }
    const int nj = (int)jobs.size();
    if (nj == 0) return LESVGPU_OK;
    CK(ctx->h_raw.reserve(raw_bytes));
    for (int k = 0; k < nj; ++k) {
        const lesvgpu_ksw_pair& P = pairs[job_pair[k]];
        memcpy(ctx->h_raw.p + jobs[k].seq_off_q, P.query, (size_t)P.qlen);
        memcpy(ctx->h_raw.p + jobs[k].seq_off_t, P.target, (size_t)P.tlen);
    }
    const int rc = run_jobs(ctx, jobs, raw_bytes, cigar_pool != nullptr, false);
    if (rc != LESVGPU_OK) return rc;
    for (int k = 0; k < nj; ++k) {
        lesvgpu_ksw_result& R = results[job_pair[k]];
        const KswOut& O = ctx->h_outs.p[k];
        R.n_cigar = O.n_cigar; R.cigar_off = jobs[k].cigar_off;
        R.aln_size = O.aln_size; R.n_match = O.n_match;
        R.ok = O.n_cigar > 0 && O.status == 1;
    }
    if (cigar_pool) *cigar_pool = ctx->h_cigar.p;
    return LESVGPU_OK;
}

extern "C" int lesvgpu_subject_load(lesvgpu_ctx* ctx, const uint8_t* subject_pac, size_t pac_bytes)
{
    if (!ctx || !subject_pac || !pac_bytes) return LESVGPU_E_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LESVGPU_E_CUDA;
    CK(ctx->d_subject.reserve(pac_bytes + 16));
    CK(cudaMemcpyAsync(ctx->d_subject.p, subject_pac, pac_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->subject_bytes = pac_bytes;
    return LESVGPU_OK;
}

// Body of lesvgpu_svsig_batch / lesvgpu_svsig_batch_contigs.  one_contig: every read lies on subject_seqs[0]
// whatever its subject_id says (the single-contig entry point takes the contig's offset and size, not a table).
static int svsig_core(lesvgpu_ctx* ctx, const uint8_t* query_pac, const size_t* query_read_offset,
                      const lesvgpu_seq* subject_seqs, int n_subject, bool one_contig,
                      const lesvgpu_svread* reads, int n, int min_indel_size, const char* trf_mask,
                      lesvgpu_svsig_read* per_read, const lesvgpu_svsig** sigs, int* n_sigs)
{
    if (!ctx || n < 0 || (n && (!reads || !per_read || !query_pac || !query_read_offset)) || !subject_seqs || n_subject <= 0) return LESVGPU_E_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LESVGPU_E_CUDA;
    if (!ctx->subject_bytes) { ctx->err = "lesvgpu_subject_load has not been called"; return LESVGPU_E_ARG; }
    ctx->timing = lesvgpu_timing{};
    ctx->sigs.clear();
    ctx->sig_subject.clear();
    if (sigs) *sigs = nullptr;
    if (n_sigs) *n_sigs = 0;
    for (int i = 0; i < n && !one_contig; ++i)
        if (reads[i].subject_id < 0 || reads[i].subject_id >= n_subject) {
            ctx->err = "SvRead names a contig the subject table does not have"; return LESVGPU_E_ARG;
        }
    auto contig_of = [&](const lesvgpu_svread& S) -> const lesvgpu_seq& { return subject_seqs[one_contig ? 0 : S.subject_id]; };
    std::vector<KswJob> jobs;
    std::vector<int> job_read;
    jobs.reserve(n); job_read.reserve(n);
    // pass 1: size the raw upload (the .pac bytes covering each read range)
    size_t raw_bytes = 0;
    std::vector<size_t> src_byte(n), dst_byte(n), nbytes(n);
    for (int i = 0; i < n; ++i) {
        lesvgpu_svsig_read& R = per_read[i];
        memset(&R, 0, sizeof R);
        const lesvgpu_svread& S = reads[i];
        const size_t subject_seq_offset = contig_of(S).seq_offset;
        const int subject_size = (int)contig_of(S).seq_size;
        const int qn = S.qend - S.qoff, sn = S.send - S.soff;
        if (qn <= 0 || sn <= 0 || S.qend > S.qsize || S.send > subject_size || S.qoff < 0 || S.soff < 0 || S.query_id < 0) {
            ctx->err = "SvRead range outside its sequence"; return LESVGPU_E_ARG;
        }
        if (!query_read_offset[S.query_id]) { ctx->err = "read not loaded (raw_read_offset_array == 0)"; return LESVGPU_E_ARG; }
        // discover_func :272-316: band (int)(dist*1.2) [0 -> 10 % rule], then max(qlen, slen)
        int w = effective_band(qn, sn, (int)(S.dist * 1.2));
        if (!band_reaches_end(qn, sn, w)) w = qn > sn ? qn : sn;
        R.band = w;
        const size_t res0 = query_read_offset[S.query_id] - 1 + (size_t)S.qoff;   // forward-strand range
        src_byte[i] = res0 >> 2;
        nbytes[i] = ((res0 + (size_t)qn - 1) >> 2) - src_byte[i] + 1;
        dst_byte[i] = raw_bytes;
        raw_bytes += nbytes[i];
        KswJob J;
        memset(&J, 0, sizeof J);
        J.qlen = qn; J.tlen = sn; J.w = w;
        J.src = 1; J.qdir = S.qdir;
        J.seq_off_q = dst_byte[i] * 4 + (res0 & 3);
        J.seq_off_t = subject_seq_offset + (size_t)S.soff;
        if (((J.seq_off_t + (size_t)sn + 3) >> 2) > ctx->subject_bytes) { ctx->err = "subject range outside the loaded .pac"; return LESVGPU_E_ARG; }
        J.min_indel = min_indel_size;
        jobs.push_back(J); job_read.push_back(i);
    }
    const int nj = (int)jobs.size();
    if (nj == 0) return LESVGPU_OK;
    CK(ctx->h_raw.reserve(raw_bytes));
    for (int i = 0; i < n; ++i) memcpy(ctx->h_raw.p + dst_byte[i], query_pac + src_byte[i], nbytes[i]);
    const int rc = run_jobs(ctx, jobs, raw_bytes, false, true);
    if (rc != LESVGPU_OK) return rc;
    // The walk counts every SV-sized gap run but keeps kSigCap records per alignment.  find_sv_signature has no such
    // limit (find_sv_signature.c:126-220), so alignments with more runs -- a small min_indel_size, a read over a
    // tandem-repeat region -- are run once more with a CIGAR, which is scanned here for the same records.
    const KswOut* outs = ctx->h_outs.p;
    const KswGapRun* runs = ctx->h_runs.p;
    std::vector<KswOut> outs_keep;
    std::vector<KswGapRun> runs_keep;
    std::vector<int> over;
    std::vector<std::vector<KswGapRun>> over_runs;       // left to right
    for (int k = 0; k < nj; ++k)
        if (outs[k].status == 1 && outs[k].n_cigar > 0 && outs[k].trunc_ok && outs[k].n_runs > kSigCap) over.push_back(k);
    if (!over.empty()) {
        outs_keep.assign(outs, outs + nj);
        runs_keep.assign(runs, runs + (size_t)nj * kSigCap);
        outs = outs_keep.data(); runs = runs_keep.data();
        const lesvgpu_timing first = ctx->timing;
        std::vector<KswJob> again;
        for (int k : over) { KswJob J = jobs[k]; J.min_indel = 0; again.push_back(J); }
        const int rc2 = run_jobs(ctx, again, raw_bytes, true, false);
        if (rc2 != LESVGPU_OK) return rc2;
        lesvgpu_timing& T = ctx->timing;
        T.h2d_ms += first.h2d_ms; T.pack_ms += first.pack_ms; T.dp_ms += first.dp_ms; T.d2h_ms += first.d2h_ms;
        T.dp_launches += first.dp_launches; T.kernel_launches += first.kernel_launches;
        T.h2d_bytes += first.h2d_bytes; T.d2h_bytes += first.d2h_bytes;      // dp_cells accumulates by itself
        over_runs.resize(over.size());
        for (size_t x = 0; x < over.size(); ++x) {
            const KswOut& O = outs[over[x]];
            const KswOut& O2 = ctx->h_outs.p[x];
            if (O2.status != 1 || O2.n_cigar <= 0) { ctx->err = "re-alignment for the signature scan failed"; return LESVGPU_E_CUDA; }
            const uint32_t* cg = ctx->h_cigar.p + again[x].cigar_off;
            int si = 0, qi = 0;                              // target / read bases consumed so far
            for (int c = 0; c < O2.n_cigar; ++c) {
                const int op = (int)(cg[c] & 15u), len = (int)(cg[c] >> 4);
                if (op == 0) { si += len; qi += len; continue; }
                const int ds = op == 2 ? len : 0, dq = op == 1 ? len : 0;
                if (len >= min_indel_size && si >= O.sb && qi >= O.qb && si + ds <= O.se && qi + dq <= O.qe) {
                    KswGapRun g;
                    if (op == 2) { g.type = 0; g.qfrom = qi; g.qto = qi + 1; g.sfrom = si; g.sto = si + len; }
                    else { g.type = 6; g.qfrom = qi; g.qto = qi + len; g.sfrom = si; g.sto = si + 1; }
                    over_runs[x].push_back(g);
                }
                si += ds; qi += dq;
            }
            if ((int)over_runs[x].size() != O.n_runs) { ctx->err = "signature scan of the CIGAR disagrees with the walk"; return LESVGPU_E_CUDA; }
        }
    }
    size_t next_over = 0;
    for (int k = 0; k < nj; ++k) {
        const int i = job_read[k];
        const lesvgpu_svread& S = reads[i];
        lesvgpu_svsig_read& R = per_read[i];
        const KswOut& O = outs[k];
        const std::vector<KswGapRun>* wide = nullptr;
        if (next_over < over.size() && over[next_over] == k) wide = &over_runs[next_over++];
        R.qb = O.qb; R.qe = O.qe; R.sb = O.sb; R.se = O.se;
        R.eff_mat = O.eff_mat; R.eff_len = O.eff_len;
        if (O.status != 1 || O.n_cigar <= 0) { R.status = 0; continue; }
        if (!O.trunc_ok) { R.status = -1; continue; }
        const size_t subject_seq_offset = contig_of(S).seq_offset;
        const int subject_size = (int)contig_of(S).seq_size;
        const int fqb = O.qb + S.qoff, fqe = O.qe + S.qoff;          // :334-337 (forward qoff even for REV)
        const int fsb = O.sb + S.soff, fse = O.se + S.soff;
        if (trf_mask) {                                              // fall_in_trf_subseq, trf_array.cpp:74-83
            size_t cnt = 0;
            for (size_t x = subject_seq_offset + fsb; x < subject_seq_offset + (size_t)fse; ++x) cnt += trf_mask[x];
            if ((size_t)(fse - fsb) - cnt <= 2000) { R.status = -2; continue; }
        }
        const double eff = O.eff_len == 0 ? 0.0 : 100.0 * O.eff_mat / O.eff_len;   // hbn_traceback_aux.c:93-94
        if (eff < 70.0) { R.status = -3; continue; }
        R.status = 1; R.n_sig = O.n_runs;
        const KswGapRun* g = wide ? wide->data() : runs + (size_t)k * kSigCap;
        for (int y = 0; y < O.n_runs; ++y) {
            const int x = wide ? y : O.n_runs - 1 - y;               // the walk found them right-to-left
            lesvgpu_svsig v;
            v.qid = S.query_id; v.qdir = S.qdir;
            v.qfrom = g[x].qfrom; v.qto = g[x].qto;
            v.fqfrom = fqb; v.fqto = fqe; v.qsize = S.qsize;
            v.sfrom = S.soff + g[x].sfrom; v.sto = S.soff + g[x].sto;
            v.fsfrom = fsb; v.fsto = fse; v.ssize = subject_size;
            v.type = g[x].type;
            ctx->sigs.push_back(v);
            ctx->sig_subject.push_back(one_contig ? 0 : S.subject_id);
        }
    }
    if (sigs) *sigs = ctx->sigs.data();
    if (n_sigs) *n_sigs = (int)ctx->sigs.size();
    return LESVGPU_OK;
}

extern "C" int lesvgpu_svsig_batch(lesvgpu_ctx* ctx,
                                   const uint8_t* query_pac, const size_t* query_read_offset,
                                   size_t subject_seq_offset, int subject_size,
                                   const lesvgpu_svread* reads, int n, int min_indel_size,
                                   const char* trf_mask,
                                   lesvgpu_svsig_read* per_read, const lesvgpu_svsig** sigs, int* n_sigs)
{
    lesvgpu_seq one;
    one.seq_offset = subject_seq_offset; one.seq_size = (uint64_t)(subject_size > 0 ? subject_size : 0);
    return svsig_core(ctx, query_pac, query_read_offset, &one, 1, true, reads, n, min_indel_size, trf_mask, per_read, sigs, n_sigs);
}

extern "C" int lesvgpu_svsig_batch_contigs(lesvgpu_ctx* ctx,
                                           const uint8_t* query_pac, const size_t* query_read_offset,
                                           const lesvgpu_seq* subject_seqs, int n_subject,
                                           const lesvgpu_svread* reads, int n, int min_indel_size,
                                           const char* trf_mask,
                                           lesvgpu_svsig_read* per_read, const lesvgpu_svsig** sigs, const int** sig_subject, int* n_sigs)
{
    if (sig_subject) *sig_subject = nullptr;
    const int rc = svsig_core(ctx, query_pac, query_read_offset, subject_seqs, n_subject, false, reads, n, min_indel_size, trf_mask,
                              per_read, sigs, n_sigs);
    if (rc == LESVGPU_OK && sig_subject && ctx) *sig_subject = ctx->sig_subject.data();
    return rc;
}

extern "C" int lesvgpu_align_refine_batch(lesvgpu_ctx* ctx, const uint8_t* query_pac, const size_t* query_read_offset,
                                          const lesvgpu_refine_job* rj, int n, lesvgpu_refine_result* results)
{
    if (!ctx || n < 0 || (n && (!rj || !results || !query_pac || !query_read_offset))) return LESVGPU_E_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return LESVGPU_E_CUDA;
    if (!ctx->subject_bytes) { ctx->err = "lesvgpu_subject_load has not been called"; return LESVGPU_E_ARG; }
    ctx->timing = lesvgpu_timing{};
    constexpr int kKswMaxBw = 30000;                      // KSW_MAX_BW, algo/ksw2_wrapper.h:10
    std::vector<KswJob> jobs;
    std::vector<int> job_idx;
    std::vector<size_t> src_byte, dst_byte, nbytes;
    size_t raw_bytes = 0;
    for (int i = 0; i < n; ++i) {
        lesvgpu_refine_result& R = results[i];
        memset(&R, 0, sizeof R);
        const lesvgpu_refine_job& A = rj[i];
        const int qn = A.query_to - A.query_from, sn = A.subject_to - A.subject_from;
        if (qn <= 0 || sn <= 0 || A.query_from < 0 || A.query_to > A.query_size || A.subject_from < 0) {
            ctx->err = "refine job range outside its sequence"; return LESVGPU_E_ARG;
        }
        const int longest = qn > sn ? qn : sn;
        int band = A.max_dist < 0 ? longest : A.max_dist;                 // align_subseqs.c:233
        int w = effective_band(qn, sn, band);                             // ksw2_wrapper.c:121
        if (!band_reaches_end(qn, sn, w)) {
            if (band < kKswMaxBw && band < longest) { band = longest; w = effective_band(qn, sn, band); }   // :255
            if (!band_reaches_end(qn, sn, w)) continue;                   // "ksw fail"
        }
        R.band = w;
        // forward-strand range of the read that strand coordinates [from, to) denote
        const int ffrom = A.query_dir ? A.query_size - A.query_to : A.query_from;
        if (!query_read_offset[A.query_id]) { ctx->err = "read not loaded (raw_read_offset_array == 0)"; return LESVGPU_E_ARG; }
        const size_t res0 = query_read_offset[A.query_id] - 1 + (size_t)ffrom;
        src_byte.push_back(res0 >> 2);
        nbytes.push_back(((res0 + (size_t)qn - 1) >> 2) - (res0 >> 2) + 1);
        dst_byte.push_back(raw_bytes);
        KswJob J;
        memset(&J, 0, sizeof J);
        J.qlen = qn; J.tlen = sn; J.w = w;
        J.src = 1; J.qdir = A.query_dir;
        J.seq_off_q = raw_bytes * 4 + (res0 & 3);
        J.seq_off_t = A.subject_seq_offset + (size_t)A.subject_from;
        if (((J.seq_off_t + (size_t)sn + 3) >> 2) > ctx->subject_bytes) { ctx->err = "subject range outside the loaded .pac"; return LESVGPU_E_ARG; }
        raw_bytes += nbytes.back();
        jobs.push_back(J); job_idx.push_back(i);
    }
    const int nj = (int)jobs.size();
    if (nj == 0) return LESVGPU_OK;
    CK(ctx->h_raw.reserve(raw_bytes));
    for (int k = 0; k < nj; ++k) memcpy(ctx->h_raw.p + dst_byte[k], query_pac + src_byte[k], nbytes[k]);
    const int rc = run_jobs(ctx, jobs, raw_bytes, false, false);
    if (rc != LESVGPU_OK) return rc;
    for (int k = 0; k < nj; ++k) {
        const lesvgpu_refine_job& A = rj[job_idx[k]];
        lesvgpu_refine_result& R = results[job_idx[k]];
        const KswOut& O = ctx->h_outs.p[k];
        if (O.status != 1 || O.n_cigar <= 0 || !O.trunc_ok) continue;    // ksw / truncate fail
        const int x = O.qe - O.qb, y = O.se - O.sb;
        if (x < jobs[k].qlen * 0.8 && y < jobs[k].tlen * 0.8) continue;   // :307 "align too short"
        R.ok = 1;
        R.qb = O.qb + A.query_from; R.qe = O.qe + A.query_from;
        R.sb = O.sb + A.subject_from; R.se = O.se + A.subject_from;
        R.mat = O.trim_mat; R.len = O.trim_len; R.eff_mat = O.eff_mat; R.eff_len = O.eff_len;
    }
    return LESVGPU_OK;
}
