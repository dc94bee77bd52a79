// ksw_emul.cpp -- single-threaded CPU run of the K-ksw kernel's per-thread update code.
// TEST VEHICLE ONLY: compiles lesv_b200/csrc/ksw_core.cuh in host mode and steps NT emulated
// threads row by row exactly as ksw_dp_kernel schedules them (block ownership, carries, pre-block
// score tracking), then backtracks over the internal traceback bytes.  tests/test_ksw_emul.py
// compares the CIGARs with the oracle so that kernel logic errors are caught without a GPU.
#include "../../lesv_b200/csrc/ksw_core.cuh"
#include <vector>
#include <cstdlib>
#include <cstring>

using namespace lesvksw;

// pack 16 bases per word, base k at bits 2(k&15)
static void pack_words(const uint8_t* s, int n, int reversed, std::vector<uint32_t>& out, int pad_front, int nwords)
{
    out.assign(nwords + pad_front, 0u);
    for (int k = 0; k < n; ++k) {
        uint32_t c = reversed ? s[n - 1 - k] : s[k];
        out[pad_front + (k >> 4)] |= (c & 3u) << (2 * (k & 15));
    }
}

static int emul_impl(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                     int w, int nt, uint32_t** cigar_out, uint8_t** tb_out, int* tb_row_bytes,
                     int skew_k = 0, unsigned seed = 0);
static int g_eager = 0;     // 1: step with thread_row_step<true> (carry resolved up front, barrier-mode kernel)
extern "C" void lesv_emul_set_eager(int on) { g_eager = on; }
extern "C" void lesv_emul_set_lean(int on) { lesvksw::g_ksw_emul_lean = on != 0; }
static int g_strip = 0;     // > 0: column strips of g_strip blocks, one strip after the other (ksw_dp_strip_kernel)
extern "C" void lesv_emul_set_strip(int blocks) { g_strip = blocks; }

extern "C" int lesv_emul_ksw_extd2(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                   int w, int nt, uint32_t** cigar_out)
{
    return emul_impl(qlen, query, tlen, target, w, nt, cigar_out, nullptr, nullptr);
}
// also returns the traceback matrix translated to the reference byte layout (natural column order,
// flags at bits 3..6, bit 7 clear) so tests can compare it cell by cell with the oracle's
extern "C" int lesv_emul_ksw_extd2_tb(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                      int w, int nt, uint32_t** cigar_out, uint8_t** tb_out, int* tb_row_bytes)
{
    return emul_impl(qlen, query, tlen, target, w, nt, cigar_out, tb_out, tb_row_bytes);
}

// Skewed schedule (kSyncSkew of ksw_dp_kernel): the 32-block group g of the band runs skew_k rows
// behind group g-1, warps exchange column-15 carries through a ring of 4*skew_k slots and meet at a
// CTA barrier only every skew_k ticks.  Inside a barrier interval the warps are run one after the
// other in a random order -- the most adversarial interleaving the barrier allows -- so a slot that
// could be read before it is written, or overwritten before it is read, shows up as a wrong CIGAR.
static int g_group = 32;   // threads per skew group: 32 = warps of one CTA, NT = the CTAs of a cluster
extern "C" int lesv_emul_ksw_extd2_skew(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                        int w, int nt, int skew_k, unsigned seed, uint32_t** cigar_out)
{
    g_group = 32;
    return emul_impl(qlen, query, tlen, target, w, nt, cigar_out, nullptr, nullptr, skew_k, seed);
}
// cluster flavour: `nt` threads in groups of `group` (one CTA each), lock-step inside a group
extern "C" int lesv_emul_ksw_extd2_skew_group(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                              int w, int nt, int group, int skew_k, unsigned seed, uint32_t** cigar_out)
{
    g_group = group;
    const int n = emul_impl(qlen, query, tlen, target, w, nt, cigar_out, nullptr, nullptr, skew_k, seed);
    g_group = 32;
    return n;
}

static int emul_impl(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                     int w, int nt, uint32_t** cigar_out, uint8_t** tb_out, int* tb_row_bytes,
                     int skew_k, unsigned seed)
{
    *cigar_out = nullptr;
    if (qlen <= 0 || tlen <= 0) return 0;
    if (w < 0) w = tlen > qlen ? tlen : qlen;
    // band must not be cut (the host library rejects those jobs before launch)
    for (int r = 0; r < qlen + tlen - 1; ++r) if (row_geom(r, qlen, tlen, w).empty()) return 0;
    const int prow_blocks = tb_row_blocks(qlen, tlen, w);
    const size_t prow = (size_t)prow_blocks * 16;
    const size_t nrow = (size_t)qlen + tlen - 1;
    std::vector<uint8_t> p(nrow * prow + 16, 0xAA);
    std::vector<uint32_t> tp, qrp;
    pack_words(target, tlen, 0, tp, 0, (tlen >> 4) + 4);
    pack_words(query, qlen, 1, qrp, 1, (qlen >> 4) + 6);
    std::vector<ThreadCtx> th(nt);
    for (int t = 0; t < nt; ++t) thread_ctx_reset(th[t], t);
    std::vector<Carry> pub(nt), pub_next(nt);
    for (auto& c : pub) c = carry_init();
    if (skew_k > 0) {
        const int G = g_group;
        if (nt % G) return -3;
        const int K = skew_k, R = 4 * K, NW = nt / G;
        for (int r = 0; r < (int)nrow; ++r) {
            const RowGeom g = row_geom(r, qlen, tlen, w);
            if ((g.wr >> 4) - (g.st0 >> 4) + 1 > nt) return -1;
        }
        const int maxB = ((tlen - 1) >> 4) + 1;
        const int nticks = (int)nrow + K * (maxB / G);
        std::vector<int> rr(nt);
        std::vector<Carry> prev(nt, carry_init()), ring((size_t)NW * R, carry_init());
        for (auto& c : ring) { c.xv = 0xDEADBEEFu; c.x2 = 0xDEADBEEFu; }
        for (int t = 0; t < nt; ++t) rr[t] = -K * (th[t].cand / G);
        std::vector<int> perm(NW);
        for (int s0 = 0; s0 < nticks; s0 += K) {
            for (int k = 0; k < NW; ++k) perm[k] = k;
            for (int k = NW - 1; k > 0; --k) { seed = seed * 1664525u + 1013904223u; std::swap(perm[k], perm[(seed >> 8) % (k + 1)]); }
            for (int wp : perm)
                for (int s = s0; s < s0 + K && s < nticks; ++s) {
                    std::vector<Carry> cin(G);
                    for (int l = 1; l < G; ++l) cin[l] = prev[wp * G + l - 1];
                    for (int l = 0; l < G; ++l) {
                        const int tid = wp * G + l;
                        ThreadCtx& c = th[tid];
                        int& r = rr[tid];
                        if (!row_is_fast(c, r))
                            while (c.cand <= maxB && (r >= (int)nrow || (r >= 0 && c.cand < (row_geom(r, qlen, tlen, w).st0 >> 4)))) {
                                c.cand += nt; r = s - K * (c.cand / G);
                            }
                        if (r >= 0 && r < (int)nrow && c.cand <= maxB) {
                            Carry cout; uint32_t tb[4];
                            const int rcur = r;
                            if (thread_row_step(c, nt, r, qlen, tlen, w, prow_blocks, tp.data(), qrp.data(),
                                                [&]() { return l ? cin[l] : ring[(size_t)wp * R + ((rcur - 1) & (R - 1))]; }, cout, tb)) {
                                if (c.tb_off < 0 || c.tb_off >= (int)prow) return -2;
                                memcpy(&p[(size_t)r * prow + (size_t)c.tb_off], tb, 16);
                                prev[tid] = cout;
                                if (l == G - 1) ring[(size_t)((wp + 1) % NW) * R + (r & (R - 1))] = cout;
                            }
                        }
                        ++r;
                    }
                }
        }
    } else if (g_strip > 0) {
        // ksw_dp_strip_kernel: every thread keeps one block for the whole pass (nt = 2^24 in thread_row_step), strip s
        // takes the column-15 values of strip s-1 from a per-row edge buffer; row range and carry wiring as in the kernel
        const int SW = g_strip;
        const int max_b = ((tlen - 1) >> 4) + 1;
        const int n_strips = (max_b + 1 + SW - 1) / SW;
        const int nrow_all = (int)nrow;
        std::vector<Carry> edge_prev(nrow, carry_init()), edge_cur(nrow, carry_init());
        for (int st = 0; st < n_strips; ++st) {
            const int B0 = st * SW;
            const int c0 = 16 * B0 - 16, c1e = 16 * (B0 + SW - 1) + 15;
            int r_lo = c0 > 2 * c0 - w ? c0 : 2 * c0 - w;
            r_lo = r_lo < 0 ? 0 : r_lo;
            long long hi = (long long)c1e + qlen - 1, hi2 = 2ll * c1e + w;
            hi = hi < hi2 ? hi : hi2;
            const int r_hi = hi < nrow_all - 1 ? (int)hi : nrow_all - 1;
            std::vector<ThreadCtx> ts(SW);
            std::vector<Carry> prev(SW, carry_init()), prev_next(SW, carry_init());
            for (int t = 0; t < SW; ++t) thread_ctx_reset(ts[t], B0 + t);
            for (auto& e : edge_cur) { e.xv = 0xDEADBEEFu; e.x2 = 0xDEADBEEFu; }
            for (int r = r_lo; r <= r_hi; ++r) {
                for (int tid = 0; tid < SW; ++tid) {
                    Carry cin;
                    if (tid == 0) { if (st > 0 && r > 0) cin = edge_prev[r - 1]; else { cin.xv = 0; cin.x2 = 0; } }
                    else cin = prev[tid - 1];
                    Carry cout; uint32_t tb[4];
                    prev_next[tid] = prev[tid];
                    if (thread_row_step<true>(ts[tid], 1 << 24, r, qlen, tlen, w, prow_blocks, tp.data(), qrp.data(),
                                              [&]() { return cin; }, cout, tb)) {
                        if (ts[tid].tb_off < 0 || ts[tid].tb_off >= (int)prow) return -2;
                        memcpy(&p[(size_t)r * prow + (size_t)ts[tid].tb_off], tb, 16);
                        prev_next[tid] = cout;
                        if (tid == SW - 1) edge_cur[r] = cout;
                    }
                }
                prev.swap(prev_next);
            }
            edge_prev.swap(edge_cur);
        }
    } else
    for (int r = 0; r < (int)nrow; ++r) {
        const RowGeom g = row_geom(r, qlen, tlen, w);
        if ((g.wr >> 4) - (g.st0 >> 4) + 1 > nt) return -1;      // window does not fit NT threads
        for (int tid = 0; tid < nt; ++tid) {
            Carry cout; uint32_t tb[4];
            const Carry cin = pub[(tid + nt - 1) % nt];
            const bool did = g_eager
                ? thread_row_step<true>(th[tid], nt, r, qlen, tlen, w, prow_blocks, tp.data(), qrp.data(), [&]() { return cin; }, cout, tb)
                : thread_row_step<false>(th[tid], nt, r, qlen, tlen, w, prow_blocks, tp.data(), qrp.data(), [&]() { return cin; }, cout, tb);
            if (did) {
                if (th[tid].tb_off < 0 || th[tid].tb_off >= (int)prow) return -2;
                memcpy(&p[(size_t)r * prow + (size_t)th[tid].tb_off], tb, 16);
                pub_next[tid] = cout;
            } else pub_next[tid] = carry_init();
        }
        pub.swap(pub_next);
    }
    // backtrack (ksw2.h:121-153) over the internal byte format
    std::vector<uint32_t> cig;
    auto push = [&](uint32_t op, int len) {
        if (cig.empty() || (cig.back() & 0xf) != op) cig.push_back((uint32_t)len << 4 | op);
        else cig.back() += (uint32_t)len << 4;
    };
    int i = tlen - 1, j = qlen - 1, state = 0;
    while (i >= 0 && j >= 0) {
        const int r = i + j;
        const RowGeom g = row_geom(r, qlen, tlen, w);
        const int off = (g.st0 >> 4) << 4, off_end = ((g.en0 >> 4) << 4) + 15;
        int force = -1;
        if (i < off) force = 2;
        if (i > off_end) force = 1;
        uint32_t byte = 0;
        if (force < 0) byte = p[(size_t)r * prow + (size_t)tb_block_slot(i >> 4, prow_blocks) * 16 + tb_byte_pos(i & 15)];
        state = tb_next_state(state, byte, force);
        if (state == 0) { push(0, 1); --i; --j; }
        else if (state == 1 || state == 3) { push(2, 1); --i; }
        else { push(1, 1); --j; }
    }
    if (i >= 0) push(2, i + 1);
    if (j >= 0) push(1, j + 1);
    const int n = (int)cig.size();
    uint32_t* out = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (int k = 0; k < n; ++k) out[k] = cig[n - 1 - k];
    *cigar_out = out;
    if (tb_out) {
        uint8_t* t = (uint8_t*)calloc(nrow * prow + 16, 1);
        for (size_t r = 0; r < nrow; ++r) {
            const RowGeom g = row_geom((int)r, qlen, tlen, w);
            const int nb = (g.en0 >> 4) - (g.st0 >> 4) + 1;
            for (int b = 0; b < nb; ++b)
                for (int c = 0; c < 16; ++c) {
                    const uint8_t x = p[r * prow + (size_t)tb_block_slot((g.st0 >> 4) + b, prow_blocks) * 16 + tb_byte_pos(c)];
                    t[r * prow + (size_t)b * 16 + c] = (uint8_t)((lesvksw::kTagS - (x & 7)) | ((x >> 7) & 1) << 3 | ((x >> 6) & 1) << 4 |
                                                             ((x >> 5) & 1) << 5 | ((x >> 4) & 1) << 6);
                }
        }
        *tb_out = t; *tb_row_bytes = (int)prow;
    }
    return n;
}

// Folded traceback of the column strips (ksw_core.cuh: fold_row_bytes): every block the kernel stores for a full-matrix
// alignment -- blocks st0 >> 4 .. en0 >> 4 of every row -- must get a slot of its own inside qlen x fold_row_bytes bytes.
// Returns the number of collisions plus out-of-range slots (0 = the layout holds).
extern "C" long lesv_emul_fold_check(int qlen, int tlen)
{
    using namespace lesvksw;
    const int w = qlen > tlen ? qlen : tlen;
    const size_t row_blocks = fold_row_bytes(qlen, tlen) / 16;
    std::vector<uint8_t> used((size_t)qlen * row_blocks, 0);
    long bad = 0;
    for (int r = 0; r < qlen + tlen - 1; ++r) {
        const RowGeom g = row_geom(r, qlen, tlen, w);
        const int pass = r / qlen;
        for (int B = g.st0 >> 4; B <= (g.en0 >> 4); ++B) {
            const size_t slot = (size_t)B + (size_t)pass;
            if (slot >= row_blocks) { ++bad; continue; }
            uint8_t& u = used[(size_t)(r % qlen) * row_blocks + slot];
            bad += u; u = 1;
        }
    }
    return bad + (fold_tb_bytes(qlen, tlen) != (size_t)qlen * row_blocks * 16);
}
