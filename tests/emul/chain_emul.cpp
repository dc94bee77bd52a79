// chain_emul.cpp -- host build of lesv_b200/csrc/chain_core.cuh (the code the chaining kernels run
// per segment).  TEST VEHICLE ONLY: tests/test_chain_emul.py compares it with the compiled
// reference's chaining_find_candidates on random and adversarial seed lists.
#include "../../lesv_b200/csrc/chain_core.cuh"
#include <vector>

using namespace lesvchain;

// hits_out: per hit 6 ints {score, qbeg, qend, sbeg, send, n_seeds}; seeds_out: concatenated seed
// indices of every hit, first seed first.  Returns the number of hits.
extern "C" int lesv_emul_chain(int n, const int* q, const int* s, const int* len, int is_mem,
                               int min_cnt, int min_score, int* hits_out, int* seeds_out)
{
    std::vector<Seed> seeds(n);
    for (int i = 0; i < n; ++i) { seeds[i].q = q[i]; seeds[i].s = s[i]; seeds[i].len = len[i]; }
    std::vector<int> f(n), p(n), t(n), v(n);
    std::vector<IntPair> u(n);
    Hit pre[kMaxChains], out[kMaxJoined];
    int members[kMaxChains], n_pre = 0;
    const Params P = make_params(min_cnt, min_score);
    const int nh = find_candidates(P, seeds.data(), n, is_mem != 0, f.data(), p.data(), t.data(), v.data(), u.data(),
                                   pre, &n_pre, out, members);
    int ns = 0;
    for (int h = 0; h < nh; ++h) {
        int cnt = 0;
        for (int m = 0; m < out[h].mem_cnt; ++m) {
            const Hit& c = pre[members[out[h].mem_off + m]];
            for (int x = c.v_cnt - 1; x >= 0; --x) { seeds_out[ns++] = v[c.v_off + x]; ++cnt; }
        }
        int* o = hits_out + 6 * h;
        o[0] = out[h].score; o[1] = out[h].qbeg; o[2] = out[h].qend; o[3] = (int)out[h].sbeg; o[4] = (int)out[h].send; o[5] = cnt;
    }
    return nh;
}

// ---- host emulation of chain_warp.cuh: score_seeds_warp, lane by lane -----------------------------------------
// The same steps as the device code with the warp intrinsics spelled out over arrays of 32 (ballot, prefix max,
// OR-reduction, shuffles).  tests/test_chain_emul.py checks f / p / v against the sequential score_seeds on
// random and tie-heavy seed lists, so the "exact replay" argument of chain_warp.cuh is tested, not just reasoned.
#include <climits>
static void score_seeds_warp_emul(const Params& P, const Seed* seeds, int n, bool is_mem, int* f, int* p, int* t, int* v)
{
    int sum_cov = 0;
    for (int i = 0; i < n; ++i) { sum_cov += seeds[i].len; t[i] = 0; f[i] = 0; p[i] = -1; }
    const int avg_cov = sum_cov / n;
    int st = 0;
    for (int i = 0; i < n; ++i) {
        const Seed si = seeds[i];
        const long long ri = si.s, qi = si.q;
        const int cov = si.len;
        int max_j = -1, max_f = cov, n_skip = 0;
        while (st < i && ri > (long long)seeds[st].s + seeds[st].len + P.max_dist_ref) ++st;
        bool stop = false;
        for (int jhi = i - 1; jhi >= st && !stop; jhi -= 32) {
            bool valid[32], marked[32]; int sc[32], pj[32];
            unsigned tgt_or = 0;
            for (int lane = 0; lane < 32; ++lane) {              // every lane reads memory BEFORE any lane writes t[]
                const int j = jhi - lane;
                valid[lane] = j >= st; sc[lane] = 0; pj[lane] = -1; marked[lane] = false;
                if (valid[lane]) {
                    const Seed sj = seeds[j];
                    if (is_mem) { if ((long long)sj.q + sj.len > qi || (long long)sj.s + sj.len > ri) valid[lane] = false; }
                    else { if (sj.q > qi || sj.s > ri) valid[lane] = false; }
                    const long long dr = ri - sj.s, dq = qi - sj.q;
                    if (dr == 0 || dq <= 0) valid[lane] = false;
                    if (dq > P.max_dist_qry || dr > P.max_dist_ref) valid[lane] = false;
                    const int dd = (int)(dr > dq ? dr - dq : dq - dr);
                    if (dd > P.band_width) valid[lane] = false;
                    if (valid[lane]) {
                        const int min_d = (int)(dq < dr ? dq : dr);
                        int x = min_d > cov ? cov : min_d;
                        const int log_dd = dd ? ilog2_32((uint32_t)dd) : 0;
                        if (!is_mem) { const double pen = dd * .01; x -= (int)(pen * avg_cov) + (log_dd >> 1); }
                        else x -= (log_dd >> 1);
                        x += f[j];
                        sc[lane] = x; pj[lane] = p[j]; marked[lane] = t[j] == i;
                    }
                }
                if (valid[lane] && pj[lane] >= 0 && pj[lane] > jhi - 32) tgt_or |= 1u << (jhi - pj[lane]);
            }
            unsigned um = 0, im = 0;
            int run = INT_MIN;                                    // exclusive prefix max of the valid scores
            bool upd[32];
            for (int lane = 0; lane < 32; ++lane) {
                if ((tgt_or >> lane) & 1u) marked[lane] = true;
                int excl = run;
                if (lane == 0 || excl < max_f) excl = max_f;
                upd[lane] = valid[lane] && sc[lane] > excl;
                if (upd[lane]) um |= 1u << lane;
                else if (valid[lane] && marked[lane]) im |= 1u << lane;
                const int x = valid[lane] ? sc[lane] : INT_MIN;
                if (x > run) run = x;
            }
            int brk = 32;
            for (unsigned ev = um | im; ev; ev &= ev - 1) {
                const int l = __builtin_ctz(ev);
                if ((um >> l) & 1u) { if (n_skip) --n_skip; }
                else if (++n_skip > P.max_skip) { brk = l; break; }
            }
            const unsigned before = brk >= 32 ? 0xFFFFFFFFu : ((1u << brk) - 1u);
            const unsigned um2 = um & before;
            if (um2) { const int l = 31 - __builtin_clz(um2); max_f = sc[l]; max_j = jhi - l; }
            for (int lane = 0; lane < 32; ++lane)
                if (valid[lane] && lane < brk && pj[lane] >= 0) t[pj[lane]] = i;
            if (brk < 32) stop = true;
        }
        f[i] = max_f; p[i] = max_j;
        v[i] = (max_j >= 0 && v[max_j] > max_f) ? v[max_j] : max_f;
    }
}

// returns the number of positions where f, p or v of the warp schedule differ from the sequential loop
extern "C" int lesv_emul_chain_warp_diff(int n, const int* q, const int* s, const int* len, int is_mem, int min_cnt, int min_score)
{
    if (n <= 0) return 0;
    std::vector<Seed> seeds(n);
    for (int i = 0; i < n; ++i) { seeds[i].q = q[i]; seeds[i].s = s[i]; seeds[i].len = len[i]; }
    std::vector<int> f(n), p(n), t(n), v(n), f2(n), p2(n), t2(n), v2(n);
    const Params P = make_params(min_cnt, min_score);
    score_seeds(P, seeds.data(), n, is_mem != 0, f.data(), p.data(), t.data(), v.data());
    score_seeds_warp_emul(P, seeds.data(), n, is_mem != 0, f2.data(), p2.data(), t2.data(), v2.data());
    int bad = 0;
    for (int i = 0; i < n; ++i) bad += f[i] != f2[i] || p[i] != p2[i] || v[i] != v2[i];
    return bad;
}
