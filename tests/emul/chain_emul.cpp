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
