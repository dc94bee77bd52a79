// trace_emul.cpp -- host build of lesv_b200/csrc/trace_core.cuh (the code one GPU thread runs per
// alignment candidate).  TEST VEHICLE ONLY: tests/test_trace_emul.py compares it with the compiled
// reference's edlib_nw / InitHitFindData_SetupMapAlignInfoFromHit / hbn_traceback.
#include "../../lesv_b200/csrc/trace_core.cuh"
#include <vector>
#include <cstring>

using namespace lesvtb;

namespace {
struct Bufs {
    std::vector<uint64_t> P, M, peq;
    std::vector<int> score;
    std::vector<int16_t> tt, ts;
    Scratch sc;
    explicit Bufs(size_t cells = 60000, size_t peq_words = 1 << 16) : P(cells), M(cells), peq(peq_words), score(cells)
    {
        sc.P = P.data(); sc.M = M.data(); sc.score = score.data(); sc.peq = peq.data();
        sc.cap_cells = cells; sc.cap_peq = peq_words;
        tt.resize(32768); ts.resize(32768);
        dalign_tables(tt.data(), ts.data(), &sc.path_ave);
        sc.trim_table = tt.data(); sc.trim_score = ts.data();
    }
};
}

// ops (forward order) of edlib_nw(query, target); returns the count or a negative status
extern "C" int lesv_emul_nw(const uint8_t* q, int n, const uint8_t* t, int m, uint8_t* ops_out, int cap)
{
    Bufs B;
    return nw_path([&](int i) { return q[i]; }, n, [&](int j) { return t[j]; }, m, B.sc, ops_out, (size_t)cap);
}

extern "C" int lesv_emul_mems(const uint8_t* read, const uint8_t* window, const int* seed_w, const int* seed_r, int n_seeds,
                              int word_size, int* out /* q,s,len triples */, int cap)
{
    std::vector<Mem> mems(cap);
    const int n = extract_mems([&](int i) { return read[i]; }, [&](int j) { return window[j]; }, seed_w, seed_r, n_seeds,
                               word_size, mems.data(), cap);
    for (int i = 0; i < n; ++i) { out[3 * i] = mems[i].q; out[3 * i + 1] = mems[i].s; out[3 * i + 2] = mems[i].len; }
    return n;
}

// res[0..8) = status, qoff, qend, soff, send, aln_len, n_match, overhang flags (left | right << 1);
// ops_out receives the final ops (aln_len bytes)
extern "C" void lesv_emul_traceback(const uint8_t* read, int qsize, const uint8_t* window, int ssize,
                                    const int* mems3, int n_mems, int min_align_size, double min_ident, int over_hang,
                                    int* res, uint8_t* ops_out)
{
    Bufs B;
    std::vector<Mem> mems(n_mems), ts(n_mems + 1);
    for (int i = 0; i < n_mems; ++i) { mems[i].q = mems3[3 * i]; mems[i].s = mems3[3 * i + 1]; mems[i].len = mems3[3 * i + 2]; }
    const size_t cap = 3 * (size_t)qsize + 2 * (size_t)ssize + 4096;
    const size_t front = n_mems ? 2 * (size_t)(mems[0].q > mems[0].s ? mems[0].q : mems[0].s) + 64 : 64;
    std::vector<uint8_t> ops(cap), blk(kBlkOpsCap);
    const Result R = traceback([&](int i) { return read[i]; }, qsize, [&](int j) { return window[j]; }, ssize,
                               mems.data(), n_mems, ts.data(), B.sc, ops.data(), cap, front < cap / 2 ? front : cap / 2,
                               blk.data(), min_align_size, min_ident, over_hang != 0);
    res[0] = R.status; res[1] = R.qoff; res[2] = R.qend; res[3] = R.soff; res[4] = R.send; res[5] = R.aln_len;
    res[6] = R.n_match; res[7] = R.overhang_left | (R.overhang_right << 1);
    if (R.status == kOk || R.aln_len > 0) memcpy(ops_out, ops.data() + R.ops_beg, (size_t)R.aln_len);
}
