// trace_core.cuh -- base-level alignment of one candidate (read strand x reference window) the way
// lesv's mapper builds it, as sequential host/device code: one GPU thread runs one candidate.
//
// Replaces (paths relative to the reference's src/):
//   InitHitFindData_SetupMapAlignInfoFromHit + s_extract_mem   algo/init_hit_finder.c:469-505, 255-297
//   compute_trace_points                                       algo/hbn_traceback.c:53-95
//   hbn_traceback (process_over_hang = 0 part)                 algo/hbn_traceback.c:313-544
//   edlib_nw / edlib_shw / edlib_extend                        algo/edlib_wrapper.c:32-111, 114-190, 238-346
//   edlibAlign, NW / SHW modes with EDLIB_TASK_PATH            algo/edlib.cpp:143-278 (vendored edlib)
//
// edlib restated.  edlibAlign computes the unit-cost edit distance with Myers' bit-vector algorithm
// inside an Ukkonen band and, for the path, re-runs it with k = the distance while storing every
// column, then walks back with the move priority "up (consume query), left (consume target),
// diagonal" (obtainAlignmentTraceback, edlib.cpp:929-1146).  Cells that lie on some optimal path are
// never pruned from that band and carry exact values, and a neighbour can only win a comparison if
// it lies on an optimal path itself, so the walk is a function of the exact DP matrix and the move
// priority alone.  Here: the same bit-vector recurrence (Myers 1999; calculateBlock, edlib.cpp:393-432)
// over full-height columns, every column stored, the same walk.  SHW = free end in the target:
// smallest target prefix with the minimal distance (edlib.cpp:640-661 keeps positions in ascending
// order and edlibAlign reports the first), accepted if <= k.
//
// Large problems (20 B x blocks x columns + 8 B x columns >= 1 MB, edlib.cpp:1175-1177) take edlib's
// Hirschberg route (obtainAlignmentHirschberg, :1218-1383): split the target in the middle, pick the
// FIRST query row whose forward + reverse scores add up to the distance, recurse; restated below with
// an explicit stack.  The overhang step (hbn_traceback.c:170-311, 517-532) runs DALIGNER's adaptive
// forward wave (Local_Alignment -> forward_wave, algo/align.c:386-1046) from the alignment's end to
// find how far the overhang aligns, then edlib_nw over that range; the wave is restated without its
// trace-point bookkeeping (only the end point and the difference count are consumed, dalign.c:88-107).
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TRC_HD __host__ __device__
#else
#define TRC_HD
#endif

namespace lesvtb {

enum { kOpDiag = 0, kOpIns = 1, kOpDel = 2 };      // Ins: read base over '-'; Del: '-' over reference base
enum Status { kOk = 1, kNoAlign = 0, kNeedsHirschberg = -1, kOpsOverflow = -2, kScratchOverflow = -3 };

struct Mem { int q, s, len; };                      // ChainSeed{qoff (read), soff (window), length}

// ---- Myers bit-vector columns --------------------------------------------------------------------
struct Scratch {
    uint64_t* P;            // [column][block]
    uint64_t* M;
    int* score;             // bottom cell of the block
    uint64_t* peq;          // [4][blocks]
    size_t cap_cells;       // block-columns that fit P/M/score
    size_t cap_peq;         // u64 words in peq
    int cols;               // columns of the last myers_fill: cell (column j, block b) sits at b * cols + j, so a walk
                            // along the diagonal stays inside a few cache lines
    const int16_t* trim_table;   // DALIGNER TABLE / SCORE (New_Align_Spec, align.c:253-296), 32768 entries each
    const int16_t* trim_score;
    int path_ave;
};
constexpr int kBlkOpsCap = 4096;   // ops of one SHW block: <= 1123 + 1459 columns (get_next_sequence_block)

TRC_HD inline int popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// one block of one column (calculateBlock): returns hout
TRC_HD inline int myers_block(uint64_t Pv, uint64_t Mv, uint64_t Eq, int hin, uint64_t& Pout, uint64_t& Mout)
{
    const uint64_t hin_neg = (uint64_t)(hin < 0);
    const uint64_t Xv = Eq | Mv;
    Eq |= hin_neg;
    const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
    uint64_t Ph = Mv | ~(Xh | Pv);
    uint64_t Mh = Pv & Xh;
    const int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
    Ph <<= 1; Mh <<= 1;
    Mh |= hin_neg;
    Ph |= (uint64_t)(hin > 0);
    Pout = Mh | ~(Xv | Ph);
    Mout = Ph & Xv;
    return hout;
}

// D[i][j] for 1 <= i <= n from the stored column j (1-based), boundary D[i][0] = i, D[0][j] = j
TRC_HD inline int cell_value(const Scratch& sc, int nb, int j, int i)
{
    if (j == 0) return i;
    if (i == 0) return j;
    const int b = (i - 1) >> 6, pos = (i - 1) & 63;
    const size_t x = (size_t)b * sc.cols + (j - 1);
    const uint64_t above = pos == 63 ? 0ull : (~0ull << (pos + 1));     // rows below row i inside the block
    return sc.score[x] - popc64(sc.P[x] & above) + popc64(sc.M[x] & above);
}

// Fills columns 1..m for query Q (n bases) x target T (m bases).  Q(i), T(j): base codes 0..3.
template <class FQ, class FT>
TRC_HD inline bool myers_fill(FQ Q, int n, FT T, int m, Scratch& sc, int* nb_out)
{
    const int nb = (n + 63) >> 6;
    *nb_out = nb;
    if ((size_t)nb * (size_t)m > sc.cap_cells || (size_t)4 * nb > sc.cap_peq) return false;
    sc.cols = m;
    for (int x = 0; x < 4 * nb; ++x) sc.peq[x] = 0;
    for (int i = 0; i < n; ++i) sc.peq[(size_t)Q(i) * nb + (i >> 6)] |= 1ull << (i & 63);
    if (n & 63) {                                    // rows past the query match anything (buildPeq, :350-353)
        const uint64_t pad = ~0ull << (n & 63);
        for (int c = 0; c < 4; ++c) sc.peq[(size_t)c * nb + nb - 1] |= pad;
    }
    for (int j = 0; j < m; ++j) {
        const uint64_t* eq = sc.peq + (size_t)T(j) * nb;
        int hout = 1;
        for (int b = 0; b < nb; ++b) {
            const size_t x = (size_t)b * m + j;
            const uint64_t Pv = j ? sc.P[x - 1] : ~0ull, Mv = j ? sc.M[x - 1] : 0ull;
            const int prev = j ? sc.score[x - 1] : (b + 1) * 64;
            hout = myers_block(Pv, Mv, eq[b], hout, sc.P[x], sc.M[x]);
            sc.score[x] = prev + hout;
        }
    }
    return true;
}

// Walk from (n, jend) back to (0, 0); writes the ops in FORWARD order into out[0..ret).
TRC_HD inline int myers_walk(const Scratch& sc, int nb, int n, int jend, uint8_t* out)
{
    int i = n, j = jend, cur = cell_value(sc, nb, jend, n), len = 0;
    uint8_t* rev = out;                              // written last-op-first, reversed in place below
    while (i > 0 && j > 0) {
        const int b = (i - 1) >> 6, pos = (i - 1) & 63;
        const size_t x = (size_t)b * sc.cols + (j - 1);
        const int u = cur - (int)((sc.P[x] >> pos) & 1) + (int)((sc.M[x] >> pos) & 1);
        if (u + 1 == cur) { rev[len++] = kOpIns; --i; cur = u; continue; }
        const int l = cell_value(sc, nb, j - 1, i);
        if (l + 1 == cur) { rev[len++] = kOpDel; --j; cur = l; continue; }
        const int ul = cell_value(sc, nb, j - 1, i - 1);
        rev[len++] = kOpDiag; --i; --j; cur = ul;
    }
    while (i > 0) { rev[len++] = kOpIns; --i; }
    while (j > 0) { rev[len++] = kOpDel; --j; }
    for (int x = 0; x < len / 2; ++x) { const uint8_t t = out[x]; out[x] = out[len - 1 - x]; out[len - 1 - x] = t; }
    return len;
}

// alignmentDataSize of obtainAlignment (edlib.cpp:1175-1176)
TRC_HD inline bool needs_hirschberg(int n, int m)
{
    const long long nb = (n + 63) >> 6;
    return 20ll * nb * m + 8ll * m >= 1024 * 1024;
}

// Last column of the NW matrix of Q[0..n) x T[0..cols): D[i] = D(i, cols), i = 0..n.  Rolling columns:
// uses sc.P[0..nb), sc.M[0..nb), sc.peq.
template <class FQ, class FT>
TRC_HD inline bool myers_last_column(FQ Q, int n, FT T, int cols, Scratch& sc, int* D)
{
    const int nb = (n + 63) >> 6;
    if ((size_t)nb > sc.cap_cells || (size_t)4 * nb > sc.cap_peq) return false;
    for (int x = 0; x < 4 * nb; ++x) sc.peq[x] = 0;
    for (int i = 0; i < n; ++i) sc.peq[(size_t)Q(i) * nb + (i >> 6)] |= 1ull << (i & 63);
    if (n & 63) {
        const uint64_t pad = ~0ull << (n & 63);
        for (int c = 0; c < 4; ++c) sc.peq[(size_t)c * nb + nb - 1] |= pad;
    }
    for (int b = 0; b < nb; ++b) { sc.P[b] = ~0ull; sc.M[b] = 0ull; }
    for (int j = 0; j < cols; ++j) {
        const uint64_t* eq = sc.peq + (size_t)T(j) * nb;
        int hout = 1;
        for (int b = 0; b < nb; ++b) hout = myers_block(sc.P[b], sc.M[b], eq[b], hout, sc.P[b], sc.M[b]);
    }
    D[0] = cols;
    for (int i = 1; i <= n; ++i) {
        const int b = (i - 1) >> 6, pos = (i - 1) & 63;
        D[i] = D[i - 1] + (int)((sc.P[b] >> pos) & 1) - (int)((sc.M[b] >> pos) & 1);
    }
    return true;
}

// obtainAlignment (edlib.cpp:1148-1216) for Q[0..n) x T[0..m) with known distance `best`.
template <class FQ, class FT>
TRC_HD inline int obtain_alignment(FQ Q, int n, FT T, int m, int best, Scratch& sc, uint8_t* out, size_t out_cap)
{
    struct Frame { int q0, n, t0, m, best; };
    Frame stack[48];
    int top = 0;
    size_t len = 0;
    stack[top++] = Frame{0, n, 0, m, best};
    while (top > 0) {
        const Frame f = stack[--top];
        if (len + (size_t)f.n + (size_t)f.m > out_cap) return kOpsOverflow;
        if (f.n == 0 || f.m == 0) {                  // :1154-1161
            for (int x = 0; x < f.n + f.m; ++x) out[len++] = f.n == 0 ? kOpDel : kOpIns;
            continue;
        }
        auto fq = [&](int i) { return Q(f.q0 + i); };
        auto ft = [&](int j) { return T(f.t0 + j); };
        if (!needs_hirschberg(f.n, f.m)) {
            int nb;
            if (!myers_fill(fq, f.n, ft, f.m, sc, &nb)) return kScratchOverflow;
            len += (size_t)myers_walk(sc, nb, f.n, f.m, out + len);
            continue;
        }
        // obtainAlignmentHirschberg (:1218-1383)
        if (2 * ((size_t)f.n + 1) > sc.cap_cells || top + 2 > 48) return kScratchOverflow;
        const int lhw = f.m / 2, rhw = f.m - lhw;
        int* L = sc.score;                            // L[i] = cost of Q[0..i) against the left half
        int* R = sc.score + f.n + 1;                  // R[i] = cost of the LAST i bases against the right half
        if (!myers_last_column(fq, f.n, ft, lhw, sc, L)) return kScratchOverflow;
        auto rq = [&](int i) { return Q(f.q0 + f.n - 1 - i); };
        auto rt = [&](int j) { return T(f.t0 + f.m - 1 - j); };
        if (!myers_last_column(rq, f.n, rt, rhw, sc, R)) return kScratchOverflow;
        int row = -2, ls = 0, rs = 0;
        for (int r = 0; r <= f.n - 2; ++r)            // first interior row whose halves add up (:1300-1310)
            if (L[r + 1] + R[f.n - r - 1] == f.best) { row = r; ls = L[r + 1]; rs = R[f.n - r - 1]; break; }
        if (row == -2 && lhw + R[f.n] == f.best) { row = -1; ls = lhw; rs = R[f.n]; }            // :1312-1319
        if (row == -2 && L[f.n] + rhw == f.best) { row = f.n - 1; ls = L[f.n]; rs = rhw; }        // :1320-1328
        if (row == -2) return kNoAlign;
        const int ul_h = row + 1;
        stack[top++] = Frame{f.q0 + ul_h, f.n - ul_h, f.t0 + lhw, rhw, rs};     // lower right, emitted second
        stack[top++] = Frame{f.q0, ul_h, f.t0, lhw, ls};                        // upper left first
    }
    return (int)len;
}

// edlib_nw (edlib_wrapper.c:32-111).  Returns ops appended (>= 0) or a negative Status.
template <class FQ, class FT>
TRC_HD inline int nw_path(FQ Q, int n, FT T, int m, Scratch& sc, uint8_t* out, size_t out_cap)
{
    if (n == 0 || m == 0) return 0;                  // :43: "return 1" with both strings empty
    if ((size_t)n + (size_t)m > out_cap) return kOpsOverflow;
    if (!needs_hirschberg(n, m)) {
        int nb;
        if (!myers_fill(Q, n, T, m, sc, &nb)) return kScratchOverflow;
        return myers_walk(sc, nb, n, m, out);        // tolerance 0.5*max, then max: never fails (:58-72)
    }
    if ((size_t)n + 1 > sc.cap_cells) return kScratchOverflow;
    if (!myers_last_column(Q, n, T, m, sc, sc.score)) return kScratchOverflow;
    return obtain_alignment(Q, n, T, m, sc.score[n], sc, out, out_cap);
}

// edlib_shw (edlib_wrapper.c:114-190): 1 + ops in out, *tend = target bases consumed; 0 = no alignment
template <class FQ, class FT>
TRC_HD inline int shw_path(FQ Q, int n, FT T, int m, Scratch& sc, uint8_t* out, int* n_ops, int* tend)
{
    *n_ops = 0; *tend = 0;
    if (n == 0 || m == 0) return 0;
    if (n + m > kBlkOpsCap) return kScratchOverflow;
    const int k = (int)((n > m ? n : m) * 0.35);     // :142
    int nb;
    if (!myers_fill(Q, n, T, m, sc, &nb)) return kScratchOverflow;
    int best = -1, bj = 0;
    for (int j = 1; j <= m; ++j) {
        const int v = cell_value(sc, nb, j, n);
        if (v <= k && (best < 0 || v < best)) { best = v; bj = j; }
    }
    if (best < 0) return 0;
    if (needs_hirschberg(n, bj)) return kNeedsHirschberg;     // cannot happen: blocks are <= 1459 x 1459
    *n_ops = myers_walk(sc, nb, n, bj, out);
    *tend = bj;
    return 1;
}

// ---- MEMs of a chain (init_hit_finder.c:469-505, 255-297) -------------------------------------------
// chain seeds: (w = window offset, r = read offset), word_size bases each, in chain order.
// s_extract_mem (init_hit_finder.c:255-297) over one co-diagonal run of seeds: ml bases starting at read
// offset ri / window offset wi.  Appends to mems[n..] (mems == nullptr: count only); returns the new n, -1 if full.
template <class FQ, class FS>
TRC_HD inline int extract_mem_run(FQ Q, FS S, int ri, int wi, int ml, int word_size, Mem* mems, int cap, int n)
{
    int x = 0;
    while (x < ml) {
        while (x < ml) { if (Q(ri) == S(wi)) break; ++ri; ++wi; ++x; }
        if (ml - x < word_size) break;
        int y = x;
        while (y < ml) { if (Q(ri) != S(wi)) break; ++ri; ++wi; ++y; }
        if (y - x >= word_size) {
            if (mems) {
                if (n >= cap) return -1;
                mems[n].len = y - x; mems[n].q = ri - (y - x); mems[n].s = wi - (y - x);
            }
            ++n;
        }
        x = y;
    }
    return n;
}

template <class FQ, class FS>
TRC_HD inline int extract_mems(FQ Q, FS S, const int* seed_w, const int* seed_r, int n_seeds, int word_size,
                               Mem* mems, int cap)
{
    int n = 0, i = 0;
    while (i < n_seeds) {
        int j = i + 1;
        while (j < n_seeds && seed_w[j] - seed_w[i] == seed_r[j] - seed_r[i]) ++j;
        const int ml = seed_w[j - 1] + word_size - seed_w[i];
        n = extract_mem_run(Q, S, seed_r[i], seed_w[i], ml, word_size, mems, cap, n);
        if (n < 0) return -1;
        i = j;
    }
    return n;
}

// compute_trace_points (hbn_traceback.c:53-95), in place: returns the number of trace seeds kept in ts[]
TRC_HD inline int trace_points(const Mem* seeds, int n, Mem* ts)
{
    if (n == 0) return 0;
    int nt = 0, i = 0;
    while (i < n) {
        int f = i;
        while (f < n - 1) {
            const int dq = seeds[f + 1].q - seeds[f].q, ds = seeds[f + 1].s - seeds[f].s;
            const int d = dq - ds;
            const bool r1 = (d < 0 ? -d : d) <= 100;
            double rel = 1.0 - 1.0 * dq / ds;
            if (rel < 0) rel = -rel;
            const bool r2 = rel <= 0.21;
            if (!r1 || !r2) break;
            ++f;
        }
        if (f > i) for (int k = i; k <= f; ++k) ts[nt++] = seeds[k];
        else if (i == 0) ts[nt++] = seeds[0];
        i = f + 1;
    }
    if (ts[nt - 1].q < seeds[n - 1].q) ts[nt++] = seeds[n - 1];
    return nt;
}

// ---- DALIGNER forward wave (align.c:386-1046) without trace points ---------------------------------
constexpr int kTrimLen = 15, kPathLen = 60, kTrimMlag = 200, kWaveLag = 30;
constexpr uint64_t kPathTop = 0x1000000000000000ull, kPathInt = 0x0fffffffffffffffull;
constexpr int kTrimMask = 0x7fff;
constexpr int kWaveHalf = 6000;                      // diagonals kept on either side of the start

// New_Align_Spec(1 - 0.35, 100, {.25,.25,.25,.25}, 1) (dalign.c:16, align.c:253-296): tables and ave_path.
// Host only; the device gets the filled tables.
inline void dalign_tables(int16_t* table, int16_t* score, int* path_ave)
{
    const double ave_corr = 1.0 - 0.35, bias_factor = 1.000;     // match = .5 -> bias index 9
    *path_ave = (int)(kPathLen * (1. - bias_factor * (1. - ave_corr)));
    const int mscore = (int)(1000 * bias_factor * (1. - ave_corr)), dscore = 1000 - mscore;
    // set_table (:246-259) unrolled: prefix bits are consumed most significant first
    for (int prefix = 0; prefix <= kTrimMask; ++prefix) {
        int sc = 0, mx = 0;
        for (int bit = kTrimLen - 1; bit >= 0; --bit) {
            if (sc > mx) mx = sc;
            sc += ((prefix >> bit) & 1) ? mscore : -dscore;
        }
        table[prefix] = (int16_t)(sc - mx);
        score[prefix] = (int16_t)sc;
    }
}

// A(i), B(j): base codes, 4 outside [0, len) (the sentinels dalign_align puts around the copies).
// Start at (0, 0) on diagonal 0.  V/M: 2*kWaveHalf+3 ints each, T: as many u64.  Returns false if the
// wave leaves the diagonal window.
template <class FA, class FB>
TRC_HD inline bool dalign_forward(FA A, FB B, const Scratch& sc, int* Vb, int* Mb, uint64_t* Tb,
                                  int* aepos, int* bepos, int* diffs)
{
    const int16_t* TABLE = sc.trim_table;
    const int16_t* SCORE = sc.trim_score;
    const int PATH_AVE = sc.path_ave;
    int* V = Vb + kWaveHalf + 1; int* M = Mb + kWaveHalf + 1; uint64_t* T = Tb + kWaveHalf + 1;
    const int IMAX = 0x7fffffff;
    int hgh = 0, low = 0, dif = 0;
    int more = 1, aclip = IMAX, bclip = -IMAX;
    const int mida = 0;
    int besta = mida, trima = mida, morea = mida, lasta = mida;
    int besty = 0, trimy = 0, morey = 0;
    int trimd = 0, mored = 0, morem = -1;
    {
        const int k = 0;
        int y = 0;
        for (;;) {
            const int c = B(y);
            if (c == 4) { more = 0; if (bclip < k) bclip = k; break; }
            const int d = A(y + k);
            if (c != d) { if (d == 4) { more = 0; aclip = k; } break; }
            y += 1;
        }
        const int c = (y << 1) + k;
        if (c > besta) { besta = trima = lasta = c; besty = trimy = y; }
        V[k] = c; T[k] = kPathInt; M[k] = kPathLen;
    }
    auto clip_fix = [&]() {
        if (B(besty) != 4 && A(besta - besty) != 4) more = 1;
        if (hgh >= aclip) {
            hgh = aclip - 1;
            if (morem <= M[aclip]) { morem = M[aclip]; morea = V[aclip]; morey = (morea - aclip) / 2; mored = dif; }
        }
        if (low <= bclip) {
            low = bclip + 1;
            if (morem <= M[bclip]) { morem = M[bclip]; morea = V[bclip]; morey = (morea - bclip) / 2; mored = dif; }
        }
        aclip = IMAX; bclip = -IMAX;
    };
    if (more == 0) clip_fix();
    while (more && lasta >= besta - kTrimMlag) {
        low -= 1; hgh += 1;
        if (low <= -kWaveHalf || hgh >= kWaveHalf) return false;
        V[low] = -1;
        int am = -1;
        V[hgh] = -1;
        dif += 1;
        int ac = -1, ap;
        V[hgh + 1] = V[low - 1] = -1;
        uint64_t t = kPathInt;
        int n = kPathLen;
        for (int k = hgh; k >= low; --k) {
            int c, m;
            uint64_t b;
            ap = ac; ac = am; am = V[k - 1];
            if (ac < am) {
                if (am < ap) { c = ap + 1; m = n; b = t; }
                else { c = am + 1; m = M[k - 1]; b = T[k - 1]; }
            } else {
                if (ac < ap) { c = ap + 1; m = n; b = t; }
                else { c = ac + 2; m = M[k]; b = T[k]; }
            }
            if ((b & kPathTop) != 0) m -= 1;
            b <<= 1;
            int y = (c - k) >> 1;
            for (;;) {
                const int cb = B(y);
                if (cb == 4) { more = 0; if (bclip < k) bclip = k; break; }
                const int d = A(y + k);
                if (cb != d) { if (d == 4) { more = 0; aclip = k; } break; }
                y += 1;
                if ((b & kPathTop) == 0) m += 1;
                b = (b << 1) | 1;
            }
            c = (y << 1) + k;
            if (c > besta) {
                besta = c; besty = y;
                if (m >= PATH_AVE) {
                    lasta = c;
                    if (TABLE[b & kTrimMask] >= 0)
                        if (TABLE[(b >> kTrimLen) & kTrimMask] + SCORE[b & kTrimMask] >= 0) { trima = c; trimy = y; trimd = dif; }
                }
            }
            t = T[k]; n = M[k];
            V[k] = c; T[k] = b; M[k] = m;
        }
        if (more == 0) clip_fix();
        const int lim = besta - kWaveLag;
        while (hgh >= low) {
            if (V[hgh] < lim) hgh -= 1;
            else { while (V[low] < lim) low += 1; break; }
        }
    }
    int trimx;
    if (morem >= 0) { trimx = morea - morey; trimy = morey; trimd = mored; }     // REACH = 1
    else trimx = trima - trimy;
    *aepos = trimx; *bepos = trimy; *diffs = trimd;
    return true;
}

// overhang_extend (hbn_traceback.c:170-212): how far do Q and S align from (0,0) (dalign_align with
// min_align_size 1, min_ident_perc 0.65), then edlib_nw over that range.  Returns ops written, 0 = the
// reference's "return 0", < 0 = Status.
template <class FQ, class FS>
TRC_HD inline int overhang_extend(FQ Q, int qlen, FS S, int slen, Scratch& sc, uint8_t* out, size_t out_cap)
{
    if (sc.cap_cells < (size_t)(4 * kWaveHalf + 8)) return kScratchOverflow;
    int* V = sc.score; int* M = sc.score + 2 * kWaveHalf + 4;
    auto A = [&](int i) -> int { return (i >= 0 && i < qlen) ? (int)Q(i) : 4; };
    auto B = [&](int j) -> int { return (j >= 0 && j < slen) ? (int)S(j) : 4; };
    int qe = 0, se = 0, diffs = 0;
    if (!dalign_forward(A, B, sc, V, M, sc.P, &qe, &se, &diffs)) return kScratchOverflow;
    if (!(qe >= 1 && se >= 1)) return 0;                                  // dalign.c:92-94
    const double ident = 100.0 - 200.0 * diffs / (qe + se);
    if (!(ident >= 0.65)) return 0;
    return nw_path(Q, qe, S, se, sc, out, out_cap);
}

// ---- the alignment under construction -----------------------------------------------------------
struct Aln {
    uint8_t* ops; size_t cap;
    size_t beg, end;        // ops[beg, end)
    int qoff, qend, soff, send;
};

struct Result {
    int status;
    int qoff, qend, soff, send;
    int aln_len, n_match;
    size_t ops_beg;         // where the final ops start inside the candidate's ops buffer
    int overhang_left, overhang_right;   // 1 = the overhang step (left_extend / right_extend) ran its wave here
};

constexpr int kMatLen = 8, kMaxOverHang = 3000, kExtendBlock = 1024;

// edlib_extend (edlib_wrapper.c:238-346).  FQ/FT give the i-th base in extension direction.  The kept
// columns go to out[0..*n_out) in extension order; *qe / *te = bases consumed.  Returns Status.
template <class FQ, class FT, class Emit>
TRC_HD inline int extend(FQ Q, int qsize, FT T, int tsize, Scratch& sc, Emit emit, uint8_t* blk, int* qe, int* te)
{
    int qidx = 0, tidx = 0, qcons = 0, tcons = 0;
    for (;;) {
        const int qleft = qsize - qidx, tleft = tsize - tidx;
        int qblk, tblk;
        bool last_block;
        if (qleft < kExtendBlock + 100 || tleft < kExtendBlock + 100) {     // get_next_sequence_block :192-236
            qblk = (int)(tleft * 1.3); qblk = qblk < qleft ? qblk : qleft;
            tblk = (int)(qleft * 1.3); tblk = tblk < tleft ? tblk : tleft;
            last_block = true;
        } else { qblk = kExtendBlock; tblk = kExtendBlock; last_block = false; }
        if (qblk == 0 || tblk == 0) break;
        auto fq = [&](int i) { return Q(qidx + i); };
        auto ft = [&](int i) { return T(tidx + i); };
        int n_ops = 0, tfae = 0;
        const int r = shw_path(fq, qblk, ft, tblk, sc, blk, &n_ops, &tfae);
        if (r < 0) return r;
        const int qfae = r ? qblk : 0;
        bool done = last_block;
        if (qblk - qfae > 100 || tblk - tfae > 100) done = true;
        // last run of 8 identical columns, scanning from the end (:289-301)
        int acnt = 0, qcnt = 0, tcnt = 0, k = n_ops - 1, m = 0;
        {
            int qi = qfae, ti = tfae;                // bases consumed after column k
            while (k >= 0) {
                const uint8_t op = blk[k];
                bool same = false;
                if (op == kOpDiag) { --qi; --ti; ++qcnt; ++tcnt; same = fq(qi) == ft(ti); }
                else if (op == kOpIns) { --qi; ++qcnt; }
                else { --ti; ++tcnt; }
                m = same ? m + 1 : 0;
                ++acnt;
                if (m == kMatLen) break;
                --k;
            }
        }
        int keep;
        if (m != kMatLen || k < 1) {                 // :303-317: fall back to the exact common prefix
            keep = 0;
            for (int i = 0; i < qblk && i < tblk; ++i) {
                if (fq(i) != ft(i)) break;
                blk[keep++] = kOpDiag;
            }
            done = true;
            qcons += keep; tcons += keep;
        } else {
            keep = n_ops - acnt;
            qidx += qfae - qcnt; tidx += tfae - tcnt;
            qcons = qidx; tcons = tidx;
            if (done) { keep += kMatLen; qcons += kMatLen; tcons += kMatLen; }
        }
        for (int i = 0; i < keep; ++i)
            if (!emit(blk[i])) return kOpsOverflow;
        if (done) break;
    }
    *qe = qcons; *te = tcons;
    return kOk;
}

// hbn_traceback (hbn_traceback.c:313-544) without the overhang step.  Q: read strand (qsize bases),
// S: window (ssize bases).  seeds = MEMs in chain order.  ops buffer: cap bytes; `front` = room kept
// before the first column for the left extension.
// The gaps between consecutive trace seeds, in the order hbn_traceback aligns them (:351-374):
// f(i, query start, query bases, window start, window bases) for i = 0 .. tsc-2.  The gaps are independent
// of each other, so they can be solved anywhere (one GPU thread each) before the alignment is assembled.
template <class F>
TRC_HD inline void enumerate_gaps(const Mem* tsa, int tsc, F f)
{
    const int E = 10, E2 = 20;
    int qfrom, sfrom;
    if (tsa[0].len > E2) { qfrom = tsa[0].q + tsa[0].len - E; sfrom = tsa[0].s + tsa[0].len - E; }
    else { qfrom = tsa[0].q; sfrom = tsa[0].s; }
    for (int i = 0; i < tsc - 1; ++i) {
        const Mem sj = tsa[i + 1];
        int qto, sto;
        if (sj.len > E2) { qto = sj.q + E; sto = sj.s + E; }
        else { qto = sj.q + sj.len / 2; sto = sj.s + sj.len / 2; }
        f(i, qfrom, qto - qfrom, sfrom, sto - sfrom);
        qfrom = qto; sfrom = sto;
        if (sj.len > E2) { qfrom = sj.q + sj.len - E; sfrom = sj.s + sj.len - E; }
    }
}

template <class FQ, class FS>
TRC_HD inline Result finish_ends(FQ Q, int qsize, FS S, int ssize, Scratch& sc, uint8_t* ops, size_t cap, size_t beg, size_t end,
                                 int qoff, int qend, int soff, int send, uint8_t* blk, bool process_over_hang);

// ident_perc = 100 * matches / columns; accept iff long and similar enough (hbn_traceback.c:534-542)
TRC_HD inline void finalize_stats(Result& R, int n_match, int min_align_size, double min_ident_perc)
{
    R.n_match = n_match;
    const double ident = 100.0 * n_match / R.aln_len;
    R.status = (R.aln_len >= min_align_size && ident >= min_ident_perc) ? kOk : kNoAlign;
}

// hbn_traceback over known trace seeds tsa[0..tsc).  gap(i, qa, n, sa, m, out, cap) delivers the ops of
// edlib_nw(Q[qa..qa+n), S[sa..sa+m)) for the i-th gap (count, or a negative Status).
template <class FQ, class FS, class FG>
TRC_HD inline Result traceback_ts(FQ Q, int qsize, FS S, int ssize, const Mem* tsa, int tsc, FG gap,
                                  Scratch& sc, uint8_t* ops, size_t cap, size_t front, uint8_t* blk,
                                  int min_align_size, double min_ident_perc, bool process_over_hang);

template <class FQ, class FS>
TRC_HD inline Result traceback(FQ Q, int qsize, FS S, int ssize, const Mem* mems, int n_mems, Mem* ts_buf,
                               Scratch& sc, uint8_t* ops, size_t cap, size_t front, uint8_t* blk,
                               int min_align_size, double min_ident_perc, bool process_over_hang)
{
    if (n_mems == 0) {
        Result R;
        R.status = kNoAlign; R.qoff = R.qend = R.soff = R.send = 0; R.aln_len = 0; R.n_match = 0; R.ops_beg = 0;
        R.overhang_left = R.overhang_right = 0;
        return R;
    }
    const int tsc = trace_points(mems, n_mems, ts_buf);
    auto gap = [&](int, int qa, int n, int sa, int m, uint8_t* out, size_t out_cap) {
        auto fq = [&](int x) { return Q(qa + x); };
        auto fs = [&](int x) { return S(sa + x); };
        return nw_path(fq, n, fs, m, sc, out, out_cap);
    };
    return traceback_ts(Q, qsize, S, ssize, ts_buf, tsc, gap, sc, ops, cap, front, blk, min_align_size, min_ident_perc,
                        process_over_hang);
}

template <class FQ, class FS, class FG>
TRC_HD inline Result traceback_ts(FQ Q, int qsize, FS S, int ssize, const Mem* tsa, int tsc, FG gap,
                                  Scratch& sc, uint8_t* ops, size_t cap, size_t front, uint8_t* blk,
                                  int min_align_size, double min_ident_perc, bool process_over_hang)
{
    Result R;
    R.status = kNoAlign; R.qoff = R.qend = R.soff = R.send = 0; R.aln_len = 0; R.n_match = 0; R.ops_beg = 0;
    R.overhang_left = R.overhang_right = 0;
    if (tsc == 0) return R;
    size_t beg = front, end = front;
    const int E = 10, E2 = 20;
    int qfrom, qto, sfrom, sto;
#define TRC_PUSH_DIAG(cnt)                                                  \
    do {                                                                    \
        const int c__ = (cnt);                                              \
        if (end + (size_t)(c__ > 0 ? c__ : 0) > cap) { R.status = kOpsOverflow; return R; } \
        for (int x__ = 0; x__ < c__; ++x__) ops[end++] = kOpDiag;           \
    } while (0)
    if (tsa[0].len > E2) {
        qfrom = tsa[0].q; qto = qfrom + tsa[0].len; sfrom = tsa[0].s; sto = sfrom + tsa[0].len;
        TRC_PUSH_DIAG(qto - E - qfrom);
        qfrom = qto - E; sfrom = sto - E;
    } else { qfrom = tsa[0].q; sfrom = tsa[0].s; }
    for (int i = 0; i < tsc - 1; ++i) {
        const Mem sj = tsa[i + 1];
        if (sj.len > E2) { qto = sj.q + E; sto = sj.s + E; }
        else { qto = sj.q + sj.len / 2; sto = sj.s + sj.len / 2; }
        {
            const int r = gap(i, qfrom, qto - qfrom, sfrom, sto - sfrom, ops + end, cap - end);
            if (r < 0) { R.status = r; return R; }
            end += (size_t)r;
        }
        qfrom = qto; sfrom = sto;
        if (sj.len > E2) {
            qto = sj.q + sj.len - E; sto = sj.s + sj.len - E;
            TRC_PUSH_DIAG(qto - qfrom);
            qfrom = qto; sfrom = sto;
        }
    }
    {
        const Mem se = tsa[tsc - 1];
        qto = se.q + se.len; sto = se.s + se.len;
        TRC_PUSH_DIAG(qto - qfrom);
    }
#undef TRC_PUSH_DIAG
    R = finish_ends(Q, qsize, S, ssize, sc, ops, cap, beg, end, tsa[0].q, qto, tsa[0].s, sto, blk, process_over_hang);
    if (R.status != kOk) return R;
    // calc_ident_perc (hbn_traceback_aux.c:4-19)
    int n_match = 0;
    {
        int qi = R.qoff, si = R.soff;
        const size_t e = R.ops_beg + (size_t)R.aln_len;
        for (size_t a = R.ops_beg; a < e; ++a) {
            const uint8_t op = ops[a];
            if (op == kOpDiag) { n_match += Q(qi) == S(si); ++qi; ++si; }
            else if (op == kOpIns) ++qi;
            else ++si;
        }
    }
    finalize_stats(R, n_match, min_align_size, min_ident_perc);
    return R;
}

// Everything hbn_traceback does after the seeds and gaps are stitched (:396-532): trim to the first / last
// run of kMatLen identical columns, edlib_extend to both sides, the overhang step.  ops[beg, end) holds the
// stitched columns of Q[qoff, qend) x S[soff, send).  Returns status kOk with coordinates, aln_len and
// ops_beg set (n_match still 0: see finalize_stats), kNoAlign, or a negative Status.
template <class FQ, class FS>
TRC_HD inline Result finish_ends(FQ Q, int qsize, FS S, int ssize, Scratch& sc, uint8_t* ops, size_t cap, size_t beg, size_t end,
                                 int qoff, int qend, int soff, int send, uint8_t* blk, bool process_over_hang)
{
    Result R;
    R.status = kNoAlign; R.qoff = R.qend = R.soff = R.send = 0; R.aln_len = 0; R.n_match = 0; R.ops_beg = 0;
    R.overhang_left = R.overhang_right = 0;

    // trim to the first / last run of kMatLen identical columns (:396-433)
    {
        int qcnt = 0, scnt = 0, m = 0;
        size_t a = beg;
        int qi = qoff, si = soff;
        while (a < end) {
            const uint8_t op = ops[a++];
            bool same = false;
            if (op == kOpDiag) { same = Q(qi) == S(si); ++qi; ++si; ++qcnt; ++scnt; }
            else if (op == kOpIns) { ++qi; ++qcnt; }
            else { ++si; ++scnt; }
            m = same ? m + 1 : 0;
            if (m == kMatLen) break;
        }
        if (m < kMatLen) return R;
        beg = a - kMatLen;
        qoff += qcnt - kMatLen; soff += scnt - kMatLen;
        qcnt = scnt = 0; m = 0;
        a = end; qi = qend; si = send;
        while (a > beg) {
            const uint8_t op = ops[--a];
            bool same = false;
            if (op == kOpDiag) { --qi; --si; ++qcnt; ++scnt; same = Q(qi) == S(si); }
            else if (op == kOpIns) { --qi; ++qcnt; }
            else { --si; ++scnt; }
            m = same ? m + 1 : 0;
            if (m == kMatLen) break;
        }
        if (a == beg) return R;
        end = a + kMatLen;
        qend -= qcnt - kMatLen; send -= scnt - kMatLen;
    }
    // extend to the left (:444-476): fragments run backwards from qoff-1 / soff-1
    {
        const int q0 = qoff - 1, s0 = soff - 1;
        auto fq = [&](int x) { return Q(q0 - x); };
        auto fs = [&](int x) { return S(s0 - x); };
        int qe = 0, te = 0;
        auto emit = [&](uint8_t op) { if (beg == 0) return false; ops[--beg] = op; return true; };   // :458-465
        const int r = extend(fq, qoff, fs, soff, sc, emit, blk, &qe, &te);
        if (r < 0) { R.status = r; return R; }
        qoff -= qe; soff -= te;
    }
    // extend to the right (:478-513)
    {
        const int q0 = qend, s0 = send;
        auto fq = [&](int x) { return Q(q0 + x); };
        auto fs = [&](int x) { return S(s0 + x); };
        int qe = 0, te = 0;
        auto emit = [&](uint8_t op) { if (end >= cap) return false; ops[end++] = op; return true; };
        const int r = extend(fq, qsize - qend, fs, ssize - send, sc, emit, blk, &qe, &te);
        if (r < 0) { R.status = r; return R; }
        qend += qe; send += te;
    }
    if (beg == end) return R;
    // overhang step (:517-532, left_extend :214-266, right_extend :268-311)
    if (process_over_hang) {
        {
            int qls = qoff, sls = soff;
            const int ls = qls < sls ? qls : sls;
            if (ls > 0 && ls <= kMaxOverHang) {
                R.overhang_left = 1;
                qls += kMatLen; sls += kMatLen;
                auto fq = [&](int x) { return Q(qls - 1 - x); };
                auto fs = [&](int x) { return S(sls - 1 - x); };
                // the new columns replace the first kMatLen ones; they are written backwards in front of them
                uint8_t* tmp = ops + end;                              // scratch room behind the alignment
                const int r = overhang_extend(fq, qls, fs, sls, sc, tmp, cap - end);
                if (r < 0) { R.status = r; return R; }
                if (r > 0) {
                    beg += kMatLen; qoff += kMatLen; soff += kMatLen;
                    if ((size_t)r > beg) { R.status = kOpsOverflow; return R; }
                    int qi = 0, si = 0;
                    for (int x = 0; x < r; ++x) {
                        const uint8_t op = tmp[x];
                        ops[--beg] = op;
                        if (op != kOpDel) ++qi;
                        if (op != kOpIns) ++si;
                    }
                    qoff -= qi; soff -= si;
                }
            }
        }
        {
            int qrs = qsize - qend, srs = ssize - send;
            const int rs = qrs < srs ? qrs : srs;
            if (rs > 0 && rs <= kMaxOverHang) {
                R.overhang_right = 1;
                qrs += kMatLen; srs += kMatLen;
                const int q0 = qsize - qrs, s0 = ssize - srs;
                auto fq = [&](int x) { return Q(q0 + x); };
                auto fs = [&](int x) { return S(s0 + x); };
                if (end < (size_t)kMatLen + beg) { R.status = kNoAlign; return R; }
                uint8_t* dst = ops + end - kMatLen;
                // the first kMatLen columns of the result re-align the last kMatLen ones: compute aside, then copy
                uint8_t* tmp = ops + end;
                const int r = overhang_extend(fq, qrs, fs, srs, sc, tmp, cap - end);
                if (r < 0) { R.status = r; return R; }
                if (r > 0) {
                    qend -= kMatLen; send -= kMatLen; end -= kMatLen;
                    int qi = 0, si = 0;
                    for (int x = 0; x < r; ++x) {
                        const uint8_t op = tmp[x];
                        dst[x] = op;
                        if (op != kOpDel) ++qi;
                        if (op != kOpIns) ++si;
                    }
                    end += (size_t)r;
                    qend += qi; send += si;
                }
            }
        }
    }
    R.qoff = qoff; R.qend = qend; R.soff = soff; R.send = send;
    R.aln_len = (int)(end - beg); R.ops_beg = beg;
    R.status = kOk;
    return R;
}

}  // namespace lesvtb
