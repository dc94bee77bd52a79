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
// Not restated yet (the candidate gets status kNeedsHirschberg / the overhang step is skipped and
// flagged): edlib's Hirschberg path, taken when 20 B x blocks x columns + 8 B x columns >= 1 MB
// (edlib.cpp:1175-1177), and the DALIGNER overhang extension (hbn_traceback.c:170-311).
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
    const size_t x = (size_t)(j - 1) * nb + b;
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
            const size_t x = (size_t)j * nb + b;
            const uint64_t Pv = j ? sc.P[x - nb] : ~0ull, Mv = j ? sc.M[x - nb] : 0ull;
            const int prev = j ? sc.score[x - nb] : (b + 1) * 64;
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
        const size_t x = (size_t)(j - 1) * nb + b;
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

// edlib_nw (edlib_wrapper.c:32-111).  Returns ops appended (>= 0) or a negative Status.
template <class FQ, class FT>
TRC_HD inline int nw_path(FQ Q, int n, FT T, int m, Scratch& sc, uint8_t* out, size_t out_cap)
{
    if (n == 0 || m == 0) return 0;                  // :43: "return 1" with both strings empty
    if ((size_t)n + (size_t)m > out_cap) return kOpsOverflow;
    if (needs_hirschberg(n, m)) return kNeedsHirschberg;
    int nb;
    if (!myers_fill(Q, n, T, m, sc, &nb)) return kScratchOverflow;
    return myers_walk(sc, nb, n, m, out);            // tolerance 0.5*max, then max: never fails (:58-72)
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
    if (needs_hirschberg(n, bj)) return kNeedsHirschberg;
    *n_ops = myers_walk(sc, nb, n, bj, out);
    *tend = bj;
    return 1;
}

// ---- MEMs of a chain (init_hit_finder.c:469-505, 255-297) -------------------------------------------
// chain seeds: (w = window offset, r = read offset), word_size bases each, in chain order.
template <class FQ, class FS>
TRC_HD inline int extract_mems(FQ Q, FS S, const int* seed_w, const int* seed_r, int n_seeds, int word_size,
                               Mem* mems, int cap)
{
    int n = 0, i = 0;
    while (i < n_seeds) {
        int j = i + 1;
        while (j < n_seeds && seed_w[j] - seed_w[i] == seed_r[j] - seed_r[i]) ++j;
        const int ml = seed_w[j - 1] + word_size - seed_w[i];
        // s_extract_mem(read, window, read off, window off, ml, word_size)
        int x = 0, ri = seed_r[i], wi = seed_w[i];
        while (x < ml) {
            while (x < ml) { if (Q(ri) == S(wi)) break; ++ri; ++wi; ++x; }
            if (ml - x < word_size) break;
            int y = x;
            while (y < ml) { if (Q(ri) != S(wi)) break; ++ri; ++wi; ++y; }
            if (y - x >= word_size) {
                if (n >= cap) return -1;
                mems[n].len = y - x; mems[n].q = ri - (y - x); mems[n].s = wi - (y - x); ++n;
            }
            x = y;
        }
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
    int overhang_left, overhang_right;   // 1 = the reference would have tried the DALIGNER overhang step here
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
template <class FQ, class FS>
TRC_HD inline Result traceback(FQ Q, int qsize, FS S, int ssize, const Mem* mems, int n_mems, Mem* ts_buf,
                               Scratch& sc, uint8_t* ops, size_t cap, size_t front, uint8_t* blk,
                               int min_align_size, double min_ident_perc)
{
    Result R;
    R.status = kNoAlign; R.qoff = R.qend = R.soff = R.send = 0; R.aln_len = 0; R.n_match = 0; R.ops_beg = 0;
    R.overhang_left = R.overhang_right = 0;
    if (n_mems == 0) return R;
    const int tsc = trace_points(mems, n_mems, ts_buf);
    const Mem* tsa = ts_buf;
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
            const int qa = qfrom, sa = sfrom;
            auto fq = [&](int x) { return Q(qa + x); };
            auto fs = [&](int x) { return S(sa + x); };
            const int r = nw_path(fq, qto - qfrom, fs, sto - sfrom, sc, ops + end, cap - end);
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
    int qoff = tsa[0].q, qend = qto, soff = tsa[0].s, send = sto;

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
    // overhang step (:517-532) not restated: flag where the reference would have run it
    {
        const int ls = qoff < soff ? qoff : soff;
        R.overhang_left = ls > 0 && ls <= kMaxOverHang;
        const int qrs = qsize - qend, srs = ssize - send;
        const int rs = qrs < srs ? qrs : srs;
        R.overhang_right = rs > 0 && rs <= kMaxOverHang;
    }
    // calc_ident_perc (hbn_traceback_aux.c:4-19)
    int n_match = 0;
    {
        int qi = qoff, si = soff;
        for (size_t a = beg; a < end; ++a) {
            const uint8_t op = ops[a];
            if (op == kOpDiag) { n_match += Q(qi) == S(si); ++qi; ++si; }
            else if (op == kOpIns) ++qi;
            else ++si;
        }
    }
    R.qoff = qoff; R.qend = qend; R.soff = soff; R.send = send;
    R.aln_len = (int)(end - beg); R.n_match = n_match; R.ops_beg = beg;
    const double ident = 100.0 * n_match / R.aln_len;
    R.status = (R.aln_len >= min_align_size && ident >= min_ident_perc) ? kOk : kNoAlign;
    return R;
}

}  // namespace lesvtb
