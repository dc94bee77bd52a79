// ksw_core.cuh -- per-thread arithmetic core of the banded dual-affine DP kernel (K-ksw).
//
// Replaces: ksw_extd2_sse (reference src/algo/ksw2_extd2_sse.c:26-393, flag = 0, zdrop = -1)
// with the scoring fixed by ksw2_extd2_set_params (src/algo/ksw2_wrapper.c:72-95):
// match +2, mismatch -5, gap = min(5 + 4l, 56 + 1l).
//
// The file is written once for device and host: on the device the packed helpers map to the
// sm_100a SIMD-in-register instructions (VIADD.16x2, VIMNMX3.S16x2, VIADDMNMX.S16x2, PRMT,
// SHF); on the host they are scalar restatements so that tests/emul can run the *same* update
// code single-threaded on a CPU and compare it with the oracle.  The host build is a test
// vehicle only -- the product never executes it.
//
// ---------------------------------------------------------------------------------------------
// Formulation.  The reference keeps, per target column t, the int8 differences u,v,x,y,x2,y2
// and per anti-diagonal r computes for every column of the 16-aligned band
//     z = max(s, a = x[t-1]+v[t-1], b = y[t]+u[t], a2 = x2[t-1]+v[t-1], b2 = y2[t]+u[t])
// (first maximum wins, order s,a,b,a2,b2), clamps z to +2, and updates
//     u' = z - v[t-1]   v' = z - u[t]
//     x' = max(a - z + q, 0) - (q+e)     (likewise y', and x2'/y2' with q2,e2)
// storing one traceback byte: winner index | 4 "gap is being extended" flags.  All of it is
// 8-bit two's-complement arithmetic WITH wrap-around, and the wrap-around is observable: columns
// of the 16-aligned band that lie outside the band proper are computed from band-edge garbage,
// overflow int8 (measured on the oracle, tests/test_oracle_ksw.py), and feed back into the
// band-edge cells of later rows.
//
// Here one 32-bit register holds TWO columns in signed 16-bit lanes, each lane = 256*value + tag:
// the int8 value sits in the HIGH byte, so VIADD.16x2's mod-2^16 wrap IS the int8 wrap and
// VIMNMX's signed 16-bit compare IS the int8 compare with the low byte breaking ties.  Pair j of a
// 16-column block holds columns (j, j+8), so "one column to the left" of pair j is pair j-1 (one
// PRMT per block fixes up pair 0).
//
// sm_100a issues the packed instructions on two pipes of half a warp-instruction per clock each
// (tools/ubench/pipes.cu, profiles/r2_pipes.txt): VIADD.16x2 and IMAD on one, VIMNMX / VIMNMX3 /
// VIADDMNMX / LOP3 / PRMT / SHF on the other.  The representation below is chosen so that the
// second pipe sees as few instructions as possible and both pipes carry about the same load.
// State per column (lane), all arithmetic mod 2^16:
//     UC = 256(u - (q+e))   VC = 256(v - (q+e))                      (low byte 0)
//     PX = 256 max(a',0) + 4    with x = PX/256 - (q+e);  PY = .. + 3,  PX2 = .. + 2,  PY2 = .. + 1
//     S  = 256 s + 5                                                  (tag = 5 - winner index)
// so that the five candidates of a row need no constants:
//     ka  = PX[t-1] + VC[t-1]        = 256a  + 4        kb  = PY  + UC        = 256b  + 3
//     ka2 = PX2[t-1] + (VC[t-1] + D) = 256a2 + 2        kb2 = PY2 + (UC + D)  = 256b2 + 1
//     (D = 256((q+e) - (q2+e2)))
//     k   = max3(max3(S, ka, kb), ka2, kb2)             tag = k & 7
//     W   = min((k | 0xFF) + K, C)   = 256(z - 2(q+e)) + 1            (one VIADDMNMX: clamp z <= 2)
//     UC' = W + ~VC[t-1]    VC' = W + ~UC                              (~w = -w-1 per lane)
//     N   = ~W + 256(q - 2(q+e)) + 2 = 256(q - z),   N2 = N + 256(q2 - q)
//     PX' = max(ka + N, 4)   PY' = max(kb + N, 3)   PX2' = max(ka2 + N2, 2)   PY2' = max(kb2 + N2, 1)
// The relu thresholds are the tags, so a clamped and an unclamped lane carry the same low byte and
// no lane sum ever carries out of its low byte into the value.  The flag "x' > -(q+e)" is
// min(PX', 256 + 4) = 256 flag + 4; the four flags and the tag are summed into the high byte of the
// lane with multiply-adds (the other pipe) and one PRMT per four columns gathers the bytes.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define KSW_HD __host__ __device__ __forceinline__
#define KSW_HDC __host__ __device__ constexpr
#else
#define KSW_HD inline
#define KSW_HDC constexpr
#endif

// kWK as a constant-bank operand: VIADDMNMX takes one immediate, and ptxas rebuilds a second constant with a MOV per use
// (8 per block row) rather than keeping it in a register; a __constant__ it cannot fold is read straight from the bank.
#define KSW_WK_VALUE 0xED02ED02u
#if defined(__CUDA_ARCH__) || defined(__CUDACC__)
static __constant__ uint32_t ksw_wk_bank = KSW_WK_VALUE;
#endif
#if defined(__CUDA_ARCH__)
#define KSW_WK ksw_wk_bank
#else
#define KSW_WK KSW_WK_VALUE
#endif

namespace lesvksw {

// ---- scoring in force (ksw2_wrapper.c:83-94) and what ksw_extd2_sse derives from it ----------
constexpr int kMatch = 2, kMismatch = -5;
constexpr int kQ = 5, kE = 4, kQ2 = 56, kE2 = 1;           // q+e <= q2+e2, no swap (:70)
constexpr int kLongThres = 16;                              // (q2-q)/(e-e2)-1, not bumped (:94-96)
constexpr int kLongDiff = kLongThres * (kE - kE2) - (kQ2 - kQ) - kE2;   // -4 (:97)
static_assert(kLongDiff == -4, "long_diff");
static_assert(!(kQ2 + kE2 + kLongThres * kE2 > kQ + kE + kLongThres * kE), "long_thres bump");

// ---- lane representation: 256*value + tag ---------------------------------------------------
constexpr int kQE = kQ + kE, kQE2 = kQ2 + kE2;
constexpr int kTagS = 5, kTagA = 4, kTagB = 3, kTagA2 = 2, kTagB2 = 1;     // none is 0: ptxas builds a zero operand of VIADDMNMX with a PRMT per use
KSW_HDC uint32_t dup(int v) { return (uint32_t)(v & 0xFFFF) | ((uint32_t)(v & 0xFFFF) << 16); }
constexpr uint32_t kD2 = dup(256 * (kQE - kQE2));           // VC/UC -> the long-gap candidates
constexpr uint32_t kWK = dup(256 * (-2 * kQE) - 254);       // (k | 0xFF) + kWK = 256(z - 2(q+e)) + 1
constexpr uint32_t kWClamp = dup(256 * (kMatch - 2 * kQE) + 1);
constexpr uint32_t kNQ = dup(256 * (kQ - 2 * kQE) + 2);     // ~W + kNQ = 256(q - z)
constexpr uint32_t kNQ2 = dup(256 * (kQ2 - kQ));            // N -> N2
constexpr uint32_t kReluA = dup(kTagA), kReluB = dup(kTagB), kReluA2 = dup(kTagA2), kReluB2 = dup(kTagB2);
constexpr uint32_t kInitUV = dup(256 * (-2 * kQE));         // u = v = -(q+e)
constexpr uint32_t kInitS = dup(256 * 0 + kTagS);           // calloc()ed s[] = 0
constexpr uint32_t kSMatch = dup(256 * kMatch + kTagS);
constexpr uint32_t kSStep = (uint32_t)((256 * (kMismatch - kMatch)) & 0xFFFF);   // 0xF900, no carry out of a lane
static_assert(256 * kMatch + kTagS + (int)kSStep < 65536, "score select stays inside its lane");
static_assert((kSStep & 0xFF) == 0, "score step is a multiple of 256: it can be pre-shifted by up to 8 bits");
// flags: min(P, 256 + tag) = 256 flag + tag; summed with weights 128/64/32/16 next to 256 * tag
constexpr uint32_t kSatA = dup(256 + kTagA), kSatB = dup(256 + kTagB), kSatA2 = dup(256 + kTagA2), kSatB2 = dup(256 + kTagB2);
constexpr int kFlagJunk = 128 * kTagA + 64 * kTagB + 32 * kTagA2 + 16 * kTagB2;   // what the tags add up to
constexpr uint32_t kFlagFix = (uint32_t)(0u - (uint32_t)kFlagJunk * 0x00010001u); // 32-bit add: both lanes at once
static_assert(kFlagJunk == 784 && kTagS < 8, "flag sum");

// value the reference writes into u[r] (top row) and uses as v1 at column 0 (:150, :154)
KSW_HD int first_gap_diff(int r)
{
    return r == 0 ? -(kQ + kE) : r < kLongThres ? -kE : r == kLongThres ? kLongDiff : -kE2;
}

// ---- packed 16x2 helpers -----------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
KSW_HD uint32_t vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
KSW_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
KSW_HD uint32_t vmin2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
KSW_HD uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }   // max(a+b, c)
KSW_HD uint32_t vaddmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2(a, b, c); }   // min(a+b, c)
KSW_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t r;                              // PTX prmt: selector bit 3 = replicate the byte's sign
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}
KSW_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
#else
KSW_HD int16_t lo16(uint32_t a) { return (int16_t)(a & 0xFFFF); }
KSW_HD int16_t hi16(uint32_t a) { return (int16_t)(a >> 16); }
KSW_HD uint32_t mk2(int lo, int hi) { return (uint32_t)(lo & 0xFFFF) | ((uint32_t)(hi & 0xFFFF) << 16); }
KSW_HD uint32_t vadd2(uint32_t a, uint32_t b) { return mk2(lo16(a) + lo16(b), hi16(a) + hi16(b)); }
KSW_HD int max_i(int a, int b) { return a > b ? a : b; }
KSW_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c)
{
    return mk2(max_i(max_i(lo16(a), lo16(b)), lo16(c)), max_i(max_i(hi16(a), hi16(b)), hi16(c)));
}
KSW_HD uint32_t vmin2(uint32_t a, uint32_t b)
{
    return mk2(lo16(a) < lo16(b) ? lo16(a) : lo16(b), hi16(a) < hi16(b) ? hi16(a) : hi16(b));
}
KSW_HD uint32_t vaddmax(uint32_t a, uint32_t b, uint32_t c)
{
    return mk2(max_i((int16_t)(lo16(a) + lo16(b)), lo16(c)), max_i((int16_t)(hi16(a) + hi16(b)), hi16(c)));
}
KSW_HD int min_i(int a, int b) { return a < b ? a : b; }
KSW_HD uint32_t vaddmin(uint32_t a, uint32_t b, uint32_t c)
{
    return mk2(min_i((int16_t)(lo16(a) + lo16(b)), lo16(c)), min_i((int16_t)(hi16(a) + hi16(b)), hi16(c)));
}
KSW_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint64_t ab = (uint64_t)b << 32 | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)((ab >> (8 * (sel & 7))) & 0xFF);
        if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
KSW_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
#endif

static_assert(kWK == KSW_WK_VALUE, "ksw_wk_bank holds kWK");

// ---- internal traceback byte -------------------------------------------------------------------
// bits 0-2: 5 - winner (winner 0 H, 1 E, 2 F, 3 E~, 4 F~); bit 7: E continues; bit 6: F continues;
// bit 5: E~ continues; bit 4: F~ continues; bit 3: 0.  (The reference's own layout, ksw2.h:117-120,
// is winner | flags at bits 3..6; only our backtrack kernel reads this one.)
KSW_HD int tb_cont_bit(int state) { return 8 - state; }
KSW_HD int tb_winner(uint32_t byte) { return kTagS - (int)(byte & 7); }
// byte position of column c (0..15) inside the 16-byte block written per thread per row
KSW_HD int tb_byte_pos(int c) { return 4 * ((c & 7) >> 1) + 2 * (c & 1) + (c >> 3); }

// ---- per-block register state -----------------------------------------------------------------
struct BlockState {
    uint32_t U[8], V[8], X[8], Y[8], X2[8], Y2[8];          // UC, VC, PX, PY, PX2, PY2 of the header comment
    // The score byte s[t] of a column is one of three values -- match, mismatch, or the calloc()ed 0
    // of a column never written yet -- so the block keeps 2 bits per column (bit 2c: mismatch, bit
    // 2c+1: never written) and rebuilds the S lanes of a row from them.  Stale bytes (columns of the
    // 16-aligned band outside [st0, wr]) are simply bits that were not refreshed.
    uint32_t sbits;
    KSW_HD void init()
    {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            U[j] = kInitUV; V[j] = kInitUV; X[j] = kReluA; Y[j] = kReluB;
            X2[j] = kReluA2; Y2[j] = kReluB2;
        }
        sbits = 0xAAAAAAAAu;
    }
};

// Left-neighbour values entering a block at column -1: PX, VC (low/high lane of xv) and PX2.
struct Carry {
    uint32_t xv;    // lo = PX[-1], hi = VC[-1]
    uint32_t x2;    // lo = PX2[-1]
};
KSW_HD Carry carry_init()
{
    Carry c; c.xv = (kInitUV & 0xFFFF0000u) | (uint32_t)kTagA; c.x2 = (uint32_t)kTagA2; return c;
}
KSW_HD Carry carry_left_boundary(int r)     // column 0, :148-151
{
    Carry c;
    c.xv = ((uint32_t)((256 * (first_gap_diff(r) - kQE)) & 0xFFFF) << 16) | (uint32_t)kTagA;
    c.x2 = (uint32_t)kTagA2;
    return c;
}


// mismatch bits of a block: bit 2c set iff target column c differs from the query base under it
KSW_HD uint32_t mismatch_bits(uint32_t tword, uint32_t qword)
{
    uint32_t x = tword ^ qword;
    return (x | (x >> 1)) & 0x55555555u;
}

// S lanes of a row from the block's score bits: one AND and one multiply-add per pair (the bit of pair
// j sits at 2j / 2j+16; kSStep is a multiple of 256, so for 2j <= 8 the shift is folded into the factor).
KSW_HD void score_lanes(uint32_t sbits, bool all_written, uint32_t S[8])
{
    const uint32_t hi = sbits >> 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int sh = j < 5 ? 2 * j : 2 * j - 8;
        const uint32_t e = (j < 5 ? sbits : hi) & (0x00010001u << sh);
        S[j] = e * (kSStep >> sh) + kSMatch;                  // per lane: match or mismatch
    }
    if (!all_written && (sbits & 0xAAAAAAAAu)) {              // leading-edge block only: columns still at calloc's 0
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t z = (sbits >> (2 * j + 1)) & 0x00010001u;
            S[j] -= z * (uint32_t)((kSMatch - kInitS) & 0xFFFF);    // match score -> 0, no borrow between lanes
        }
    }
}

// One pair of columns, one anti-diagonal: x1, v1, x21 are the previous-row PX, VC, PX2 of the columns to the
// left.  Updates pair j of the block in place and returns the lane word whose HIGH bytes are the traceback bytes.
KSW_HD uint32_t pair_row(BlockState& b, int j, uint32_t s, uint32_t x1, uint32_t v1, uint32_t x21)
{
    const uint32_t u0 = b.U[j];
    const uint32_t ka = vadd2(x1, v1), kb = vadd2(b.Y[j], u0);
    const uint32_t ka2 = vadd2(x21, vadd2(v1, kD2)), kb2 = vadd2(b.Y2[j], vadd2(u0, kD2));
    const uint32_t k = vmax3(vmax3(s, ka, kb), ka2, kb2);
    const uint32_t w = vaddmin(k | 0x00FF00FFu, KSW_WK, kWClamp);
    const uint32_t n = vadd2(~w, kNQ), n2 = vadd2(n, kNQ2);
    b.U[j] = vadd2(w, ~v1);
    b.V[j] = vadd2(w, ~u0);
    const uint32_t px = vaddmax(ka, n, kReluA), py = vaddmax(kb, n, kReluB);
    const uint32_t px2 = vaddmax(ka2, n2, kReluA2), py2 = vaddmax(kb2, n2, kReluB2);
    b.X[j] = px; b.Y[j] = py; b.X2[j] = px2; b.Y2[j] = py2;
    // high byte of a lane: E continues | F continues | E~ continues | F~ continues | 0 | tag (k's low byte; the
    // value byte of k lands in the neighbouring lane's low byte, which nobody reads)
    uint32_t t = k * 256u + kFlagFix;
    t += vmin2(py2, kSatB2) * 16u;
    t += vmin2(px2, kSatA2) * 32u;
    t += vmin2(py, kSatB) * 64u;
    t += vmin2(px, kSatA) * 128u;
    return t;
}

// One anti-diagonal step of a 16-column block.  `carry()` returns the previous-row values of the
// column left of the block; it is called as late as possible (just before pair 0, the last pair
// processed) so that a neighbour that is still finishing its previous row is not waited for.  On
// return `cout` holds this row's values of column 15 and tb[0..3] the 16 traceback bytes (column
// order given by tb_byte_pos).
template <class CarryFn>
KSW_HD void block_row(BlockState& b, const uint32_t S[8], CarryFn&& carry, Carry& cout, uint32_t tb[4])
{
    // previous-row values of column 7 = left neighbours of column 8 (high lane of pair 0)
    const uint32_t x7 = b.X[7], v7 = b.V[7], x27 = b.X2[7];
    uint32_t t[8];
#pragma unroll
    for (int j = 7; j > 0; --j) {           // descending: pair j-1 still holds the previous row
        t[j] = pair_row(b, j, S[j], b.X[j - 1], b.V[j - 1], b.X2[j - 1]);
        if (j == 7) {
            cout.xv = prmt(b.X[7], b.V[7], 0x7632);      // lo = PX[7].hi, hi = VC[7].hi  (column 15)
            cout.x2 = b.X2[7] >> 16;
        }
        if (j & 1) continue;
        tb[j >> 1] = prmt(t[j], t[j + 1], 0x7531);
    }
    const Carry cin = carry();
    const uint32_t xl0 = prmt(cin.xv, x7, 0x5410);        // lo = carried PX (column -1), hi = PX of column 7
    const uint32_t vl0 = prmt(cin.xv, v7, 0x5432);        // lo = carried VC, hi = VC of column 7
    const uint32_t x2l0 = prmt(cin.x2, x27, 0x5410);
    t[0] = pair_row(b, 0, S[0], xl0, vl0, x2l0);
    tb[0] = prmt(t[0], t[1], 0x7531);
}

// Backtrack state machine, one step (ksw_backtrack, ksw2.h:121-153, is_rot = 1, min_intron_len = 0).
// `byte` is the internal traceback byte (ignored when force >= 0).  Returns the new state:
// 0 = diagonal move, 1/3 = deletion (consume target), 2/4 = insertion (consume query).
KSW_HD int tb_next_state(int state, uint32_t byte, int force)
{
    if (force >= 0) return force;
    if (state != 0 && !((byte >> tb_cont_bit(state)) & 1)) state = 0;
    if (state == 0) state = tb_winner(byte);
    return state;
}

// ---- band geometry (ksw2_extd2_sse.c:124-139) -------------------------------------------------
struct RowGeom {
    int st0, en0;      // band proper
    int wr;            // last column whose score byte is rewritten this row (:158)
    KSW_HD bool empty() const { return st0 > en0; }
};
KSW_HD RowGeom row_geom(int r, int qlen, int tlen, int w)
{
    RowGeom g;
    int st = 0, en = tlen - 1;
    if (st < r - qlen + 1) st = r - qlen + 1;
    if (en > r) en = r;
    if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
    if (en > ((r + w) >> 1)) en = (r + w) >> 1;
    g.st0 = st; g.en0 = en;
    g.wr = st + (((en - st) >> 4) + 1) * 16 - 1;
    return g;
}
// blocks per traceback row = n_col_ of the reference (:85-86)
KSW_HD int tb_row_blocks(int qlen, int tlen, int w)
{
    int n = qlen < tlen ? qlen : tlen;
    n = n < w + 1 ? n : w + 1;
    return (n + 15) / 16 + 1;
}

// Traceback of a full-matrix alignment (column strips): without a band every column t is live in exactly the qlen rows
// t .. t+qlen-1, so rows r and r +- qlen never share a column and the row-major layout of the banded kernels (one slot per
// anti-diagonal, sized for the widest one) would leave half of the arena empty.  The strips fold the rows modulo qlen: row
// r lies in physical row r % qlen, block B at slot B + r / qlen -- the shift keeps the 16-aligned blocks of two passes,
// which may share a boundary block, apart.
KSW_HD int fold_passes(int qlen, int tlen) { return (qlen + tlen - 2) / qlen + 1; }
KSW_HD size_t fold_row_bytes(int qlen, int tlen) { return ((size_t)((tlen - 1) >> 4) + 1 + (size_t)fold_passes(qlen, tlen)) * 16; }
KSW_HD size_t fold_tb_bytes(int qlen, int tlen) { return (size_t)qlen * fold_row_bytes(qlen, tlen); }

// What a block at an edge of the band has to know about row r.  The numbers are the same for every thread, so a
// kernel may compute them once per row for the whole CTA (a table in shared memory, filled a chunk of rows ahead)
// and hand thread_row_step a lookup instead of the arithmetic.
struct RowEdges {
    int st0;        // first column of the band proper; its block st0 >> 4 is the left-most block of the row
    int wr;         // last column whose score byte is rewritten; wr >> 4 is the right-most block that exists
    int ben;        // en0 >> 4: blocks beyond it only keep score bytes (pre-blocks)
    int bst_prev;   // left-most block of row r-1 (INT_MAX for r = 0): block B-1 was computed there iff B-1 >= bst_prev
};
KSW_HD RowEdges row_edges(int r, int qlen, int tlen, int w)
{
    const RowGeom g = row_geom(r, qlen, tlen, w);
    RowEdges e;
    e.st0 = g.st0; e.wr = g.wr; e.ben = g.en0 >> 4;
    e.bst_prev = r > 0 ? row_geom(r - 1, qlen, tlen, w).st0 >> 4 : 0x7fffffff;
    return e;
}
struct ComputeEdges {
    int qlen, tlen, w;
    KSW_HD RowEdges operator()(int r) const { return row_edges(r, qlen, tlen, w); }
};

// ---- the per-thread row step shared by the kernel and the host emulation ----------------------
// A thread (slot) owns, at any row, the unique block B == slot (mod nt) of the live window.
struct ThreadCtx {
    BlockState b;
    int myB;               // block currently held (-1 = none)
    int cand;              // next block this thread can own: == slot (mod nt), >= every block it has held
    uint32_t tword;        // 16 target bases of block myB, 2 bits each, column c at bits 2c
    uint32_t qlo, qhi;     // window into the packed reversed query: words qwi, qwi+1 of qrp
    int qwi;
    int fast_lo, fast_n;   // fast_n rows from fast_lo on, block myB lies strictly inside the band (fast path)
    int tb_off;            // byte offset of block myB inside a traceback row: 16 * (myB mod tb_blocks)
};

KSW_HD void thread_ctx_reset(ThreadCtx& c, int slot)
{
    c.myB = -1; c.cand = slot; c.qwi = -0x7fffffff; c.tword = 0; c.qlo = c.qhi = 0;
    c.fast_lo = 0; c.fast_n = 0; c.tb_off = 0;
    c.b.init();
}
KSW_HD bool row_is_fast(const ThreadCtx& c, int r) { return (unsigned)(r - c.fast_lo) < (unsigned)c.fast_n; }

// Where block B of a row is stored: the live window never spans more than tb_blocks blocks
// (ksw2_extd2_sse.c:182 asserts the same), so B mod tb_blocks is collision-free within a row.
KSW_HD int tb_block_slot(int B, int tb_blocks) { return B % tb_blocks; }

// 16 bases of the reversed query under block B in row r: the index under column 0 is qlen-1-r+16B, it slides down
// by one per row; the shift is the same for every block of the row.
KSW_HD uint32_t query_window(ThreadCtx& c, const uint32_t* qrp, int qlen, int r, int B)
{
    const int qrel = qlen - 1 - r;
    const int wi = (qrel >> 4) + 1 + B;     // arithmetic shift = floor; +1 skips the pad word
    if (wi != c.qwi) {
        if (wi == c.qwi - 1) { c.qhi = c.qlo; c.qlo = qrp[wi]; }
        else { c.qlo = qrp[wi]; c.qhi = qrp[wi + 1]; }
        c.qwi = wi;
    }
    return funnel_r(c.qlo, c.qhi, 2 * (qrel & 15));
}

#if defined(__CUDA_ARCH__)
KSW_HD bool warp_all(bool p) { return __all_sync(0xFFFFFFFFu, p); }
#else
// Host emulation (tests/emul): one thread is its own warp.  With the switch off every row goes through the band-edge
// path, which is what a block strictly inside the band executes on the GPU whenever another lane of its warp is not.
static bool g_ksw_emul_lean = true;
KSW_HD bool warp_all(bool p) { return p && g_ksw_emul_lean; }
#endif

// tp: packed target words (block B -> tp[B]); qrp: packed reversed query, word index offset by
// +1 (qrp[0] is a zero pad word so that index -1..-16 reads zeros).  `carry()` fetches what the
// owner of block B-1 published for row r-1 (called at most once, and only if needed).
// Returns true if the thread computed DP cells this row (then cout/tb are valid and the 16 bytes
// belong at byte offset c.tb_off of traceback row r).
// kEager: the neighbour's value is already at hand (the previous row ended with a CTA barrier), so
// `carry()` is evaluated up front and the band-edge cases overwrite it inside the slow path; the
// block update then sees one plain value and no per-row policy select.
// kVote: the function is entered by whole warps.  A warp whose 32 blocks all lie strictly inside the band -- all
// but the one or two warps at the band edges -- then skips everything that concerns the edges with ONE uniform
// branch; the edge warps are the critical path of a row (every warp of the CTA waits for them at the row barrier),
// so their extra work is kept short as well: geometry once per warp, the rare cases behind their own branches.
template <bool kEager = false, bool kVote = true, class CarryFn, class EdgeFn>
KSW_HD bool thread_row_step(ThreadCtx& c, int nt, int r, int qlen, int tlen, int w, int tb_blocks,
                            const uint32_t* tp, const uint32_t* qrp,
                            CarryFn&& carry, Carry& cout, uint32_t tb[4], EdgeFn&& edges)
{
    Carry eager = carry_init();
    if (kEager) eager = carry();
    const bool fast = row_is_fast(c, r);
    // carry policy of this row: 0 = take what the left neighbour published, 1 = matrix edge
    // (column 0), 2 = the left block was not computed one row earlier (initial values)
    int policy = 0;
    uint32_t S[8];
    if (kVote ? warp_all(fast) : fast) {
        // every column of the block is rewritten this row and the block was held one row ago: no geometry, no score-bit
        // bookkeeping, and the query window moves by one base (a new word every 16 rows)
        const int qrel = qlen - 1 - r;
        if ((qrel & 15) == 15) { c.qhi = c.qlo; c.qwi = (qrel >> 4) + 1 + c.myB; c.qlo = qrp[c.qwi]; }
        const uint32_t x = c.tword ^ funnel_r(c.qlo, c.qhi, 2 * qrel);     // the shift amount is taken mod 32
        c.b.sbits = (x | (x >> 1)) & 0x55555555u;                           // every column fresh: the row's mismatch bits
        score_lanes(c.b.sbits, true, S);
    } else {
        uint32_t fm = 0xFFFFFFFFu;                 // score bits refreshed this row (2 per column)
        bool pre_block = false;
        if (!fast) {
            const RowEdges g = edges(r);
            const int bst = g.st0 >> 4, bwr = g.wr >> 4;
            // the band left this thread's block behind: take the next one.  One step is the rule (the band moves by at most one
            // block per row); the loop behind it serves the kernels that visit a slot only now and then, and the empty asm
            // keeps the compiler from turning it into a division inside every row.
            if (c.cand < bst) {
                c.cand += nt;
                while (c.cand < bst) {
#if defined(__CUDA_ARCH__)
                    asm volatile("" : "+r"(c.cand));
#endif
                    c.cand += nt;
                }
            }
            const int B = c.cand;
            if (B > bwr) return false;
            if (B != c.myB) {                   // take ownership of a new block
                c.myB = B;
                c.b.init();
                c.tword = tp[B];
                c.qwi = -0x7fffffff;
                c.tb_off = 16 * tb_block_slot(B, tb_blocks);
                // rows with st0 <= 16B-1 (so block B-1 was live one row earlier) and en0 >= 16B+15:
                //   en0 = min(tlen-1, r, (r+w)>>1) >= c15  <=>  r >= max(c15, 2*c15 - w)   (c15 <= tlen-1);
                //   row c15 itself still gets the top-row reset of column c15 below, so start one row later
                //   st0 = max(0, r-qlen+1, (r-w+1)>>1) <= c0-1  <=>  r <= min(c0+qlen-2, 2*(c0-1)+w)
                const int c0 = 16 * B, c15 = c0 + 15;
                c.fast_lo = 0; c.fast_n = 0;
                if (B > 0 && c15 <= tlen - 1) {
                    const int lo = c15 + 1 > 2 * c15 - w ? c15 + 1 : 2 * c15 - w;
                    const int h1 = c0 + qlen - 2, h2 = 2 * (c0 - 1) + w;
                    const int hi = h1 < h2 ? h1 : h2;
                    c.fast_lo = lo; c.fast_n = hi >= lo ? hi - lo + 1 : 0;
                }
            }
            // fresh columns st0 <= 16B + col <= wr: only the left-most and the right-most block of the row are cut
            // (bst <= B <= bwr here)
            if (B == bst) fm = 0xFFFFFFFFu << (2 * (g.st0 & 15));
            if (B == bwr) fm &= 0xFFFFFFFFu >> (30 - 2 * (g.wr & 15));
            pre_block = B > g.ben;              // pre-block: only its score bytes exist so far
            if (!pre_block) {
                if ((r >> 4) == B) {            // top row of the matrix reaches column r: :152-155
                    const int col = r & 15, j = col & 7;
                    const uint32_t m = col < 8 ? 0x0000FFFFu : 0xFFFF0000u;
                    const uint32_t uv = dup(256 * (first_gap_diff(r) - kQE));
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj)
                        if (jj == j) {
                            c.b.U[jj] = (c.b.U[jj] & ~m) | (uv & m);
                            c.b.Y[jj] = (c.b.Y[jj] & ~m) | (kReluB & m);      // y = -(q+e), y2 = -(q2+e2)
                            c.b.Y2[jj] = (c.b.Y2[jj] & ~m) | (kReluB2 & m);
                        }
                }
                // block B-1 was computed in row r-1 iff it was inside that row's aligned band (:142-147)
                if (B == 0) policy = 1;
                else if (B - 1 < g.bst_prev) policy = 2;
                if (kEager && policy) eager = policy == 1 ? carry_left_boundary(r) : carry_init();
            }
        }
        const uint32_t qword = query_window(c, qrp, qlen, r, c.myB);
        const uint32_t x = c.tword ^ qword;
        // mismatch of column c at bit 2c, fresh columns only (a block strictly inside the band has fm = all ones)
        const uint32_t sb = c.b.sbits = ((x | (x >> 1)) & 0x55555555u & fm) | (c.b.sbits & ~fm);
        if (pre_block) return false;
        score_lanes(sb, false, S);
    }
    if (kEager) block_row(c.b, S, [&]() -> Carry { return eager; }, cout, tb);
    else
        block_row(c.b, S, [&]() -> Carry {
            if (policy == 0) return carry();
            return policy == 1 ? carry_left_boundary(r) : carry_init();
        }, cout, tb);
    return true;
}

template <bool kEager = false, bool kVote = true, class CarryFn>
KSW_HD bool thread_row_step(ThreadCtx& c, int nt, int r, int qlen, int tlen, int w, int tb_blocks,
                            const uint32_t* tp, const uint32_t* qrp,
                            CarryFn&& carry, Carry& cout, uint32_t tb[4])
{
    return thread_row_step<kEager, kVote>(c, nt, r, qlen, tlen, w, tb_blocks, tp, qrp, carry, cout, tb, ComputeEdges{qlen, tlen, w});
}

}  // namespace lesvksw
