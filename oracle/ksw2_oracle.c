/* ksw2_oracle.c -- scalar restatement of the reference's dual-affine banded DP.
 * TEST INFRASTRUCTURE ONLY (see lesv_oracle.h).
 *
 * What is restated, and why it is written the way it is:
 *   reference: src/algo/ksw2_extd2_sse.c:26-393 (ksw_extd2_sse, flag = 0, zdrop = -1, m = 5),
 *              src/algo/ksw2.h:121-153 (ksw_backtrack), src/algo/ksw2_wrapper.c:97-211.
 * The reference works on 16-lane int8 SSE vectors over seven byte arrays u,v,x,y,x2,y2,s that
 * sit back to back in one calloc()ed buffer, followed by the target copy `sf` and the reversed
 * query `qr`.  Three side effects of that layout are observable in the traceback matrix and
 * therefore in the CIGAR, so this restatement keeps the same flat buffer and the same write
 * order, one scalar lane at a time:
 *   (1) every anti-diagonal is computed over the 16-ALIGNED column range [st&~15, en|15], so
 *       cells outside the band proper are computed from whatever the arrays hold and their
 *       results feed band-edge cells of the next rows (:139, :181-266);
 *   (2) the match/mismatch byte s[] is refreshed only on [st0, st0+16k) (:158-172), so aligned
 *       lanes outside that range reuse a byte written by an earlier row (or calloc's zero);
 *   (3) the last 16-byte store of that loop can run past the end of s[] into the first bytes of
 *       sf[] (the target copy), and its loads can run past sf[] into qr[] (:160-171).
 * All arithmetic is 8-bit two's complement with wrap-around, as _mm_add_epi8/_mm_sub_epi8 do.
 */
#include "lesv_oracle.h"
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-0x40000000)   /* KSW_NEG_INF, ksw2.h:6 */

void lesv_oracle_ksw_default_params(lesv_oracle_ksw_params* p)
{
    /* ksw2_extd2_set_params, ksw2_wrapper.c:72-95 */
    p->match = 2; p->mismatch = -5; p->ambi = -1;
    p->q = 5; p->e = 4; p->q2 = 56; p->e2 = 1;
}

static inline int8_t add8(int8_t a, int8_t b) { return (int8_t)((uint8_t)a + (uint8_t)b); }
static inline int8_t sub8(int8_t a, int8_t b) { return (int8_t)((uint8_t)a - (uint8_t)b); }

/* row bounds, ksw2_extd2_sse.c:124-138; returns 0 if the row is empty */
static inline int row_bounds(int r, int qlen, int tlen, int w, int* st0, int* en0)
{
    int st = 0, en = tlen - 1;
    if (st < r - qlen + 1) st = r - qlen + 1;
    if (en > r) en = r;
    if (st < ((r - w + 1) >> 1)) st = (r - w + 1) >> 1;
    if (en > ((r + w) >> 1)) en = (r + w) >> 1;
    *st0 = st; *en0 = en;
    return st <= en;
}

int64_t lesv_oracle_ksw_cells(int qlen, int tlen, int w, int* ok)
{
    int64_t cells = 0;
    if (w < 0) w = tlen > qlen ? tlen : qlen;
    if (ok) *ok = 1;
    for (int r = 0; r < qlen + tlen - 1; ++r) {
        int st0, en0;
        if (!row_bounds(r, qlen, tlen, w, &st0, &en0)) { if (ok) *ok = 0; break; }
        cells += en0 - st0 + 1;
    }
    return cells;
}

static uint32_t* push_cigar(int* n, int* m, uint32_t* cigar, uint32_t op, int len)
{
    /* ksw_push_cigar, ksw2.h:102-113 */
    if (*n == 0 || op != (cigar[*n - 1] & 0xf)) {
        if (*n == *m) {
            *m = *m ? (*m) << 1 : 4;
            cigar = (uint32_t*)realloc(cigar, (size_t)(*m) << 2);
        }
        cigar[(*n)++] = (uint32_t)len << 4 | op;
    } else cigar[*n - 1] += (uint32_t)len << 4;
    return cigar;
}

static int ksw_extd2_impl(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                          const lesv_oracle_ksw_params* prm, int w,
                          uint32_t** cigar_out, int* score_out, uint8_t** tb_out, int* tb_row_bytes);

int lesv_oracle_ksw_extd2(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                          const lesv_oracle_ksw_params* prm, int w,
                          uint32_t** cigar_out, int* score_out)
{
    return ksw_extd2_impl(qlen, query, tlen, target, prm, w, cigar_out, score_out, NULL, NULL);
}

/* Same, but also hands back the whole traceback matrix p[] (reference byte layout, ksw2.h:117-120;
 * row r starts at r * *tb_row_bytes and covers columns off[r] = st0/16*16 onward). Used by the
 * tests to compare every in-matrix cell of the aligned band, not just the cells on the path. */
int lesv_oracle_ksw_extd2_tb(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                             const lesv_oracle_ksw_params* prm, int w,
                             uint32_t** cigar_out, uint8_t** tb_out, int* tb_row_bytes)
{
    return ksw_extd2_impl(qlen, query, tlen, target, prm, w, cigar_out, NULL, tb_out, tb_row_bytes);
}

static int ksw_extd2_impl(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                          const lesv_oracle_ksw_params* prm, int w,
                          uint32_t** cigar_out, int* score_out, uint8_t** tb_out, int* tb_row_bytes)
{
    int8_t q = prm->q, e = prm->e, q2 = prm->q2, e2 = prm->e2;
    const int8_t sc_mch = prm->match, sc_mis = prm->mismatch;
    const int m = 5;
    /* sc_N_: mat[m*m-1] is sc_ambi (!= 0) -> that value (:79) */
    const int8_t sc_N = prm->ambi == 0 ? (int8_t)-e2 : (int8_t)(prm->ambi > 0 ? -prm->ambi : prm->ambi);
    *cigar_out = NULL;
    if (tb_out) *tb_out = NULL;
    if (score_out) *score_out = NEG_INF;
    if (qlen <= 0 || tlen <= 0) return 0;
    if (q2 + e2 < q + e) { int8_t t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }   /* :70 */
    const int qe = q + e;
    const int8_t qe_ = (int8_t)(q + e), qe2_ = (int8_t)(q2 + e2);
    if (w < 0) w = tlen > qlen ? tlen : qlen;                                        /* :82 */
    const int tlen_ = (tlen + 15) / 16, qlen_ = (qlen + 15) / 16;
    int n_col_ = qlen < tlen ? qlen : tlen;
    n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;                      /* :85-86 */
    {   /* :88-92 */
        int min_sc = sc_mis < sc_N ? sc_mis : sc_N;
        if (-min_sc > 2 * (q + e)) return 0;
    }
    int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;                          /* :94-97 */
    if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
    const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;

    /* one flat buffer, same layout as :99-102 */
    const size_t A = (size_t)tlen_ * 16;
    const size_t total = ((size_t)tlen_ * 8 + qlen_ + 1) * 16;
    uint8_t* mem = (uint8_t*)calloc(total + 16, 1);
    int8_t *u = (int8_t*)mem, *v = u + A, *x = v + A, *y = x + A, *x2 = y + A, *y2 = x2 + A;
    int8_t *s = y2 + A;
    uint8_t *sf = (uint8_t*)(s + A), *qr = sf + A;
    memset(u, -q - e, A); memset(v, -q - e, A); memset(x, -q - e, A); memset(y, -q - e, A);
    memset(x2, -q2 - e2, A); memset(y2, -q2 - e2, A);
    int32_t* H = (int32_t*)malloc(A * sizeof(int32_t));
    for (size_t t = 0; t < A; ++t) H[t] = NEG_INF;
    const size_t nrow = (size_t)qlen + tlen - 1;
    const size_t prow = (size_t)n_col_ * 16;
    uint8_t* p = (uint8_t*)calloc(nrow * prow + 16, 1);
    int* off = (int*)malloc(nrow * sizeof(int) * 2);
    int* off_end = off + nrow;
    for (int t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
    memcpy(sf, target, tlen);

    int zdropped = 0, score = NEG_INF, last_st = -1, last_en = -1;
    int32_t ez_max = 0; int ez_max_t = -1, ez_max_q = -1;
    for (int r = 0; r < (int)nrow; ++r) {
        int st0, en0;
        if (!row_bounds(r, qlen, tlen, w, &st0, &en0)) { zdropped = 1; break; }
        const int st = st0 / 16 * 16, en = (en0 + 16) / 16 * 16 - 1;                  /* :139 */
        const uint8_t* qrr = qr + (qlen - 1 - r);
        int8_t x1, x21, v1;
        if (st > 0) {                                                               /* :141-151 */
            if (st - 1 >= last_st && st - 1 <= last_en) { x1 = x[st - 1]; x21 = x2[st - 1]; v1 = v[st - 1]; }
            else { x1 = (int8_t)(-q - e); x21 = (int8_t)(-q2 - e2); v1 = (int8_t)(-q - e); }
        } else {
            x1 = (int8_t)(-q - e); x21 = (int8_t)(-q2 - e2);
            v1 = r == 0 ? (int8_t)(-q - e) : r < long_thres ? (int8_t)-e : r == long_thres ? (int8_t)long_diff : (int8_t)-e2;
        }
        if (en >= r) {                                                              /* :152-155 */
            y[r] = (int8_t)(-q - e); y2[r] = (int8_t)(-q2 - e2);
            u[r] = r == 0 ? (int8_t)(-q - e) : r < long_thres ? (int8_t)-e : r == long_thres ? (int8_t)long_diff : (int8_t)-e2;
        }
        /* scores: 16 bytes at a time from st0, may spill (:158-172) */
        for (int t = st0; t <= en0; t += 16) {
            int8_t tmp[16];
            for (int k = 0; k < 16; ++k) {
                uint8_t a = sf[t + k], b = qrr[t + k];
                int8_t sc = a == b ? sc_mch : sc_mis;
                if (a == (uint8_t)(m - 1) || b == (uint8_t)(m - 1)) sc = sc_N;
                tmp[k] = sc;
            }
            memcpy(s + t, tmp, 16);   /* all 16 loads happen before the store, as with SSE */
        }
        /* core loop, left-aligned gaps (:220-266), lane by lane */
        uint8_t* pr = p + (size_t)r * prow - st;
        off[r] = st; off_end[r] = en;
        int8_t xc = x1, x2c = x21, vc = v1;        /* carries = previous-row values at t-1 */
        for (int t = st; t <= en; ++t) {
            int8_t z = s[t];
            int8_t xt1 = xc, vt1 = vc, x2t1 = x2c;
            xc = x[t]; vc = v[t]; x2c = x2[t];
            int8_t ut = u[t];
            int8_t a = add8(xt1, vt1), b = add8(y[t], ut);
            int8_t a2 = add8(x2t1, vt1), b2 = add8(y2[t], ut);
            uint8_t d = 0;
            if (a > z) { d = 1; z = a; }
            if (b > z) { d = 2; z = b; }
            if (a2 > z) { d = 3; z = a2; }
            if (b2 > z) { d = 4; z = b2; }
            if (sc_mch < z) z = sc_mch;
            u[t] = sub8(z, vt1);
            v[t] = sub8(z, ut);
            int8_t tmp = sub8(z, q);
            a = sub8(a, tmp); b = sub8(b, tmp);
            tmp = sub8(z, q2);
            a2 = sub8(a2, tmp); b2 = sub8(b2, tmp);
            x[t]  = sub8(a  > 0 ? a  : 0, qe_);  if (a  > 0) d |= 0x08;
            y[t]  = sub8(b  > 0 ? b  : 0, qe_);  if (b  > 0) d |= 0x10;
            x2[t] = sub8(a2 > 0 ? a2 : 0, qe2_); if (a2 > 0) d |= 0x20;
            y2[t] = sub8(b2 > 0 ? b2 : 0, qe2_); if (b2 > 0) d |= 0x40;
            pr[t] = d;
        }
        /* exact H tracking and row maximum (:315-358).  zdrop is disabled (-1) on this path, so
         * ksw_apply_zdrop (ksw2.h:162-179) only maintains ez->max / max_t / max_q, which decide
         * where the partial backtrack starts when the band is cut later on. */
        {
            int32_t max_H, max_t;
            if (r > 0) {
                int32_t HH[4], tt[4];
                const int en1 = st0 + (en0 - st0) / 4 * 4;
                int t;
                max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0];
                max_t = en0;
                for (int k = 0; k < 4; ++k) { HH[k] = max_H; tt[k] = max_t; }
                for (t = st0; t < en1; t += 4)          /* four interleaved running maxima */
                    for (int k = 0; k < 4; ++k) {
                        H[t + k] += v[t + k];
                        if (H[t + k] > HH[k]) { HH[k] = H[t + k]; tt[k] = t; }
                    }
                for (int k = 0; k < 4; ++k)
                    if (max_H < HH[k]) { max_H = HH[k]; max_t = tt[k] + k; }
                for (; t < en0; ++t) {
                    H[t] += v[t];
                    if (H[t] > max_H) { max_H = H[t]; max_t = t; }
                }
            } else { H[0] = v[0] - qe; max_H = H[0]; max_t = 0; }
            if (max_H > ez_max) { ez_max = max_H; ez_max_t = max_t; ez_max_q = r - max_t; }
        }
        if (r == qlen + tlen - 2 && en0 == tlen - 1) score = H[tlen - 1];
        last_st = st; last_en = en;
    }
    if (score_out) *score_out = score;

    int n_cigar = 0, m_cigar = 0;
    uint32_t* cigar = NULL;
    if (!zdropped || (ez_max_t >= 0 && ez_max_q >= 0)) {
        /* ksw_backtrack(is_rot=1, is_rev=0, min_intron_len=0), ksw2.h:121-153; a cut band
         * backtracks from the best cell seen (ksw2_extd2_sse.c:388-389) -- nw_ksw2_extd2 then
         * rejects that partial CIGAR (ksw2_wrapper.c:183-187). */
        int i = zdropped ? ez_max_t : tlen - 1, j = zdropped ? ez_max_q : qlen - 1, state = 0;
        while (i >= 0 && j >= 0) {
            int force_state = -1, r = i + j;
            uint32_t tmp;
            if (i < off[r]) force_state = 2;
            if (i > off_end[r]) force_state = 1;
            tmp = force_state < 0 ? p[(size_t)r * prow + i - off[r]] : 0;
            if (state == 0) state = tmp & 7;
            else if (!(tmp >> (state + 2) & 1)) state = 0;
            if (state == 0) state = tmp & 7;
            if (force_state >= 0) state = force_state;
            if (state == 0) { cigar = push_cigar(&n_cigar, &m_cigar, cigar, 0, 1); --i; --j; }
            else if (state == 1 || state == 3) { cigar = push_cigar(&n_cigar, &m_cigar, cigar, 2, 1); --i; }
            else { cigar = push_cigar(&n_cigar, &m_cigar, cigar, 1, 1); --j; }
        }
        if (i >= 0) cigar = push_cigar(&n_cigar, &m_cigar, cigar, 2, i + 1);
        if (j >= 0) cigar = push_cigar(&n_cigar, &m_cigar, cigar, 1, j + 1);
        for (i = 0; i < n_cigar >> 1; ++i) {
            uint32_t t = cigar[i]; cigar[i] = cigar[n_cigar - 1 - i]; cigar[n_cigar - 1 - i] = t;
        }
    }
    free(mem); free(H); free(off);
    if (tb_out) { *tb_out = p; *tb_row_bytes = (int)prow; } else free(p);
    *cigar_out = cigar;
    return n_cigar;
}

int lesv_oracle_nw_ksw2_extd2(const uint8_t* query, int qfrom, int qto,
                              const uint8_t* subject, int sfrom, int sto,
                              double min_ident_perc, int max_distance,
                              int* qoff, int* qend, int* soff, int* send, double* ident,
                              char** qaln_out, char** saln_out, int* aln_size_out)
{
    /* ksw2_wrapper.c:97-211 */
    lesv_oracle_ksw_params prm;
    lesv_oracle_ksw_default_params(&prm);
    const int ql = qto - qfrom, sl = sto - sfrom;
    if (max_distance == 0) max_distance = (int)((ql > sl ? ql : sl) * 0.1);           /* :121 */
    const uint8_t* qs = query + qfrom;
    const uint8_t* ss = subject + sfrom;
    uint32_t* cigar = NULL;
    *qaln_out = *saln_out = NULL; *aln_size_out = 0;
    int n_cigar = lesv_oracle_ksw_extd2(ql, qs, sl, ss, &prm, max_distance, &cigar, NULL);
    if (n_cigar == 0) return 0;
    size_t cap = (size_t)ql + sl + 1;
    char* qa = (char*)malloc(cap), *sa = (char*)malloc(cap);
    int qi = 0, si = 0, n = 0;
    for (int k = 0; k < n_cigar; ++k) {
        int len = cigar[k] >> 4, op = cigar[k] & 0xf;
        for (int t = 0; t < len; ++t) {
            if (op == 0) { qa[n] = "ACGT-"[qs[qi++]]; sa[n] = "ACGT-"[ss[si++]]; }
            else if (op == 1) { qa[n] = "ACGT-"[qs[qi++]]; sa[n] = '-'; }
            else { qa[n] = '-'; sa[n] = "ACGT-"[ss[si++]]; }
            ++n;
        }
    }
    free(cigar);
    qa[n] = sa[n] = 0;
    if (qi < ql || si < sl) { free(qa); free(sa); return 0; }
    *ident = lesv_oracle_calc_ident_perc(qa, sa, n, NULL, NULL);
    if (*ident < min_ident_perc) { free(qa); free(sa); return 0; }
    *qoff = qfrom; *qend = qto; *soff = sfrom; *send = sto;
    *qaln_out = qa; *saln_out = sa; *aln_size_out = n;
    return 1;
}
