/* svsig_oracle.c -- scalar restatement of the alignment scans and of qx2svsig's per-read body.
 * TEST INFRASTRUCTURE ONLY (see lesv_oracle.h).
 *
 * reference: src/algo/hbn_traceback.c:547-606 (truncate_align_bad_ends),
 *            src/algo/hbn_traceback_aux.c:4-95 (calc_ident_perc, calc_effective_ident_perc),
 *            src/app/necat2sv/find_sv_signature.c:126-220 (find_sv_signature), :234-380 (discover_func).
 */
#include "lesv_oracle.h"
#include <stdlib.h>
#include <string.h>

#define K_MAT_LEN 8      /* kMatLen, hbn_traceback.c:8 */
#define GAP '-'

int lesv_oracle_truncate_align_bad_ends(const char* qaln, const char* saln, int aln_size,
                                        int* qoff, int* qend, int* soff, int* send,
                                        int* from, int* to)
{
    int qcnt = 0, scnt = 0, m = 0, i = 0;
    /* first run of 8 identical columns from the left */
    while (i < aln_size) {
        int qc = qaln[i], sc = saln[i];
        ++i;
        if (qc != GAP) ++qcnt;
        if (sc != GAP) ++scnt;
        m = (qc == sc) ? m + 1 : 0;
        if (m == K_MAT_LEN) break;
    }
    if (m < K_MAT_LEN || i == aln_size) return 0;
    *from = i - K_MAT_LEN;
    *qoff += qcnt - K_MAT_LEN;
    *soff += scnt - K_MAT_LEN;
    /* last run of 8 identical columns from the right */
    qcnt = scnt = m = 0;
    i = aln_size;
    while (i > 0) {
        --i;
        int qc = qaln[i], sc = saln[i];
        if (qc != GAP) ++qcnt;
        if (sc != GAP) ++scnt;
        m = (qc == sc) ? m + 1 : 0;
        if (m == K_MAT_LEN) break;
    }
    if (i == 0) return 0;
    *to = i + K_MAT_LEN;
    *qend -= qcnt - K_MAT_LEN;
    *send -= scnt - K_MAT_LEN;
    return 1;
}

double lesv_oracle_calc_ident_perc(const char* qa, const char* sa, int n, int* dist, int* score)
{
    if (n == 0) return 0.0;
    int mat = 0;
    for (int i = 0; i < n; ++i) mat += qa[i] == sa[i];
    if (dist) *dist = n - mat;
    if (score) *score = n * 8 + mat * -1;      /* MATCH_REWARD 8, MISMATCH_PENALTY -1 (hbn_defs.h:76-77) */
    return 100.0 * mat / n;
}

double lesv_oracle_calc_effective_ident_perc(const char* qa, const char* sa, int n)
{
    const int E = 20;
    int eff_len = 0, eff_mat = 0, i = 0;
    while (i < n) {
        int qc = qa[i], tc = sa[i];
        if (qc != GAP && tc != GAP) { eff_mat += qc == tc; ++eff_len; ++i; continue; }
        if (qc == GAP && tc == GAP) { ++i; continue; }
        /* a run of gaps on one side; double-gap columns inside the run are skipped over */
        const char* g = qc == GAP ? qa : sa;
        int j = i + 1;
        while (j < n) {
            if (qa[j] == GAP && sa[j] == GAP) { ++j; continue; }
            if (g[j] != GAP) break;
            ++j;
        }
        if (j - i < E) {
            for (int k = i; k < j; ++k) {
                if (qa[k] == GAP && sa[k] == GAP) continue;
                eff_mat += qa[k] == sa[k];
                ++eff_len;
            }
        }
        i = j;
    }
    if (eff_len == 0) return 0.0;
    return 100.0 * eff_mat / eff_len;
}

int lesv_oracle_find_sv_signature(int qid, int qdir, const char* qas, const char* sas, int aln_size,
                                  int subject_offset, int qb, int qe, int fqb, int fqe, int ql,
                                  int sb, int se, int fsb, int fse, int sl, int min_indel_size,
                                  lesv_oracle_svsig* out, int cap)
{
    int i = 0, qi = qb, si = subject_offset + sb, n_out = 0;
    (void)qe; (void)se;
    while (i < aln_size) {
        if (qas[i] != GAP && sas[i] != GAP) { ++i; ++qi; ++si; continue; }
        const int del = qas[i] == GAP;           /* gap in the read = deletion */
        const char* g = del ? qas : sas;
        int j = i + 1;
        while (j < aln_size && g[j] == GAP) ++j;
        int n = j - i;
        if (n >= min_indel_size && n_out < cap) {
            lesv_oracle_svsig* s = &out[n_out++];
            s->qid = qid; s->qdir = qdir;
            s->fqfrom = fqb; s->fqto = fqe; s->qsize = ql;
            s->fsfrom = fsb; s->fsto = fse; s->ssize = sl;
            if (del) { s->qfrom = qi; s->qto = qi + 1; s->sfrom = si; s->sto = si + n; s->type = 0; }
            else     { s->qfrom = qi; s->qto = qi + n; s->sfrom = si; s->sto = si + 1; s->type = 6; }
        }
        if (del) si += n; else qi += n;
        i = j;
    }
    return n_out;
}

int lesv_oracle_discover_one(const lesv_oracle_svread* svr, const uint8_t* read_fwd, int read_len,
                             const uint8_t* contig, int contig_len, int min_indel_size,
                             lesv_oracle_svsig* out, int cap)
{
    /* find_sv_signature.c:241-374 */
    const int qn = svr->qend - svr->qoff;
    uint8_t* q = (uint8_t*)malloc(qn > 0 ? qn : 1);
    if (svr->qdir == 1) {      /* REV: reverse complement of fwd[qoff,qend) (raw_reads_reader.c:83-124) */
        for (int k = 0; k < qn; ++k) q[k] = 3 - read_fwd[svr->qend - 1 - k];
    } else memcpy(q, read_fwd + svr->qoff, qn);
    const uint8_t* s = contig + svr->soff;
    const int sn = svr->send - svr->soff;
    int qb, qe, sb, se, aln_size = 0;
    double ident;
    char *qaln = NULL, *saln = NULL;
    int r = lesv_oracle_nw_ksw2_extd2(q, 0, qn, s, 0, sn, 0.0, (int)(svr->dist * 1.2),
                                      &qb, &qe, &sb, &se, &ident, &qaln, &saln, &aln_size);
    if (!r) r = lesv_oracle_nw_ksw2_extd2(q, 0, qn, s, 0, sn, 0.0, qn > sn ? qn : sn,
                                          &qb, &qe, &sb, &se, &ident, &qaln, &saln, &aln_size);
    free(q);
    if (!r) return -1;
    int from = 0, to = 0, n_out = -1;
    if (lesv_oracle_truncate_align_bad_ends(qaln, saln, aln_size, &qb, &qe, &sb, &se, &from, &to)) {
        const char* qas = qaln + from;
        const char* sas = saln + from;
        aln_size = to - from;
        int fqb = qb + svr->qoff, fqe = qe + svr->qoff;
        int fsb = sb + svr->soff, fse = se + svr->soff;
        double eff = lesv_oracle_calc_effective_ident_perc(qas, sas, aln_size);
        if (!(eff < 70.0))
            n_out = lesv_oracle_find_sv_signature(svr->qid, svr->qdir, qas, sas, aln_size, svr->soff,
                                                  qb, qe, fqb, fqe, read_len, sb, se, fsb, fse,
                                                  contig_len, min_indel_size, out, cap);
    }
    free(qaln); free(saln);
    return n_out;
}
