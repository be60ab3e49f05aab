/*
 * jaccard_oracle.c — plain-C CPU restatement of the Jaccard scorer + top-K.  TEST INFRASTRUCTURE ONLY
 * (see oracle/jaccard_oracle.py for who may use it).  Follows /root/reference/retrieval_data_annotation.py:
 *   set(seq_i) & set(seq_j), set(seq_i) | set(seq_j)        :12-13   -> sorted-unique id lists + merge
 *   ratio = len(intersection) / len(union); 0 if empty      :10-14   -> exact integer pair (inter, union)
 *   np.argsort(-row)[:k]  (canonical: stable)                :101     -> insertion by (score desc, index asc),
 *                                                                        scores compared as exact rationals
 *   np.fill_diagonal(m, 0)                                   :172-173 -> zero_diag
 * Build: gcc -O2 -shared -fPIC -o libjoracle.so jaccard_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int cmp_i32(const void* a, const void* b) {
    int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
    return (x > y) - (x < y);
}

/* sort + unique every CSR row; returns new arrays (caller frees) */
static int uniq_rows(const int32_t* pos, const int64_t* off, int64_t n, int32_t** out_pos, int64_t** out_off) {
    int32_t* p = (int32_t*)malloc(sizeof(int32_t) * (size_t)(off[n] > 0 ? off[n] : 1));
    int64_t* o = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
    if (!p || !o) return -1;
    int64_t w = 0;
    o[0] = 0;
    for (int64_t r = 0; r < n; ++r) {
        int64_t len = off[r + 1] - off[r];
        memcpy(p + w, pos + off[r], sizeof(int32_t) * (size_t)len);
        qsort(p + w, (size_t)len, sizeof(int32_t), cmp_i32);
        int64_t u = 0;
        for (int64_t i = 0; i < len; ++i)
            if (i == 0 || p[w + i] != p[w + u - 1]) p[w + u++] = p[w + i];
        w += u;
        o[r + 1] = w;
    }
    *out_pos = p;
    *out_off = o;
    return 0;
}

static int64_t intersect(const int32_t* a, int64_t na, const int32_t* b, int64_t nb) {
    int64_t i = 0, j = 0, c = 0;
    while (i < na && j < nb) {
        if (a[i] < b[j]) ++i;
        else if (a[i] > b[j]) ++j;
        else { ++c; ++i; ++j; }
    }
    return c;
}

/* does (i1/u1, x1) rank before (i2/u2, x2)?  exact: i1*u2 vs i2*u1, then smaller index */
static int better(int64_t i1, int64_t u1, int32_t x1, int64_t i2, int64_t u2, int32_t x2) {
    int64_t l = i1 * u2, r = i2 * u1;
    if (l != r) return l > r;
    return x1 < x2;
}

int joracle_counts(const int32_t* q_pos, const int64_t* q_off, int64_t nq, const int32_t* p_pos, const int64_t* p_off,
                   int64_t np, int32_t* inter, int32_t* uni) {
    int32_t *qp, *pp;
    int64_t *qo, *po;
    if (uniq_rows(q_pos, q_off, nq, &qp, &qo) || uniq_rows(p_pos, p_off, np, &pp, &po)) return -1;
    for (int64_t q = 0; q < nq; ++q) {
        int64_t nqe = qo[q + 1] - qo[q];
        for (int64_t p = 0; p < np; ++p) {
            int64_t npe = po[p + 1] - po[p];
            int64_t c = intersect(qp + qo[q], nqe, pp + po[p], npe);
            inter[q * np + p] = (int32_t)c;
            uni[q * np + p] = (int32_t)(nqe + npe - c);
        }
    }
    free(qp); free(qo); free(pp); free(po);
    return 0;
}

int joracle_topk(const int32_t* q_pos, const int64_t* q_off, int64_t nq, const int32_t* p_pos, const int64_t* p_off,
                 int64_t np, int32_t k, int32_t zero_diag, int64_t query_base, int64_t pool_base, int64_t* top_inter,
                 int64_t* top_union, int32_t* top_idx) {
    int32_t *qp, *pp;
    int64_t *qo, *po;
    if (uniq_rows(q_pos, q_off, nq, &qp, &qo) || uniq_rows(p_pos, p_off, np, &pp, &po)) return -1;
    for (int64_t q = 0; q < nq; ++q) {
        int64_t* ti = top_inter + q * k;
        int64_t* tu = top_union + q * k;
        int32_t* tx = top_idx + q * k;
        int32_t filled = 0;
        int64_t nqe = qo[q + 1] - qo[q];
        for (int64_t p = 0; p < np; ++p) {
            int64_t npe = po[p + 1] - po[p];
            int64_t c = intersect(qp + qo[q], nqe, pp + po[p], npe);
            if (zero_diag && query_base + q == pool_base + p) c = 0;
            int64_t u = nqe + npe - c;
            if (u == 0) u = 1; /* both empty: score 0 */
            int32_t gx = (int32_t)(pool_base + p);
            if (filled == k && !better(c, u, gx, ti[k - 1], tu[k - 1], tx[k - 1])) continue;
            int32_t pos = filled < k ? filled : k - 1;
            while (pos > 0 && better(c, u, gx, ti[pos - 1], tu[pos - 1], tx[pos - 1])) {
                ti[pos] = ti[pos - 1]; tu[pos] = tu[pos - 1]; tx[pos] = tx[pos - 1];
                --pos;
            }
            ti[pos] = c; tu[pos] = u; tx[pos] = gx;
            if (filled < k) ++filled;
        }
    }
    free(qp); free(qo); free(pp); free(po);
    return 0;
}
