/* Plain-C restatement of the reference's CPU hot path (TEST / BASELINE INFRASTRUCTURE ONLY).
 *
 * Used as (a) a second, scalar-loop oracle for the tests and (b) the timed CPU baseline in
 * bench.py (`cpu_baseline`, `--impl reference`): the reference is pure Python + numba and
 * cannot travel to the GPU box (/root/reference does not exist there), so its algorithm is
 * restated here loop for loop.  Parity pinned by tests/test_oracle_golden.py against
 * outputs of the unmodified reference (tests/golden/).
 *
 *   euclid / canberra loops        nabo/_mapping.py:16-26, 29-45   (scalar FP64, ascending k,
 *                                  no FMA contraction: compile with -ffp-contract=off)
 *   full-row sort                  nabo/_mapping.py:135-146        (the reference argsorts the
 *                                  WHOLE row; so does this port - it is the reference's cost)
 *   snn counts                     nabo/_mapping.py:186-198
 *   score accumulation             nabo/_graph.py:643-653
 * The reference is single-threaded (numba without parallel=True).  These functions are
 * single-threaded too; oracle/c_port.py runs them on disjoint target slices from a thread
 * pool (ctypes drops the GIL) for the "reference x P threads" baseline.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double d; int32_t i; } pair_t;

static int cmp_pair(const void* a, const void* b) {
    const pair_t* x = (const pair_t*)a; const pair_t* y = (const pair_t*)b;
    /* NaN last (numpy sort semantics), ties by index */
    int xn = isnan(x->d), yn = isnan(y->d);
    if (xn != yn) return xn - yn;
    if (!xn) { if (x->d < y->d) return -1; if (x->d > y->d) return 1; }
    return (x->i > y->i) - (x->i < y->i);
}

void oracle_euclidean_dist(const double* x, const double* y, double* d, int m, int n, int g) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double td = 0.0;
            for (int k = 0; k < g; ++k) { double t = x[(size_t)i * g + k] - y[(size_t)j * g + k]; td += t * t; }
            d[(size_t)i * n + j] = sqrt(td);
        }
}

void oracle_mod_canberra_dist(const double* x, const double* y, double* d, int m, int n, int g, double f) {
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            double dist = 0.0;
            for (int k = 0; k < g; ++k) {
                double absx = fabs(x[(size_t)i * g + k]);
                double num = fabs(x[(size_t)i * g + k] - y[(size_t)j * g + k]);
                if (num < f * absx) {
                    double absy = fabs(y[(size_t)j * g + k]);
                    double den = (absx + absy + 0.01);
                    dist += num / den;
                } else dist += 1;
            }
            d[(size_t)i * n + j] = dist;
        }
}

/* metric 0 = euclidean, 1 = modified Canberra.  mask may be NULL.  Returns 0. */
int oracle_knn(const double* q, const double* r, int n_query, int n_ref, int g, int k, int metric, double f,
               const uint8_t* mask, int drop_first, int32_t* out_idx, double* out_dist) {
    {
        double* row = (double*)malloc(sizeof(double) * (size_t)n_ref);
        pair_t* pr = (pair_t*)malloc(sizeof(pair_t) * (size_t)n_ref);
        for (int t = 0; t < n_query; ++t) {
            if (metric == 0) oracle_euclidean_dist(q + (size_t)t * g, r, row, 1, n_ref, g);
            else oracle_mod_canberra_dist(q + (size_t)t * g, r, row, 1, n_ref, g, f);
            for (int j = 0; j < n_ref; ++j) { pr[j].d = (mask && mask[j]) ? NAN : row[j]; pr[j].i = j; }
            qsort(pr, (size_t)n_ref, sizeof(pair_t), cmp_pair);          /* full argsort, as the reference */
            for (int j = 0; j < k; ++j) {
                int s = j + (drop_first ? 1 : 0);
                out_idx[(size_t)t * k + j] = s < n_ref ? pr[s].i : -1;
                out_dist[(size_t)t * k + j] = s < n_ref ? pr[s].d : NAN;
            }
        }
        free(row); free(pr);
    }
    return 0;
}

void oracle_snn_counts(const int32_t* tgt_knn, const int32_t* ref_knn, int n_query, int k, uint8_t* out) {
    for (int t = 0; t < n_query; ++t) {
        const int32_t* a = tgt_knn + (size_t)t * k;
        for (int jj = 0; jj < k; ++jj) {
            const int32_t* b = ref_knn + (size_t)a[jj] * k;
            int c = 0;
            for (int u = 0; u < k; ++u) { int hit = 0; for (int v = 0; v < k; ++v) hit |= (a[v] == b[u]); c += hit; }
            out[(size_t)t * k + jj] = (uint8_t)c;
        }
    }
}

void oracle_scores(const int32_t* tgt_knn, const uint8_t* counts, const double* lut, int n_query, int k, int n_ref,
                   double mult, double* out) {
    memset(out, 0, sizeof(double) * (size_t)n_ref);
    for (int t = 0; t < n_query; ++t)
        for (int j = 0; j < k; ++j)
            if (counts[(size_t)t * k + j] > 0) out[tgt_knn[(size_t)t * k + j]] += lut[counts[(size_t)t * k + j]];
    for (int r = 0; r < n_ref; ++r) out[r] = mult * out[r] / n_query;
}
