// nabo_knn: engine selection and workspace layout for the fused distance + top-k path
// (replaces _calc_dist, nabo/_mapping.py:48-148).
#include "common.cuh"
#include "knn_internal.cuh"

extern "C" size_t nabo_knn_workspace_bytes(int n_query, int n_ref, int g, int k, int metric, int mode) {
    if (mode == NABO_MODE_EXACT) return 256;
    return nabo_fast_workspace_bytes(n_query, n_ref, g, k, metric);
}

extern "C" int nabo_knn(const double* q, int ldq, const double* r, int ldr, int n_query, int n_ref, int g, int k,
                        int metric, double dist_factor, const uint8_t* ref_mask, int drop_first, int idx_offset,
                        int mode, int32_t* out_idx, double* out_dist, void* workspace, size_t workspace_bytes,
                        int64_t* stats_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NABO_ARG(n_query >= 0 && n_ref >= 1 && g >= 1 && k >= 1, "knn: bad sizes n_query=%d n_ref=%d g=%d k=%d",
             n_query, n_ref, g, k);
    NABO_ARG(ldq >= g && ldr >= g, "knn: leading dimension smaller than g");
    NABO_ARG(metric == NABO_EUCLIDEAN || metric == NABO_MOD_CANBERRA || metric == NABO_COSINE,
             "knn: unknown metric %d", metric);
    NABO_ARG(metric != NABO_MOD_CANBERRA || dist_factor > 0.0, "knn: dist_factor must be > 0");
    if (n_query == 0) return 0;
    NABO_ARG(q && r && out_idx && out_dist, "knn: null pointer");
    if (stats_host) for (int i = 0; i < 8; ++i) stats_host[i] = 0;
    if (mode == NABO_MODE_EXACT) {
        NaboStageTimer tm(stats_host != nullptr, st);
        tm.begin();
        int rc = nabo_knn_exact_launch(q, ldq, r, ldr, n_query, n_ref, g, k, metric, dist_factor, ref_mask,
                                       drop_first, idx_offset, nullptr, nullptr, out_idx, out_dist, st);
        tm.end(0);
        if (rc) return rc;
        if (stats_host) {
            NABO_CUDA(cudaStreamSynchronize(st));
            stats_host[1] = n_query;
            stats_host[3] = 1;
            stats_host[4] = tm.ns(0);
        }
        return 0;
    }
    NABO_ARG(mode == NABO_MODE_FAST, "knn: unknown mode %d", mode);
    return nabo_knn_fast(q, ldq, r, ldr, n_query, n_ref, g, k, metric, dist_factor, ref_mask, drop_first,
                         idx_offset, out_idx, out_dist, workspace, workspace_bytes, stats_host, st);
}
