// nabo_knn: engine selection and workspace layout for the fused distance + top-k path
// (replaces _calc_dist, nabo/_mapping.py:48-148).
#include "common.cuh"
#include "knn_internal.cuh"

extern "C" size_t nabo_knn_workspace_bytes(int n_query, int n_ref, int g, int k, int metric, int mode) {
    if (mode == NABO_MODE_EXACT) return 256;
    return nabo_fast_workspace_bytes(n_query, n_ref, g, k, metric);
}

static int knn_impl(const double* q, int ldq, const double* r, int ldr, int n_query, int n_ref, int g, int k,
                    int metric, double dist_factor, const uint8_t* ref_mask, int drop_first, int idx_offset,
                    int mode, int32_t* out_idx, double* out_dist, const NaboRoute& route, void* workspace,
                    size_t workspace_bytes, int64_t* stats_host, void* stream);

extern "C" int nabo_knn(const double* q, int ldq, const double* r, int ldr, int n_query, int n_ref, int g, int k,
                        int metric, double dist_factor, const uint8_t* ref_mask, int drop_first, int idx_offset,
                        int mode, int32_t* out_idx, double* out_dist, void* workspace, size_t workspace_bytes,
                        int64_t* stats_host, void* stream) {
    NaboRoute plain;
    plain.n_parts = 0;
    return knn_impl(q, ldq, r, ldr, n_query, n_ref, g, k, metric, dist_factor, ref_mask, drop_first, idx_offset, mode,
                    out_idx, out_dist, plain, workspace, workspace_bytes, stats_host, stream);
}

extern "C" int nabo_knn_routed(const double* q, int ldq, const double* r, int ldr, int n_query, int n_ref, int g,
                               int k, int metric, double dist_factor, const uint8_t* ref_mask, int drop_first,
                               int idx_offset, int mode, int n_parts, const int* part_bounds_host,
                               int32_t* const* part_idx_host, double* const* part_dist_host, void* workspace,
                               size_t workspace_bytes, int64_t* stats_host, void* stream) {
    NABO_ARG(n_parts >= 1 && n_parts <= NABO_MAX_PARTS, "knn_routed: n_parts=%d outside 1..%d", n_parts, NABO_MAX_PARTS);
    NABO_ARG(part_bounds_host && part_idx_host && part_dist_host, "knn_routed: null table");
    NABO_ARG(part_bounds_host[0] == 0 && part_bounds_host[n_parts] == n_query,
             "knn_routed: part bounds must run from 0 to n_query");
    NaboRoute route;
    route.n_parts = n_parts;
    for (int p = 0; p <= n_parts; ++p) route.bounds[p] = part_bounds_host[p];
    for (int p = 0; p < n_parts; ++p) {
        NABO_ARG(route.bounds[p] <= route.bounds[p + 1], "knn_routed: part bounds must not decrease");
        NABO_ARG(route.bounds[p] == route.bounds[p + 1] || (part_idx_host[p] && part_dist_host[p]),
                 "knn_routed: part %d has rows but no buffers", p);
        route.idx[p] = part_idx_host[p];
        route.dist[p] = part_dist_host[p];
    }
    return knn_impl(q, ldq, r, ldr, n_query, n_ref, g, k, metric, dist_factor, ref_mask, drop_first, idx_offset, mode,
                    nullptr, nullptr, route, workspace, workspace_bytes, stats_host, stream);
}

static int knn_impl(const double* q, int ldq, const double* r, int ldr, int n_query, int n_ref, int g, int k,
                    int metric, double dist_factor, const uint8_t* ref_mask, int drop_first, int idx_offset,
                    int mode, int32_t* out_idx, double* out_dist, const NaboRoute& route, void* workspace,
                    size_t workspace_bytes, int64_t* stats_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NABO_ARG(n_query >= 0 && n_ref >= 1 && g >= 1 && k >= 1, "knn: bad sizes n_query=%d n_ref=%d g=%d k=%d",
             n_query, n_ref, g, k);
    NABO_ARG(ldq >= g && ldr >= g, "knn: leading dimension smaller than g");
    NABO_ARG(metric == NABO_EUCLIDEAN || metric == NABO_MOD_CANBERRA || metric == NABO_COSINE,
             "knn: unknown metric %d", metric);
    NABO_ARG(metric != NABO_MOD_CANBERRA || dist_factor > 0.0, "knn: dist_factor must be > 0");
    if (n_query == 0) return 0;
    NABO_ARG(q && r && (route.n_parts > 0 || (out_idx && out_dist)), "knn: null pointer");
    if (stats_host) for (int i = 0; i < 8; ++i) stats_host[i] = 0;
    if (mode == NABO_MODE_EXACT) {
        NaboStageTimer tm(stats_host != nullptr, st);
        tm.begin();
        int rc = nabo_knn_exact_launch(q, ldq, r, ldr, n_query, n_ref, g, k, metric, dist_factor, ref_mask,
                                       drop_first, idx_offset, nullptr, nullptr, out_idx, out_dist, route, st);
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
                         idx_offset, out_idx, out_dist, route, workspace, workspace_bytes, stats_host, st);
}
