// Fast kNN engine: candidate pass + exact FP64 re-rank + certificate + exact fallback.
// (candidate kernels are added per metric; until one exists for a metric the engine
//  reports NABO_EUNSUPPORTED rather than silently running something else)
#include "knn_internal.cuh"

size_t nabo_fast_workspace_bytes(int n_query, int n_ref, int g, int k, int metric) {
    (void)n_query; (void)n_ref; (void)g; (void)k; (void)metric;
    return 256;
}

int nabo_knn_fast(const double*, int, const double*, int, int, int, int, int, int metric, double, const uint8_t*,
                  int, int, int32_t*, double*, void*, size_t, int64_t*, cudaStream_t) {
    return nabo_set_error(NABO_EUNSUPPORTED, "knn: fast engine not built for metric %d", metric);
}
