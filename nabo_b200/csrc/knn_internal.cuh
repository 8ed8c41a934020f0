// Internal (non-ABI) entry points shared between the kNN translation units.
#pragma once
#include "common.cuh"

int nabo_knn_exact_launch(const double* q, int ldq, const double* r, int ldr, int n_query, int n_ref, int g,
                          int k, int metric, double f, const uint8_t* mask, int drop_first, int idx_offset,
                          const int* row_ids, const int* n_rows_dev, int32_t* out_idx, double* out_dist,
                          cudaStream_t st);

int nabo_rerank_launch(const double* q, int ldq, const double* r, int ldr, int n_query, int n_ref, int g, int k,
                       int metric, double f, const uint8_t* mask, int drop_first, int idx_offset,
                       const int32_t* cand, int n_cand, const float* cert_tau, const float* cert_eps,
                       int* fail_rows, int* fail_count, int32_t* out_idx, double* out_dist, cudaStream_t st);

size_t nabo_fast_workspace_bytes(int n_query, int n_ref, int g, int k, int metric);
int nabo_knn_fast(const double* q, int ldq, const double* r, int ldr, int n_query, int n_ref, int g, int k,
                  int metric, double f, const uint8_t* mask, int drop_first, int idx_offset, int32_t* out_idx,
                  double* out_dist, void* workspace, size_t workspace_bytes, int64_t* stats_host,
                  cudaStream_t st);
