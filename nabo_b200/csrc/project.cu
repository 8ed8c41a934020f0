// Scaling + PCA projection of target counts into the reference's HVG/PCA space.
//   get_scaled_values core   nabo/_dataset.py:905-913   z = ((a * sf_i)[f32] - mu) / sigma  [f64]
//   transform_pca            nabo/_dataset.py:1028      P = z @ C^T - mean @ C^T            [f64]
// FP64 on the CUDA cores on purpose: the re-rank that makes neighbour indices bit-exact
// needs the projected coordinates at FP64 accuracy (a tensor-core projection would move
// k-th/k+1-th neighbour boundaries by ~1e-3 relative), and the FP64 work is < 1 % of the
// kNN pass that follows (DESIGN.md "projection").
//   dense: tiled GEMM, 32 cells x all components per block, z formed on the fly in smem.
//   CSR:   P = P0 + sum over non-zeros (z - z0_g) * C[:, g]; one warp per cell, lanes own
//          components, so the sum runs in column order and needs no reduction.
#include "common.cuh"

constexpr int PJ_CELLS = 32;    // cells per block
constexpr int PJ_GK = 32;       // genes per smem step
constexpr int PJ_NT = 256;

__device__ __forceinline__ double scaled_value(float a, float sf, double mu, double sigma) {
    float p = __fmul_rn(a, sf);                       // float32 product, as in the reference
    return __ddiv_rn(__dsub_rn((double)p, mu), sigma);
}

template <int CPT>   // components per thread (strided by 8)
__global__ void __launch_bounds__(PJ_NT)
project_dense_kernel(const float* __restrict__ counts, int ld, int n_cells, const int32_t* __restrict__ gene_idx,
                     int G, const float* __restrict__ sf, const double* __restrict__ mu,
                     const double* __restrict__ sigma, const double* __restrict__ comps,
                     const double* __restrict__ mean, int nc, double* __restrict__ out, int ldo) {
    __shared__ double zs[PJ_GK][PJ_CELLS + 1];
    extern __shared__ double cs[];                    // [nc][PJ_GK + 1]
    const int cell_l = threadIdx.x >> 3;              // 0..31
    const int cg = threadIdx.x & 7;                   // component group
    const int cell0 = blockIdx.x * PJ_CELLS;
    double acc[CPT], macc[CPT];
#pragma unroll
    for (int u = 0; u < CPT; ++u) { acc[u] = 0.0; macc[u] = 0.0; }
    for (int g0 = 0; g0 < G; g0 += PJ_GK) {
        __syncthreads();
        for (int e = threadIdx.x; e < PJ_GK * PJ_CELLS; e += PJ_NT) {
            const int cl = e / PJ_GK, gg = e - cl * PJ_GK;
            const int g = g0 + gg, cell = cell0 + cl;
            double z = 0.0;
            if (g < G && cell < n_cells) {
                const int col = gene_idx[g];
                const float a = col >= 0 ? counts[(long long)cell * ld + col] : 0.0f;
                z = scaled_value(a, sf[cell], mu[g], sigma[g]);
            }
            zs[gg][cl] = z;
        }
        for (int e = threadIdx.x; e < nc * PJ_GK; e += PJ_NT) {
            const int c = e / PJ_GK, gg = e - c * PJ_GK;
            cs[c * (PJ_GK + 1) + gg] = (g0 + gg) < G ? comps[(long long)c * G + g0 + gg] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int gg = 0; gg < PJ_GK; ++gg) {
            const double z = zs[gg][cell_l];
            const double mg = (g0 + gg) < G ? mean[g0 + gg] : 0.0;
#pragma unroll
            for (int u = 0; u < CPT; ++u) {
                const int c = cg + 8 * u;
                if (c < nc) {
                    const double cv = cs[c * (PJ_GK + 1) + gg];
                    acc[u] = fma(z, cv, acc[u]);
                    macc[u] = fma(mg, cv, macc[u]);
                }
            }
        }
    }
    const int cell = cell0 + cell_l;
    if (cell < n_cells) {
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
            const int c = cg + 8 * u;
            if (c < nc) out[(long long)cell * ldo + c] = acc[u] - macc[u];   // X@C^T - mean@C^T
        }
    }
}

extern "C" int nabo_project_dense(const float* counts, int ld, int n_cells, const int32_t* gene_idx, int G,
                                  const float* sf, const double* mu, const double* sigma,
                                  const double* components, const double* mean, int n_comps, double* out,
                                  int ldo, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NABO_ARG(n_cells >= 0 && G >= 1 && n_comps >= 1, "project: bad sizes");
    NABO_ARG(n_comps <= 128, "project: n_comps=%d > 128 unsupported", n_comps);
    NABO_ARG(ldo >= n_comps, "project: ldo < n_comps");
    if (n_cells == 0) return 0;
    NABO_ARG(counts && gene_idx && sf && mu && sigma && components && mean && out, "project: null pointer");
    const size_t smem = (size_t)n_comps * (PJ_GK + 1) * sizeof(double);
    const int grid = (n_cells + PJ_CELLS - 1) / PJ_CELLS;
#define LAUNCH(CPT)                                                                                     \
    project_dense_kernel<CPT><<<grid, PJ_NT, smem, st>>>(counts, ld, n_cells, gene_idx, G, sf, mu, sigma, \
                                                         components, mean, n_comps, out, ldo)
    if (n_comps <= 32) LAUNCH(4);
    else if (n_comps <= 64) LAUNCH(8);
    else LAUNCH(16);
#undef LAUNCH
    NABO_LAUNCH_CHECK("project_dense_kernel");
    return 0;
}

// ------------------------------------------------------------------ scaling alone (get_scaled_values)
__global__ void __launch_bounds__(256)
scale_dense_kernel(const float* __restrict__ counts, int ld, int n_cells, const int32_t* __restrict__ gene_idx,
                   int G, const float* __restrict__ sf, const double* __restrict__ mu,
                   const double* __restrict__ sigma, double* __restrict__ out, int ldo) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)n_cells * G) return;
    const int cell = (int)(e / G), g = (int)(e - (long long)cell * G);
    const int col = gene_idx[g];
    const float a = col >= 0 ? counts[(long long)cell * ld + col] : 0.0f;
    out[(long long)cell * ldo + g] = scaled_value(a, sf[cell], mu[g], sigma[g]);
}

extern "C" int nabo_scale_dense(const float* counts, int ld, int n_cells, const int32_t* gene_idx, int G,
                                const float* sf, const double* mu, const double* sigma, double* out, int ldo,
                                void* stream) {
    NABO_ARG(n_cells >= 0 && G >= 1 && ldo >= G, "scale: bad sizes");
    if (n_cells == 0) return 0;
    NABO_ARG(counts && gene_idx && sf && mu && sigma && out, "scale: null pointer");
    const long long tot = (long long)n_cells * G;
    scale_dense_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(counts, ld, n_cells, gene_idx, G,
                                                                                     sf, mu, sigma, out, ldo);
    NABO_LAUNCH_CHECK("scale_dense_kernel");
    return 0;
}

// ------------------------------------------------------------------ CSR form
// workspace: ct [G][nc] (components transposed), z0 [G], p0 [nc]
__global__ void __launch_bounds__(256)
csr_prep_kernel(const double* __restrict__ comps, const double* __restrict__ mu, const double* __restrict__ sigma,
                int G, int nc, double* __restrict__ ct, double* __restrict__ z0) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < (long long)G * nc) {
        const int g = (int)(e / nc), c = (int)(e - (long long)g * nc);
        ct[e] = comps[(long long)c * G + g];
    }
    if (e < G) z0[e] = scaled_value(0.0f, 1.0f, mu[e], sigma[e]);   // (0*sf - mu)/sigma
}

// p0[c] = sum_g (z0[g] - mean[g]) * C[c][g]  in ascending g (one thread per component)
__global__ void csr_p0_kernel(const double* __restrict__ ct, const double* __restrict__ z0,
                              const double* __restrict__ mean, int G, int nc, double* __restrict__ p0) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    double s = 0.0, m = 0.0;
    for (int g = 0; g < G; ++g) {
        const double cv = ct[(long long)g * nc + c];
        s = fma(z0[g], cv, s);
        m = fma(mean[g], cv, m);
    }
    p0[c] = s - m;
}

__global__ void __launch_bounds__(256)
project_csr_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ col,
                   const float* __restrict__ val, int n_cells, const int32_t* __restrict__ gene_pos,
                   int n_genes_total, const float* __restrict__ sf, const double* __restrict__ mu,
                   const double* __restrict__ sigma, const double* __restrict__ ct,
                   const double* __restrict__ z0, const double* __restrict__ p0, int nc,
                   double* __restrict__ out, int ldo) {
    const int lane = threadIdx.x & 31;
    const int cell = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (cell >= n_cells) return;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const long long lo = indptr[cell], hi = indptr[cell + 1];
    const float s = sf[cell];
    for (long long p0i = lo; p0i < hi; p0i += 32) {
        // lanes fetch 32 non-zeros, then the warp walks them in order
        const long long p = p0i + lane;
        int pos = -1;
        double dz = 0.0;
        if (p < hi) {
            const int cidx = col[p];
            if (cidx >= 0 && cidx < n_genes_total) pos = gene_pos[cidx];
            if (pos >= 0) dz = scaled_value(val[p], s, mu[pos], sigma[pos]) - z0[pos];
        }
        const int nloc = (int)((hi - p0i) < 32 ? (hi - p0i) : 32);
        for (int t = 0; t < nloc; ++t) {
            const int pt = __shfl_sync(0xffffffffu, pos, t);
            const double dt = __shfl_sync(0xffffffffu, dz, t);
            if (pt < 0) continue;
            const double* crow = ct + (long long)pt * nc;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = lane + 32 * u;
                if (c < nc) acc[u] = fma(dt, crow[c], acc[u]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int c = lane + 32 * u;
        if (c < nc) out[(long long)cell * ldo + c] = p0[c] + acc[u];
    }
}

extern "C" size_t nabo_project_csr_workspace_bytes(int G, int n_comps) {
    return nabo_align_up((size_t)G * n_comps * sizeof(double), 256) + nabo_align_up((size_t)G * sizeof(double), 256) +
           nabo_align_up((size_t)n_comps * sizeof(double), 256) + 1024;
}

extern "C" int nabo_project_csr(const int64_t* indptr, const int32_t* col, const float* val, int n_cells,
                                const int32_t* gene_pos, int n_genes_total, int G, const float* sf,
                                const double* mu, const double* sigma, const double* components,
                                const double* mean, int n_comps, double* out, int ldo, void* workspace,
                                size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NABO_ARG(n_cells >= 0 && G >= 1 && n_comps >= 1 && n_comps <= 128, "project_csr: bad sizes (n_comps <= 128)");
    NABO_ARG(ldo >= n_comps, "project_csr: ldo < n_comps");
    if (n_cells == 0) return 0;
    NABO_ARG(indptr && gene_pos && sf && mu && sigma && components && mean && out, "project_csr: null pointer");
    NaboArena ar(workspace, workspace_bytes);
    double* ct = ar.take<double>((size_t)G * n_comps);
    double* z0 = ar.take<double>(G);
    double* p0 = ar.take<double>(n_comps);
    if (!ar.ok) return nabo_set_error(NABO_EWORKSPACE, "project_csr: workspace too small (%zu < %zu)",
                                      workspace_bytes, nabo_project_csr_workspace_bytes(G, n_comps));
    const long long tot = (long long)G * n_comps;
    csr_prep_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(components, mu, sigma, G, n_comps, ct, z0);
    csr_p0_kernel<<<(n_comps + 63) / 64, 64, 0, st>>>(ct, z0, mean, G, n_comps, p0);
    project_csr_kernel<<<(n_cells + 7) / 8, 256, 0, st>>>(indptr, col, val, n_cells, gene_pos, n_genes_total, sf,
                                                         mu, sigma, ct, z0, p0, n_comps, out, ldo);
    NABO_LAUNCH_CHECK("project_csr_kernel");
    return 0;
}
