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


// ------------------------------------------------------------------ dense projection on the FP64 tensor cores
// P[i][c] = sum_g p[i][g] * W[g][c] - bias[c],   p = (a * sf_i) in float32 (exact as a double),
//           W[g][c] = C[c][g] / sigma_g,  bias[c] = sum_g (mu_g / sigma_g + mean_g) * C[c][g]
// which is ((p - mu) / sigma - mean) . C[c] with the per-gene constants folded into the (G x n_comps) matrix once
// per call: the per-(cell, gene) FP64 divide and the per-cell recomputation of mean . C of the simple kernel
// above disappear, and what is left is one FP64 GEMM (cells x G) . (G x n_comps) on DMMA (mma.sync m8n8k4.f64;
// tcgen05 has no FP64 kind), FP64 accumulate - the accuracy the exact re-rank downstream needs.
//   block = 8 warps = 128 cells; a warp owns 16 cells x all components (NB blocks of 8 columns);
//   K loop over genes in chunks of 32, double-buffered in shared memory: the counts of the next chunk are
//   gathered (gene_idx), scaled and widened in registers while the tensor cores work on the current one.
constexpr int PD_CELLS = 128, PD_KC = 32, PD_SA = PD_KC + 4;     // A row stride = 4 mod 16 doubles: conflict-free

__global__ void __launch_bounds__(256)
project_prep_kernel(const double* __restrict__ comps, const double* __restrict__ mu, const double* __restrict__ sigma,
                    const double* __restrict__ mean, int G, int nc, int ncp, double* __restrict__ w) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < (long long)G * ncp) {
        const int g = (int)(e / ncp), c = (int)(e - (long long)g * ncp);
        w[e] = c < nc ? __ddiv_rn(comps[(long long)c * G + g], sigma[g]) : 0.0;
    }
}

// bias[c] = sum_g (mu_g / sigma_g + mean_g) * C[c][g]: one warp per component, lane-strided partial sums
__global__ void __launch_bounds__(256)
project_bias_kernel(const double* __restrict__ comps, const double* __restrict__ mu, const double* __restrict__ sigma,
                    const double* __restrict__ mean, int G, int nc, double* __restrict__ bias) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= nc) return;
    double b = 0.0;
    for (int g = lane; g < G; g += 32)
        b = fma(__dadd_rn(__ddiv_rn(mu[g], sigma[g]), mean[g]), comps[(long long)c * G + g], b);
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (lane == 0) bias[c] = b;
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NB>     // component blocks of 8 (n_comps padded to 8 * NB)
__global__ void __launch_bounds__(256, 2)
project_dmma_kernel(const float* __restrict__ counts, int ld, int n_cells, const int32_t* __restrict__ gene_idx, int G,
                    const float* __restrict__ sf, const double* __restrict__ w, const double* __restrict__ bias, int nc,
                    double* __restrict__ out, int ldo) {
    constexpr int NCP = 8 * NB, SB = NCP + 4;                    // B row stride = 4 or 12 mod 16 doubles
    extern __shared__ double psm[];
    double* as[2] = {psm, psm + PD_CELLS * PD_SA};
    double* bs[2] = {psm + 2 * PD_CELLS * PD_SA, psm + 2 * PD_CELLS * PD_SA + PD_KC * SB};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cell0 = blockIdx.x * PD_CELLS;
    // gather assignment: a warp gathers its own 16 cells, lane = gene of the chunk (one coalesced row segment per
    // load when the model genes are contiguous columns); lane r < 16 keeps the size factor of the warp's cell r
    const int wc0 = cell0 + warp * 16;
    const float my_sf = (lane < 16 && wc0 + lane < n_cells) ? sf[wc0 + lane] : 0.f;
    float av[16];
    auto gather = [&](int gbase) {
        const int g = gbase + lane;
        const int col = g < G ? gene_idx[g] : -1;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            float a = 0.f;
            if (col >= 0 && wc0 + r < n_cells) a = __ldg(counts + (long long)(wc0 + r) * ld + col);
            av[r] = __fmul_rn(a, __shfl_sync(0xffffffffu, my_sf, r));   // float32 product, as in the reference
        }
    };
    auto stage = [&](int buf, int gbase) {
        double* a = as[buf] + (warp * 16) * PD_SA + lane;
#pragma unroll
        for (int r = 0; r < 16; ++r) a[r * PD_SA] = (double)av[r];
        for (int e = threadIdx.x; e < PD_KC * NCP / 2; e += 256) {
            const int k = e / (NCP / 2), c = (e - k * (NCP / 2)) * 2;
            double2 v = make_double2(0.0, 0.0);
            if (gbase + k < G) v = *reinterpret_cast<const double2*>(w + (long long)(gbase + k) * NCP + c);
            *reinterpret_cast<double2*>(bs[buf] + k * SB + c) = v;
        }
    };
    double acc[2][NB][2];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < NB; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
    const int n_chunks = (G + PD_KC - 1) / PD_KC;
    gather(0);
    stage(0, 0);
    __syncthreads();
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < n_chunks) gather((ch + 1) * PD_KC);          // global loads in flight during the MMAs
        const double* a = as[buf] + (warp * 16 + (lane >> 2)) * PD_SA + (lane & 3);
        const double* b = bs[buf] + (lane & 3) * SB + (lane >> 2);
#pragma unroll
        for (int k4 = 0; k4 < PD_KC / 4; ++k4) {
            const double a0 = a[k4 * 4], a1 = a[8 * PD_SA + k4 * 4];
#pragma unroll
            for (int n = 0; n < NB; ++n) {
                const double bv = b[k4 * 4 * SB + n * 8];
                dmma_m8n8k4(acc[0][n][0], acc[0][n][1], a0, bv);
                dmma_m8n8k4(acc[1][n][0], acc[1][n][1], a1, bv);
            }
        }
        if (ch + 1 < n_chunks) stage(buf ^ 1, (ch + 1) * PD_KC);
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int cell = cell0 + warp * 16 + m * 8 + (lane >> 2);
        if (cell >= n_cells) continue;
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            const int c = n * 8 + 2 * (lane & 3);
            if (c < nc) out[(long long)cell * ldo + c] = acc[m][n][0] - bias[c];
            if (c + 1 < nc) out[(long long)cell * ldo + c + 1] = acc[m][n][1] - bias[c + 1];
        }
    }
}

extern "C" size_t nabo_project_dense_workspace_bytes(int G, int n_comps) {
    const int ncp = (n_comps + 7) / 8 * 8;
    return nabo_align_up((size_t)G * ncp * sizeof(double), 256) + nabo_align_up((size_t)n_comps * sizeof(double), 256) + 512;
}

extern "C" int nabo_project_dense_mma(const float* counts, int ld, int n_cells, const int32_t* gene_idx, int G,
                                      const float* sf, const double* mu, const double* sigma,
                                      const double* components, const double* mean, int n_comps, double* out,
                                      int ldo, void* workspace, size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NABO_ARG(n_cells >= 0 && G >= 1 && n_comps >= 1, "project: bad sizes");
    NABO_ARG(n_comps <= 128, "project: n_comps=%d > 128 unsupported", n_comps);
    NABO_ARG(ldo >= n_comps, "project: ldo < n_comps");
    if (n_cells == 0) return 0;
    NABO_ARG(counts && gene_idx && sf && mu && sigma && components && mean && out, "project: null pointer");
    const int nb = (n_comps + 7) / 8, ncp = nb * 8;
    NaboArena ar(workspace, workspace_bytes);
    double* w = ar.take<double>((size_t)G * ncp);
    double* bias = ar.take<double>(n_comps);
    if (!ar.ok) return nabo_set_error(NABO_EWORKSPACE, "project: workspace too small (%zu < %zu)", workspace_bytes,
                                      nabo_project_dense_workspace_bytes(G, n_comps));
    const long long tot = (long long)G * ncp;
    project_prep_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(components, mu, sigma, mean, G, n_comps, ncp, w);
    project_bias_kernel<<<(n_comps + 7) / 8, 256, 0, st>>>(components, mu, sigma, mean, G, n_comps, bias);
    NABO_LAUNCH_CHECK("project_prep_kernel");
    const int grid = (n_cells + PD_CELLS - 1) / PD_CELLS;
#define LAUNCH(NB_)                                                                                               \
    do {                                                                                                          \
        const size_t smem = (size_t)(2 * PD_CELLS * PD_SA + 2 * PD_KC * (8 * NB_ + 4)) * sizeof(double);          \
        NABO_CUDA(cudaFuncSetAttribute(project_dmma_kernel<NB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        project_dmma_kernel<NB_><<<grid, 256, smem, st>>>(counts, ld, n_cells, gene_idx, G, sf, w, bias, n_comps, out, ldo); \
    } while (0)
    switch (nb) {
        case 1: LAUNCH(1); break;  case 2: LAUNCH(2); break;  case 3: LAUNCH(3); break;  case 4: LAUNCH(4); break;
        case 5: LAUNCH(5); break;  case 6: LAUNCH(6); break;  case 7: LAUNCH(7); break;  case 8: LAUNCH(8); break;
        case 9: LAUNCH(9); break;  case 10: LAUNCH(10); break; case 11: LAUNCH(11); break; case 12: LAUNCH(12); break;
        case 13: LAUNCH(13); break; case 14: LAUNCH(14); break; case 15: LAUNCH(15); break; default: LAUNCH(16); break;
    }
#undef LAUNCH
    NABO_LAUNCH_CHECK("project_dmma_kernel");
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
