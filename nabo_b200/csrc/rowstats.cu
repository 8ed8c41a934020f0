// Per-row float32 statistics of a sparse count matrix in NumPy's own summation order: the device form of
// Dataset.set_sf (nabo/_dataset.py:548-592, `temp[keepGenesIdx].sum()` per cell) and Dataset.set_gene_stats
// (:594-637, `temp.mean()`, `temp[temp > 0].mean()`, `temp.var()`, `(temp > 0).sum()` per gene, with
// temp = densified counts of the kept cells times their size factors, float32).
//
// mu and sigma of the reference model come from these float32 reductions, and the projected coordinates -
// hence neighbour indices - are only bit-identical if they are.  NumPy reduces a contiguous float32 vector
// with its pairwise scheme (blocks of <= 128 elements, eight interleaved accumulators, halves split at a
// multiple of eight), so the kernel walks exactly that tree: one thread per row streams the row's dense
// vector (zeros included - they do not change a partial sum, but they decide where the tree splits) by
// merging the sorted non-zero list with the dense positions.  All rows have the same dense length, so the
// threads of a warp follow the same control flow.  No FMA contraction: every operation is an explicit
// round-to-nearest intrinsic.
#include "common.cuh"

namespace rowstats {

struct Cursor {
    const int32_t* idx;
    const float* val;
    const int32_t* pos_of;    // column -> dense position, -1 = not part of the dense vector
    const float* scale;       // per dense position (may be null = 1)
    long long p, end;
    long long next_pos;       // dense position of the next non-zero, or a value past the end
    float next_val;
    long long j;              // dense position the stream is at

    __device__ void seek() {
        next_pos = 0x7fffffffffffffffLL;
        while (p < end) {
            const int pos = pos_of[idx[p]];
            if (pos >= 0) {
                next_pos = pos;
                const float v = val[p];
                next_val = scale ? __fmul_rn(v, scale[pos]) : v;
                return;
            }
            ++p;
        }
    }
    __device__ void rewind(long long p0) { p = p0; j = 0; seek(); }
    // value at the current dense position, then advance
    __device__ float get() {
        float v = 0.f;
        if (next_pos == j) { v = next_val; ++p; seek(); }
        ++j;
        return v;
    }
    // next strictly positive value of the compressed vector temp[temp > 0]
    __device__ float get_nz() {
        for (;;) {
            const float v = next_val;
            ++p; seek();
            if (v > 0.f) return v;
        }
    }
};

// mode 0: x, 1: (x - c)^2, 2: compressed positive values
template <int MODE>
__device__ __forceinline__ float fetch(Cursor& c, float center) {
    if (MODE == 2) return c.get_nz();
    const float x = c.get();
    if (MODE == 0) return x;
    const float d = __fsub_rn(x, center);
    return __fmul_rn(d, d);
}

// numpy/_core/src/umath/loops_utils.h.src: @TYPE@_pairwise_sum, restated
template <int MODE>
__device__ float leaf_sum(Cursor& c, float center, long long n) {
    if (n < 8) {
        float res = 0.f;
        for (long long i = 0; i < n; ++i) res = __fadd_rn(res, fetch<MODE>(c, center));
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = fetch<MODE>(c, center);
    long long i = 8;
    const long long lim = n - (n % 8);
    for (; i < lim; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], fetch<MODE>(c, center));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, fetch<MODE>(c, center));
    return res;
}

template <int MODE>
__device__ float pairwise_sum(Cursor& c, float center, long long n) {
    // explicit recursion stack: size, state (0 = nothing done, 1 = left half done), left result
    long long sz[48];
    float left[48];
    int st[48];
    int sp = 0;
    sz[0] = n; st[0] = 0;
    float ret = 0.f;
    bool have_ret = false;
    for (;;) {
        if (!have_ret) {
            const long long m = sz[sp];
            if (m <= 128) {
                ret = leaf_sum<MODE>(c, center, m);
                have_ret = true;
            } else {
                long long n2 = m / 2;
                n2 -= n2 % 8;
                st[sp] = 0;
                ++sp;
                sz[sp] = n2;                     // descend into the left half
                continue;
            }
        }
        // a value is ready for the frame below
        if (sp == 0) return ret;
        const int parent = sp - 1;
        if (st[parent] == 0) {                   // it was the left half: keep it, descend into the right half
            left[parent] = ret;
            st[parent] = 1;
            long long n2 = sz[parent] / 2;
            n2 -= n2 % 8;
            sz[sp] = sz[parent] - n2;
            have_ret = false;
        } else {                                 // right half: combine, return to the grandparent
            ret = __fadd_rn(left[parent], ret);
            --sp;
        }
    }
}

__global__ void __launch_bounds__(128)
rowstats_kernel(const long long* __restrict__ indptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                int n_rows, const int32_t* __restrict__ pos_of, const float* __restrict__ scale, long long n_dense,
                int want_moments, float* __restrict__ out_sum, float* __restrict__ out_mean,
                float* __restrict__ out_nzmean, float* __restrict__ out_var, int32_t* __restrict__ out_npos) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    Cursor c;
    c.idx = idx; c.val = val; c.pos_of = pos_of; c.scale = scale;
    const long long p0 = indptr[row];
    c.end = indptr[row + 1];
    c.rewind(p0);
    const float total = pairwise_sum<0>(c, 0.f, n_dense);
    out_sum[row] = total;
    if (!want_moments) return;
    const float nf = (float)n_dense;
    const float mean = __fdiv_rn(total, nf);
    out_mean[row] = mean;
    // positive entries of the dense vector
    int npos = 0;
    c.rewind(p0);
    while (c.next_pos != 0x7fffffffffffffffLL) {
        if (c.next_val > 0.f) ++npos;
        ++c.p; c.seek();
    }
    out_npos[row] = npos;
    c.rewind(p0);
    const float ss = pairwise_sum<1>(c, mean, n_dense);
    out_var[row] = __fdiv_rn(ss, nf);
    float nzm = 0.f;
    if (npos > 0) {
        c.rewind(p0);
        nzm = __fdiv_rn(pairwise_sum<2>(c, 0.f, npos), (float)npos);
    }
    out_nzmean[row] = nzm;
}

}  // namespace rowstats

extern "C" int nabo_sparse_row_stats(const long long* indptr, const int32_t* idx, const float* val, int n_rows,
                                     int n_cols, const int32_t* pos_of_col, const float* scale, long long n_dense,
                                     int want_moments, float* out_sum, float* out_mean, float* out_nzmean,
                                     float* out_var, int32_t* out_npos, void* stream) {
    NABO_ARG(n_rows >= 0 && n_cols > 0 && n_dense >= 0, "sparse_row_stats: bad sizes");
    NABO_ARG(indptr && pos_of_col && out_sum, "sparse_row_stats: null pointer");
    NABO_ARG(!want_moments || (out_mean && out_nzmean && out_var && out_npos), "sparse_row_stats: null output");
    NABO_ARG(!want_moments || n_dense > 0, "sparse_row_stats: moments of an empty vector");
    if (n_rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    rowstats::rowstats_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(indptr, idx, val, n_rows, pos_of_col, scale, n_dense,
                                                                    want_moments, out_sum, out_mean, out_nzmean, out_var,
                                                                    out_npos);
    NABO_LAUNCH_CHECK("rowstats_kernel");
    return 0;
}
