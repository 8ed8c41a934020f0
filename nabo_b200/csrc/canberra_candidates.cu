// FP32 CUDA-core candidate pass for the modified Canberra metric (nabo/_mapping.py:29-45).
//
// The metric is not a contraction (per dimension: compare, divide, select), so tensor cores do not
// apply; the bound is the FP32 / SFU pipes.  Two phases per 64 x 64 tile of (query, reference) pairs:
//
//   phase 1  no division: count the dimensions where |x - y| < f|x| (the only ones whose term is
//            below 1).  d(x, y) >= g - count, so a pair whose bound already reaches the query's
//            running threshold tau is rejected for 3 instructions per dimension.
//   phase 2  the survivors (a few per cent once tau is warm) go to a shared-memory work list and
//            are evaluated densely, one pair per thread: term = |x - y| * rcp(|x| + |y| + 0.01).
//
// Both phases evaluate the predicate with an optimistic margin (f|x|(1 + delta)), so the FP32
// score is a LOWER bound of the exact FP64 distance up to eps = nabo_cb_eps(g).  Per-query top-K' lives in
// shared memory (packed sortable keys, warp bitonic compaction).  The exact FP64 re-rank and the
// certificate D_k < tau - eps (rerank_kernel, NABO_CERT_LINEAR) make the final result identical to
// the exact engine; uncertified rows fall back to it.
#include "common.cuh"
#include "knn_internal.cuh"

namespace cb {

constexpr int TQ = 64, TR = 64, NT = 512;
constexpr int CAP = 128;                // candidate keys per query in shared memory

struct Smem {
    float *xs, *xt, *xa;                // [g][TQ]  x, f|x|(1+delta), |x|
    float *ys0, *ys1;                   // [g][TR]  y, double-buffered (cp.async prefetch of the next tile)
    unsigned long long* keys;           // [TQ][CAP]
    float* tau;                         // [TQ]
    int* cnt;                           // [TQ]
    unsigned short* work;               // [TQ*TR] packed (q << 8 | r)
    int* nwork;
};

__device__ __forceinline__ Smem carve(unsigned char* base, int g) {
    Smem s;
    s.keys = reinterpret_cast<unsigned long long*>(base);
    float* f = reinterpret_cast<float*>(s.keys + (size_t)TQ * CAP);
    s.xs = f; f += (size_t)g * TQ;
    s.xt = f; f += (size_t)g * TQ;
    s.xa = f; f += (size_t)g * TQ;
    s.ys0 = f; f += (size_t)g * TR;
    s.ys1 = f; f += (size_t)g * TR;
    s.tau = f; f += TQ;
    s.cnt = reinterpret_cast<int*>(f);
    s.nwork = s.cnt + TQ;
    s.work = reinterpret_cast<unsigned short*>(s.nwork + 4);
    return s;
}
static size_t smem_bytes(int g) {
    return (size_t)TQ * CAP * 8 + ((size_t)g * (3 * TQ + 2 * TR) + TQ) * 4 + (TQ + 4) * 4 + (size_t)TQ * TR * 2;
}

// FP32 copy of a row-major FP64 matrix, pre-tiled k-major: out[(tile * g + k) * 64 + row_in_tile].
// A tile of the main kernel is then a straight, coalesced float4 copy.
__global__ void __launch_bounds__(256)
pretile_kernel(const double* __restrict__ x, int ld, int n, int g, float* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)((n + 63) / 64) * g * 64;
    if (e >= total) return;
    const int rl = (int)(e & 63);
    const long long tk = e >> 6;
    const int k = (int)(tk % g);
    const long long tile = tk / g;
    const long long row = tile * 64 + rl;
    out[e] = row < n ? (float)x[row * ld + k] : 0.f;
}

__device__ __forceinline__ void compact(const Smem& s, int q, int kprime, int lane) {
    unsigned long long* k = s.keys + (size_t)q * CAP;
    const int n = s.cnt[q];
    for (int i = n + lane; i < CAP; i += 32) k[i] = ~0ull;
    __syncwarp();
    warp_bitonic_sort_u64(k, CAP, lane);
    if (lane == 0 && n >= kprime) {
        s.cnt[q] = kprime;
        s.tau[q] = sortable_to_float((uint32_t)(k[kprime - 1] >> 32));
    }
    __syncwarp();
}

__global__ void __launch_bounds__(NT, 1)
candidates_kernel(const float* __restrict__ qt, const float* __restrict__ rt, int n_query, int n_ref, int g, float fm,
                  const uint8_t* __restrict__ mask, int kprime, int32_t* __restrict__ cand,
                  float* __restrict__ tau_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem s = carve(smem_raw, g);
    const int q0 = blockIdx.x * TQ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tr = threadIdx.x & 15, tq = threadIdx.x >> 4;      // 16 reference groups of 4, 32 query groups of 2
    const int nvec = g * TQ / 4;

    {
        const float4* src = reinterpret_cast<const float4*>(qt + (size_t)blockIdx.x * g * TQ);
        for (int e = threadIdx.x; e < nvec; e += NT) {
            const float4 x = src[e];
            reinterpret_cast<float4*>(s.xs)[e] = x;
            reinterpret_cast<float4*>(s.xa)[e] = make_float4(fabsf(x.x), fabsf(x.y), fabsf(x.z), fabsf(x.w));
            // NaN x -> NaN threshold -> predicate false -> term 1 (as the reference)
            reinterpret_cast<float4*>(s.xt)[e] = make_float4(fm * fabsf(x.x), fm * fabsf(x.y), fm * fabsf(x.z), fm * fabsf(x.w));
        }
    }
    if (threadIdx.x < TQ) { s.cnt[threadIdx.x] = 0; s.tau[threadIdx.x] = CUDART_INF_F; }
    if (threadIdx.x < 2) s.nwork[threadIdx.x] = 0;

    // reference tiles are double-buffered: tile j+1 streams in with cp.async while tile j is processed
    auto prefetch = [&](int tile, float* dst) {
        const float4* src = reinterpret_cast<const float4*>(rt + (size_t)tile * g * TR);
        for (int e = threadIdx.x; e < nvec; e += NT) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(reinterpret_cast<float4*>(dst) + e);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + e) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int n_tiles = (n_ref + TR - 1) / TR;
    prefetch(0, s.ys0);
    for (int tile = 0; tile < n_tiles; ++tile) {
        const int r0 = tile * TR;
        float* ys = (tile & 1) ? s.ys1 : s.ys0;
        int* nwork = s.nwork + (tile & 1);                 // work-list counters alternate between tiles
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                   // tile landed; phase 2 of the previous tile finished
        if (tile + 1 < n_tiles) prefetch(tile + 1, (tile & 1) ? s.ys0 : s.ys1);
        // compaction of the previous tile's appends rides inside the phase-1 region (no extra barrier):
        // phase 1 only reads tau, and a slightly stale tau is still a valid threshold
        for (int ql = warp; ql < TQ; ql += NT / 32)
            if (s.cnt[ql] > CAP - TR) compact(s, ql, kprime, lane);
        // ---- phase 1: unsaturated-dimension counts of a 2 x 4 micro tile (float counters: FSET + FADD)
        float c[2][4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) c[a][b] = 0.f;
#pragma unroll 2
        for (int k = 0; k < g; ++k) {
            const float2 xv = *reinterpret_cast<const float2*>(s.xs + k * TQ + tq * 2);
            const float2 tv = *reinterpret_cast<const float2*>(s.xt + k * TQ + tq * 2);
            const float4 yv = *reinterpret_cast<const float4*>(ys + k * TR + tr * 4);
            const float xx[2] = {xv.x, xv.y}, tt[2] = {tv.x, tv.y};
            const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) c[a][b] += (fabsf(xx[a] - yy[b]) < tt[a]) ? 1.0f : 0.0f;
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int ql = tq * 2 + a;
            if (q0 + ql >= n_query) continue;
            const float tau = s.tau[ql];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int rl = tr * 4 + b, j = r0 + rl;
                if (j >= n_ref || (mask && mask[j])) continue;
                if ((float)g - c[a][b] < tau) {                   // d >= g - count: bound still below tau
                    const int pos = atomicAdd(nwork, 1);
                    s.work[pos] = (unsigned short)((ql << 8) | rl);
                }
            }
        }
        __syncthreads();
        // ---- phase 2: dense evaluation of the survivors
        const int nw = *nwork;
        if (threadIdx.x == 0) s.nwork[(tile + 1) & 1] = 0;       // nobody touches the other counter until the next tile
        for (int w = threadIdx.x; w < nw; w += NT) {
            const int ql = s.work[w] >> 8, rl = s.work[w] & 255;
            float acc = 0.f;
            for (int k = 0; k < g; ++k) {
                const float x = s.xs[k * TQ + ql], y = ys[k * TR + rl];
                const float num = fabsf(x - y);
                const float term = __fdividef(num, s.xa[k * TQ + ql] + (fabsf(y) + 0.01f));
                acc += (num < s.xt[k * TQ + ql]) ? term : 1.0f;
            }
            if (acc < s.tau[ql]) {
                const int pos = atomicAdd(&s.cnt[ql], 1);
                s.keys[(size_t)ql * CAP + pos] =
                    ((unsigned long long)float_to_sortable(acc) << 32) | (unsigned)(r0 + rl);
            }
        }
    }
    __syncthreads();
    for (int ql = warp; ql < TQ; ql += NT / 32) {
        const int qi = q0 + ql;
        if (qi >= n_query) continue;
        const int n = s.cnt[ql];
        const float old_tau = s.tau[ql];
        compact(s, ql, kprime, lane);
        const unsigned long long* k = s.keys + (size_t)ql * CAP;
        const int nc = n < kprime ? n : kprime;
        for (int i = lane; i < kprime; i += 32) cand[(long long)qi * kprime + i] = i < nc ? (int32_t)(uint32_t)k[i] : -1;
        if (lane == 0) tau_out[qi] = n >= kprime ? s.tau[ql] : old_tau;
    }
}

}  // namespace cb

bool nabo_cb_supported(int g, int k, int drop_first) {
    const int ksel = k + (drop_first ? 1 : 0);
    return cb::smem_bytes(g) <= 227 * 1024 && ksel + 8 <= cb::CAP - cb::TR;
}

int nabo_cb_kprime(int k, int drop_first) {
    const int ksel = k + (drop_first ? 1 : 0);
    int kp = ksel + (ksel / 4 > 8 ? ksel / 4 : 8);
    if (kp > cb::CAP - cb::TR) kp = cb::CAP - cb::TR;
    return kp;
}

size_t nabo_cb_pretile_floats(int n, int g) { return (size_t)((n + 63) / 64) * g * 64; }

int nabo_cb_candidates(const double* q, int ldq, const double* r, int ldr, int n_query, int n_ref, int g, int k,
                       double f, const uint8_t* mask, int drop_first, float* qt, float* rt, int32_t* cand, float* tau,
                       cudaStream_t st) {
    const int kprime = nabo_cb_kprime(k, drop_first);
    {
        const size_t tq = nabo_cb_pretile_floats(n_query, g), tr = nabo_cb_pretile_floats(n_ref, g);
        cb::pretile_kernel<<<(unsigned)((tq + 255) / 256), 256, 0, st>>>(q, ldq, n_query, g, qt);
        cb::pretile_kernel<<<(unsigned)((tr + 255) / 256), 256, 0, st>>>(r, ldr, n_ref, g, rt);
        NABO_LAUNCH_CHECK("cb::pretile_kernel");
    }
    const size_t smem = cb::smem_bytes(g);
    NABO_CUDA(cudaFuncSetAttribute(cb::candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // Optimistic margin of the saturation predicate: whenever the exact FP64 test |x-y| < f|x| holds, the
    // FP32 test on the rounded inputs must hold too.  Near the threshold |y| <= (1+f)|x|, so the FP32
    // difference is off by at most ~1.2e-7 (2+f)|x|; delta covers it with a 3x reserve.
    const double delta = fmax(1e-5, 4e-7 * (2.0 + f) / f);
    const float fm = (float)(f * (1.0 + delta));
    cb::candidates_kernel<<<(n_query + cb::TQ - 1) / cb::TQ, cb::NT, smem, st>>>(qt, rt, n_query, n_ref, g, fm,
                                                                              mask, kprime, cand, tau);
    NABO_LAUNCH_CHECK("cb::candidates_kernel");
    return 0;
}

// |FP32 score - (a lower bound of) the exact distance|: sequential FP32 sum of g terms <= 1 plus the
// per-term rounding of the difference, the denominator and the approximate division.
double nabo_cb_eps(int g) { return 1.2e-7 * (double)g * g + 2e-6 * g; }
