// Mapping specificity (nabo/_graph.py:794-824): for every target cell, the mean of the unweighted
// shortest-path lengths, in the reference graph, between all pairs of reference cells it maps to.
// The reference runs one networkx BFS per pair (k(k-1)/2 per target); here ONE bit-parallel multi-source
// search per target finds all pair distances at once:
//
//   * source i of the target owns bit i of a 64-bit mask; cur[v] = sources whose ball contains node v;
//   * level L: every frontier node u pushes the bits that reached it at level L-1 (delta[u]) to its
//     neighbours: acc[v] |= delta[u] & ~cur[v]; a node whose acc was zero joins the next frontier;
//   * commit: delta[v] = acc[v] & ~cur[v].  All balls grow in lock step, so the first time bit i arrives
//     at a node that already holds bit j the pair is 2L - 1 apart, and two bits arriving together are 2L
//     apart (pass A before pass B, so the shorter meeting always wins): every pair is settled after
//     ceil(d / 2) levels - the balls stay small - instead of the d levels a one-sided search needs;
//   * a source that has met every partner stops propagating; the search ends when all k(k-1)/2 pairs
//     are settled (or the frontier dies: no path, reported to the host, where the reference would raise
//     NetworkXNoPath).
//
// One CTA per target at a time, persistent over the targets; per-CTA scratch (cur, acc, frontier lists)
// lives in the caller's workspace and is cleaned by walking the list of touched nodes, so the arrays are
// zero again when the next target starts.  Integer outputs (sum of distances, pairs found): exact and
// independent of scheduling.
#include "common.cuh"

namespace spec {

constexpr int NT = 256;
constexpr int MAXK = 64;

struct Scratch {
    unsigned long long* cur;     // [m]
    unsigned long long* acc;     // [m]
    unsigned long long* dl[2];   // [m] delta of the frontier entries (parallel to fr[])
    int* fr[2];                  // [m] frontier node lists
    int* touched;                // [m]
    uint8_t* tpos;               // [m] 0 = not a mapped cell of the current target, else position + 1
};

static size_t scratch_bytes(int m) {
    const size_t mm = nabo_align_up((size_t)m, 64);
    return mm * (8 + 8 + 16 + 8 + 4 + 1) + 1024;
}

__global__ void __launch_bounds__(NT)
specificity_kernel(const long long* __restrict__ indptr, const int32_t* __restrict__ indices, int m,
                   const int32_t* __restrict__ tgt_knn, const uint8_t* __restrict__ counts, int n_query, int k,
                   long long* __restrict__ out_sum, int32_t* __restrict__ out_pairs, int32_t* __restrict__ out_nmapped,
                   unsigned char* __restrict__ ws, size_t per_block) {
    __shared__ int s_src[MAXK];
    __shared__ int s_nt, s_nf[2], s_ntouched;
    __shared__ unsigned long long s_sum;
    __shared__ int s_pairs;
    __shared__ unsigned long long s_found[MAXK];   // s_found[lo] bit hi: pair {lo, hi} has its distance (lo < hi)
    __shared__ int s_nfound[MAXK];                 // partners found per source
    __shared__ unsigned long long s_active;        // sources still missing a partner: only their bits keep travelling

    Scratch sc;
    {
        const size_t mm = ((size_t)m + 63) / 64 * 64;
        unsigned char* b = ws + (size_t)blockIdx.x * per_block;
        sc.cur = reinterpret_cast<unsigned long long*>(b); b += mm * 8;
        sc.acc = reinterpret_cast<unsigned long long*>(b); b += mm * 8;
        sc.dl[0] = reinterpret_cast<unsigned long long*>(b); b += mm * 8;
        sc.dl[1] = reinterpret_cast<unsigned long long*>(b); b += mm * 8;
        sc.fr[0] = reinterpret_cast<int*>(b); b += mm * 4;
        sc.fr[1] = reinterpret_cast<int*>(b); b += mm * 4;
        sc.touched = reinterpret_cast<int*>(b); b += mm * 4;
        sc.tpos = b;
    }

    // pair {i, j} met with total path length dist: count it once
    auto record = [&](int i, int j, int dist) {
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        const unsigned long long bit = 1ull << hi;
        if (!(atomicOr(&s_found[lo], bit) & bit)) {
            atomicAdd(&s_sum, (unsigned long long)dist);
            atomicAdd(&s_pairs, 1);
            atomicAdd(&s_nfound[lo], 1);
            atomicAdd(&s_nfound[hi], 1);
        }
    };

    for (int t = blockIdx.x; t < n_query; t += gridDim.x) {
        // ---- the target's mapped reference cells = its edges (snn count > 0), nabo/_mapping.py:195-198
        if (threadIdx.x == 0) {
            int nt = 0;
            for (int j = 0; j < k; ++j) {
                const int r = tgt_knn[(size_t)t * k + j];
                if (r >= 0 && r < m && counts[(size_t)t * k + j] > 0) s_src[nt++] = r;
            }
            s_nt = nt;
            s_sum = 0ull;
            s_pairs = 0;
            s_nf[0] = nt;
            s_nf[1] = 0;
            s_ntouched = nt;
            s_active = nt >= 64 ? ~0ull : ((1ull << nt) - 1ull);
        }
        if (threadIdx.x < MAXK) { s_found[threadIdx.x] = 0ull; s_nfound[threadIdx.x] = 0; }
        __syncthreads();
        const int nt = s_nt;
        if (nt < 2) {
            if (threadIdx.x == 0) { out_sum[t] = 0; out_pairs[t] = 0; out_nmapped[t] = nt; }
            __syncthreads();
            continue;
        }
        if (threadIdx.x < nt) {
            const int v = s_src[threadIdx.x];
            sc.cur[v] = 1ull << threadIdx.x;
            sc.fr[0][threadIdx.x] = v;
            sc.dl[0][threadIdx.x] = 1ull << threadIdx.x;
            sc.touched[threadIdx.x] = v;
        }
        __syncthreads();

        // All sources grow their balls in lock step.  When ball i (radius L) first meets ball j at a node, the
        // pair's distance is 2L - 1 if j was there since the level before, 2L if both arrive now: every pair is
        // settled after ceil(d / 2) levels, long before a one-sided search from i would reach j.
        const int want = nt * (nt - 1) / 2;
        int cb = 0;
        for (int level = 1;; ++level) {
            const int nf = s_nf[cb];
            const unsigned long long active = s_active;
            // ---- expand: one warp per frontier node, lanes over its adjacency
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            for (int f = warp; f < nf; f += NT / 32) {
                const int u = sc.fr[cb][f];
                const unsigned long long du = sc.dl[cb][f] & active;      // a source that has found every partner stops here
                if (du == 0ull) continue;
                const long long e0 = indptr[u], e1 = indptr[u + 1];
                for (long long e = e0 + lane; e < e1; e += 32) {
                    const int v = indices[e];
                    const unsigned long long nb = du & ~sc.cur[v];
                    if (nb) {
                        const unsigned long long old = atomicOr(&sc.acc[v], nb);
                        if (old == 0ull) {
                            const int pos = atomicAdd(&s_nf[cb ^ 1], 1);
                            sc.fr[cb ^ 1][pos] = v;
                        }
                    }
                }
            }
            __syncthreads();
            // ---- commit pass A: new bits meet the bits that were already there (distance 2L - 1)
            const int nn = s_nf[cb ^ 1];
            for (int f = threadIdx.x; f < nn; f += NT) {
                const int v = sc.fr[cb ^ 1][f];
                const unsigned long long c = sc.cur[v];
                const unsigned long long d = sc.acc[v] & ~c;
                sc.acc[v] = 0ull;
                sc.cur[v] = c | d;
                sc.dl[cb ^ 1][f] = d;
                if (c == 0ull) sc.touched[atomicAdd(&s_ntouched, 1)] = v;      // first visit of v
                if (c) {
                    unsigned long long di = d;
                    while (di) {
                        const int i = __ffsll((long long)di) - 1;
                        di &= di - 1;
                        unsigned long long cj = c;
                        while (cj) {
                            const int j = __ffsll((long long)cj) - 1;
                            cj &= cj - 1;
                            record(i, j, 2 * level - 1);
                        }
                    }
                }
            }
            __syncthreads();
            // ---- commit pass B: bits that arrive together (distance 2L), after every 2L - 1 meeting is known
            if (s_pairs < want) {
                for (int f = threadIdx.x; f < nn; f += NT) {
                    unsigned long long di = sc.dl[cb ^ 1][f];
                    if (di & (di - 1)) {                                       // at least two new bits
                        while (di) {
                            const int i = __ffsll((long long)di) - 1;
                            di &= di - 1;
                            unsigned long long dj = di;
                            while (dj) {
                                const int j = __ffsll((long long)dj) - 1;
                                dj &= dj - 1;
                                record(i, j, 2 * level);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x < nt && s_nfound[threadIdx.x] >= nt - 1) atomicAnd(&s_active, ~(1ull << threadIdx.x));
            if (threadIdx.x == 0) s_nf[cb] = 0;
            cb ^= 1;
            const bool done = s_pairs >= want || nn == 0;
            __syncthreads();
            if (done) break;
        }
        // ---- clean the scratch for the next target
        const int ntouched = s_ntouched;
        for (int f = threadIdx.x; f < ntouched; f += NT) {
            const int v = sc.touched[f];
            sc.cur[v] = 0ull;
            sc.acc[v] = 0ull;
        }
        if (threadIdx.x == 0) {
            out_sum[t] = 2 * (long long)s_sum;          // ordered-pair convention of the interface
            out_pairs[t] = 2 * s_pairs;
            out_nmapped[t] = nt;
            s_nf[0] = 0;
            s_nf[1] = 0;
        }
        __syncthreads();
    }
}

}  // namespace spec

static int spec_blocks(int m, size_t workspace_bytes, int n_query) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long b = (long long)(workspace_bytes / spec::scratch_bytes(m));
    if (b > 4LL * sms) b = 4LL * sms;
    if (b > n_query) b = n_query;
    return (int)b;
}

extern "C" size_t nabo_specificity_workspace_bytes(int n_ref, int n_query) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long b = 4LL * sms;
    if (b > n_query) b = n_query;
    if (b < 1) b = 1;
    // cap the request at 8 GB: fewer CTAs are used for very large graphs
    size_t need = (size_t)b * spec::scratch_bytes(n_ref);
    const size_t cap = (size_t)8 << 30;
    if (need > cap) {
        need = cap;
        if (need < spec::scratch_bytes(n_ref)) need = spec::scratch_bytes(n_ref);
    }
    return need;
}

extern "C" int nabo_mapping_specificity(const long long* indptr, const int32_t* indices, int n_ref,
                                        const int32_t* tgt_knn, const uint8_t* counts, int n_query, int k,
                                        long long* out_sum, int32_t* out_pairs, int32_t* out_nmapped,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    NABO_ARG(n_ref > 0 && n_query >= 0 && k > 0, "mapping_specificity: bad sizes");
    NABO_ARG(k <= spec::MAXK, "mapping_specificity: at most %d mapped cells per target", spec::MAXK);
    NABO_ARG(indptr && indices && tgt_knn && counts && out_sum && out_pairs && out_nmapped, "mapping_specificity: null pointer");
    if (n_query == 0) return 0;
    const int blocks = spec_blocks(n_ref, workspace_bytes, n_query);
    if (workspace == nullptr || blocks < 1)
        return nabo_set_error(NABO_EWORKSPACE, "mapping_specificity: workspace too small (need %zu bytes per CTA)",
                              spec::scratch_bytes(n_ref));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per_block = spec::scratch_bytes(n_ref);
    NABO_CUDA(cudaMemsetAsync(workspace, 0, (size_t)blocks * per_block, st));
    spec::specificity_kernel<<<blocks, spec::NT, 0, st>>>(indptr, indices, n_ref, tgt_knn, counts, n_query, k, out_sum,
                                                          out_pairs, out_nmapped, (unsigned char*)workspace, per_block);
    NABO_LAUNCH_CHECK("specificity_kernel");
    return 0;
}
