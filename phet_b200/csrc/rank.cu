// Ranking of the per-gene scores on the device.
// Replaces the ordering step of /root/reference/src/utility/utils.py:107-136 (sort_features:
// df.sort_values(by=["score"], ascending=False)) for a (G,) score vector: order[r] = gene at rank r.
//
// Rank by counting: rank(i) = #{ j : s_j > s_i } + #{ j < i : s_j == s_i }.  G^2 integer comparisons (9e8 at
// G = 30k) are a fraction of a millisecond on 148 SMs, need no scratch and no multi-pass scatter, and the result
// is a permutation by construction (the comparison is a strict total order).  Ties keep ascending gene order,
// which is what pandas' reverse / stable argsort / reverse yields (`kind="stable"`); the reference's default
// `kind="quicksort"` leaves the order INSIDE a tie group unspecified, so parity is asserted on tie groups as sets.
// NaN scores rank last (pandas na_position="last"), among themselves by gene index.
#include "common.cuh"

namespace phet {

constexpr int kRankThreads = 256;
constexpr int kRankTile = 2048;   // keys staged per trip (16 KB)

// Order-preserving unsigned key of a double; every NaN maps to 0 (below -inf).
__device__ __forceinline__ unsigned long long rank_key(double v) {
    if (v != v) return 0ull;
    if (v == 0.0) v = 0.0;   // -0.0 == +0.0 in the sort
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void k_rank_keys(const double *__restrict__ s, int64_t G, unsigned long long *__restrict__ key) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) key[g] = rank_key(s[g]);
}

// One thread per gene; the key vector streams through shared memory in tiles that every thread of the CTA
// compares against (broadcast LDS.128: two keys per load).  Tiles wholly before the CTA's genes count
// `>=` (an equal key with a smaller index ranks first), tiles wholly after count `>`, the one or two
// tiles that overlap the CTA's own range take the full predicate.
__global__ void __launch_bounds__(kRankThreads) k_rank_count(const unsigned long long *__restrict__ key,
                                                             const double *__restrict__ s, int64_t G,
                                                             int32_t *__restrict__ order,
                                                             double *__restrict__ sorted) {
    __shared__ __align__(16) unsigned long long tile[kRankTile];
    const int64_t first = (int64_t)blockIdx.x * kRankThreads;
    const int64_t i = first + threadIdx.x;
    const unsigned long long ki = (i < G) ? key[i] : ~0ull;
    unsigned int cnt = 0;
    for (int64_t t0 = 0; t0 < G; t0 += kRankTile) {
        const int len = (int)((G - t0 < kRankTile) ? (G - t0) : kRankTile);
        __syncthreads();
        for (int e = threadIdx.x; e < kRankTile; e += kRankThreads) tile[e] = (e < len) ? key[t0 + e] : 0ull;
        __syncthreads();
        const ulonglong2 *t2 = reinterpret_cast<const ulonglong2 *>(tile);
        if (t0 + len <= first) {   // all j < i
#pragma unroll 8
            for (int e = 0; e < kRankTile / 2; ++e) {
                const ulonglong2 k = t2[e];
                cnt += (k.x >= ki) + (k.y >= ki);
            }
            // the zero padding of a short tile: counted only by a thread whose key is 0 (NaN); the last tile is
            // never wholly before a CTA's genes, so this branch never sees padding
        } else if (t0 >= first + kRankThreads) {   // all j > i
#pragma unroll 8
            for (int e = 0; e < kRankTile / 2; ++e) {
                const ulonglong2 k = t2[e];
                cnt += (k.x > ki) + (k.y > ki);
            }
        } else {
            for (int e = 0; e < len; ++e) {
                const unsigned long long k = tile[e];
                cnt += (k > ki) || (k == ki && t0 + e < i);
            }
        }
    }
    if (i < G) {
        order[cnt] = (int32_t)i;
        sorted[cnt] = s[i];
    }
}

int launch_rank(phet_ctx *c, const double *scores_dev, int64_t G) {
    if (int e = c->rank_keys.alloc((size_t)G)) return e;
    if (int e = c->rank_order.alloc((size_t)G)) return e;
    if (int e = c->rank_sorted.alloc((size_t)G)) return e;
    const unsigned grid = (unsigned)((G + kRankThreads - 1) / kRankThreads);
    k_rank_keys<<<grid, kRankThreads, 0, c->stream>>>(scores_dev, G, c->rank_keys.p);
    k_rank_count<<<grid, kRankThreads, 0, c->stream>>>(c->rank_keys.p, scores_dev, G, c->rank_order.p, c->rank_sorted.p);
    PHET_CUDA(cudaGetLastError());
    c->launches += 2;
    return PHET_OK;
}

}  // namespace phet
