// Warp-level primitives for the subsample kernels: register-resident sorting network,
// butterfly reductions, and 1-D TMA bulk staging of a gene vector into shared memory.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace phet {

constexpr unsigned kFull = 0xffffffffu;

template <typename T> struct Lim;
template <> struct Lim<float> {
    static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
};
template <> struct Lim<double> {
    static __device__ __forceinline__ double inf() { return CUDART_INF; }
};

__device__ __forceinline__ float tmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float tmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double tmin(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ double tmax(double a, double b) { return fmax(a, b); }

constexpr int next_pow2(int e) {
    int p = 1;
    while (p < e) p <<= 1;
    return p;
}

// Sort 32*E values held as v[r] in lane `lane` (global position p = 32*r + lane) ascending in p.
// Flip-form bitonic network (every comparator points the same way) on 32*EP positions,
// EP = next_pow2(E); registers r >= E are virtual +inf pads and every comparator touching one
// is a no-op, so it is pruned at compile time.  Comparators with partner distance < 32 are
// shuffles (SHFL + FMNMX), the rest are register-to-register min/max.
template <typename T, int E>
__device__ __forceinline__ void warp_sort_asc(T (&v)[E], const int lane) {
    constexpr int EP = next_pow2(E);
    constexpr int NTOT = 32 * EP;
#pragma unroll
    for (int k = 2; k <= NTOT; k <<= 1) {
        const int x = k - 1;
        const int xr = x >> 5;
        if (xr == 0) {
            const bool lower = (lane & (k >> 1)) == 0;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const T o = __shfl_xor_sync(kFull, v[r], x);
                v[r] = lower ? tmin(v[r], o) : tmax(v[r], o);
            }
        } else {
            T nv[E];
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int rp = r ^ xr;
                if (rp < E) {
                    const T o = __shfl_xor_sync(kFull, v[rp], 31);
                    nv[r] = (r < rp) ? tmin(v[r], o) : tmax(v[r], o);
                } else {
                    nv[r] = v[r];
                }
            }
#pragma unroll
            for (int r = 0; r < E; ++r) v[r] = nv[r];
        }
#pragma unroll
        for (int j = k >> 2; j >= 1; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const int rp = r ^ jr;
                    if (r < rp && rp < E) {
                        const T lo = tmin(v[r], v[rp]);
                        const T hi = tmax(v[r], v[rp]);
                        v[r] = lo;
                        v[rp] = hi;
                    }
                }
            } else {
                const bool lower = (lane & j) == 0;
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const T o = __shfl_xor_sync(kFull, v[r], j);
                    v[r] = lower ? tmin(v[r], o) : tmax(v[r], o);
                }
            }
        }
    }
}

// Value at sorted position p (warp-uniform) of a sorted register array.
template <typename T, int E>
__device__ __forceinline__ T warp_pick(const T (&v)[E], const int p) {
    const int rr = p >> 5;
    T x = v[0];
#pragma unroll
    for (int r = 1; r < E; ++r)
        if (r == rr) x = v[r];
    return __shfl_sync(kFull, x, p & 31);
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    return x;
}

__device__ __forceinline__ double warp_min(double x) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) x = fmin(x, __shfl_xor_sync(kFull, x, o));
    return x;
}

// ---- mbarrier + 1-D bulk async copy (TMA unit, SASS UBLKCP) ---------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// One thread stages `bytes` (multiple of 16, both addresses 16-byte aligned) of a gene row.
__device__ __forceinline__ void stage_row(void *smem_dst, const void *gmem_src, size_t bytes, uint64_t *bar) {
    constexpr unsigned kChunk = 32768;
    mbar_expect_tx(bar, static_cast<unsigned>(bytes));
    size_t off = 0;
    while (off < bytes) {
        const unsigned n = static_cast<unsigned>(bytes - off < kChunk ? bytes - off : kChunk);
        bulk_g2s(static_cast<char *>(smem_dst) + off, static_cast<const char *>(gmem_src) + off, n, bar);
        off += n;
    }
}

}  // namespace phet
