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

// Ask the TMA unit to pull `bytes` (multiple of 16, 16-byte aligned) into L2 without a destination:
// the HBM read of the NEXT vector a CTA will stage then overlaps the work on the current one.
__device__ __forceinline__ void prefetch_row_l2(const void *gmem_src, size_t bytes) {
    constexpr unsigned kChunk = 32768;
    size_t off = 0;
    while (off < bytes) {
        const unsigned n = static_cast<unsigned>(bytes - off < kChunk ? bytes - off : kChunk);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(static_cast<const char *>(gmem_src) + off), "r"(n)
                     : "memory");
        off += n;
    }
}

// =================================================================================================
// Half-warp merge-split sorter (the fast path for m <= 256).
//
// Each half-warp (16 lanes) sorts its own 16*E values; the two groups of a unit (control, case)
// are sorted at the same time by the two halves of one warp.  Every lane first sorts its E
// values in registers (Batcher merge-exchange network, arbitrary E), then the 16 lanes run a
// flip-form bitonic network of 10 stages whose "comparator" is a merge-split of two sorted
// E-blocks: the lower lane keeps the E smallest of the 2E, the upper lane the E largest.
//
// To make that step branch- and predicate-free the blocks are kept in "x-space": a lane whose
// role in the coming stage is "upper" stores the NEGATED block, ascending.  Then for both roles
//     w[i] = min(x[i], -partner_x[i])          (one SHFL + one FMNMX per element, same index i)
// is the kept half, and in both roles w is an increasing-then-decreasing sequence, so one
// pruned half-cleaner network (virtual -inf pads in front) re-sorts it.  Changing role between
// stages is x'[i] = -x[E-1-i], done arithmetically as alpha*x[i] + beta*x[E-1-i] with
// (alpha, beta) in {(1,0),(0,-1)} on the otherwise idle FMA pipe.  Missing elements are +BIG
// sentinels (finite, so 0*BIG == 0).  Versus the full-warp network this needs ~45 % of the
// shuffles and ~70 % of the min/max instructions per unit.
// =================================================================================================

template <typename T> struct Big;
template <> struct Big<float> {
    static __host__ __device__ __forceinline__ float v() { return 3.402823466e+38f; }
};
template <> struct Big<double> {
    static __host__ __device__ __forceinline__ double v() { return 1.7976931348623157e+308; }
};

template <typename T>
__device__ __forceinline__ void cmpex(T &a, T &b) {
    const T lo = tmin(a, b);
    const T hi = tmax(a, b);
    a = lo;
    b = hi;
}

// Batcher's merge-exchange sort (Knuth 5.2.2 Algorithm M) as a compile-time network, any E.
template <typename T, int E>
__device__ __forceinline__ void local_sort_asc(T (&x)[E]) {
#pragma unroll
    for (int p = 1; p < E; p <<= 1) {
#pragma unroll
        for (int k = p; k >= 1; k >>= 1) {
#pragma unroll
            for (int j = k % p; j <= E - 1 - k; j += 2 * k) {
#pragma unroll
                for (int i = 0; i <= k - 1; ++i) {
                    if (i <= E - j - k - 1 && (i + j) / (2 * p) == (i + j + k) / (2 * p)) cmpex(x[i + j], x[i + j + k]);
                }
            }
        }
    }
}

// Sort an increasing-then-decreasing block ascending: half-cleaner network on next_pow2(E)
// positions with the E real values at the top and virtual -inf pads (no-op comparators) below.
template <typename T, int E>
__device__ __forceinline__ void bitonic_cleanup_asc(T (&w)[E]) {
    constexpr int EP = next_pow2(E);
    constexpr int OFF = EP - E;
#pragma unroll
    for (int d = EP / 2; d >= 1; d >>= 1) {
#pragma unroll
        for (int pos = OFF; pos < EP; ++pos) {
            if ((pos & d) == 0) cmpex(w[pos - OFF], w[pos + d - OFF]);
        }
    }
}

// x'[i] = alpha*x[i] + beta*x[E-1-i]  ((1,0): unchanged; (0,-1): negate and reverse).
template <typename T, int E>
__device__ __forceinline__ void role_flip(T (&x)[E], const T alpha, const T beta) {
    T y[E];
#pragma unroll
    for (int i = 0; i < E; ++i) y[i] = fma(beta, x[E - 1 - i], alpha * x[i]);
#pragma unroll
    for (int i = 0; i < E; ++i) x[i] = y[i];
}

// Per-lane role-change coefficients of the 11 transitions of the 16-lane network; only six are
// distinct.  alpha = 1 keeps the block, alpha = 0 (beta = -1) negates and reverses it.
template <typename T>
struct HalfSortCoef {
    T a_1, a_12, a_14, a_42, a_18, a_84;  // alpha for (init|final <-> bit1), (1<->2), (1->4), (4->2), (1->8), (8->4)
    __device__ __forceinline__ explicit HalfSortCoef(int h) {
        const int b1 = h & 1, b2 = (h >> 1) & 1, b4 = (h >> 2) & 1, b8 = (h >> 3) & 1;
        a_1 = b1 ? T(0) : T(1);
        a_12 = (b1 ^ b2) ? T(0) : T(1);
        a_14 = (b1 ^ b4) ? T(0) : T(1);
        a_42 = (b4 ^ b2) ? T(0) : T(1);
        a_18 = (b1 ^ b8) ? T(0) : T(1);
        a_84 = (b8 ^ b4) ? T(0) : T(1);
    }
};

template <typename T, int E>
__device__ __forceinline__ void merge_split_stage(T (&x)[E], const int mask) {
    T w[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const T o = __shfl_xor_sync(kFull, x[i], mask);
        w[i] = tmin(x[i], -o);
    }
    bitonic_cleanup_asc<T, E>(w);
#pragma unroll
    for (int i = 0; i < E; ++i) x[i] = w[i];
}

// In: E unsorted values per lane (pads = Big<T>::v()).  Out: ascending over the half-warp in
// blocked layout, sorted position p = (lane & 15) * E + r.
template <typename T, int E>
__device__ __forceinline__ void halfwarp_sort_asc(T (&x)[E], const HalfSortCoef<T> &c) {
    local_sort_asc<T, E>(x);
#define PHET_STAGE(ALPHA, MASK)                 \
    role_flip<T, E>(x, (ALPHA), (ALPHA)-T(1));  \
    merge_split_stage<T, E>(x, (MASK));
    PHET_STAGE(c.a_1, 1)     // k = 2 : flip  (role bit 1)
    PHET_STAGE(c.a_12, 3)    // k = 4 : flip  (role bit 2)
    PHET_STAGE(c.a_12, 1)    //         j = 1 (role bit 1)
    PHET_STAGE(c.a_14, 7)    // k = 8 : flip  (role bit 4)
    PHET_STAGE(c.a_42, 2)    //         j = 2
    PHET_STAGE(c.a_12, 1)    //         j = 1
    PHET_STAGE(c.a_18, 15)   // k = 16: flip  (role bit 8)
    PHET_STAGE(c.a_84, 4)    //         j = 4
    PHET_STAGE(c.a_42, 2)    //         j = 2
    PHET_STAGE(c.a_12, 1)    //         j = 1
#undef PHET_STAGE
    role_flip<T, E>(x, c.a_1, c.a_1 - T(1));  // back to true values, ascending
}

// Value at sorted position p (uniform) of a half-warp-sorted blocked array: lane hp = p / E of
// each half, register rp = p % E.  Both halves pick at once (shuffle width 16).
// Explicit shared-window loads: one address add + LDS (the generic-pointer path re-derives the
// shared window base for every access).
__device__ __forceinline__ float lds_at(uint32_t base, int idx, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + 4u * (uint32_t)idx));
    return v;
}
__device__ __forceinline__ double lds_at(uint32_t base, int idx, double) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + 8u * (uint32_t)idx));
    return v;
}

__device__ __forceinline__ double half_sum(double x) {
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    return x;
}

}  // namespace phet
