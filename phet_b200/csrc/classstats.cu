// Whole-vector statistics per gene: exact order statistics (for the IQR), mean and sum of squared
// deviations of the control cells, the case cells and all cells of every gene.  This is what the two
// sibling IQR scorers of the reference need:
//
//   HIQR           src/model/hvf.py:59-96          iqr(X[:, g]) or sum over classes of iqr(X[class, g])
//   DeltaIQRMean   src/model/deltaiqrmean.py:53-85 iqr / mean / ttest_ind of class 0 against class 1
//
// They are the S = 1, m = class-size corner of PHet's Step 1, where sorting networks and the
// thread-per-sample select do not apply (m is 10^4 .. 10^5); the kernel is a block-wide exact MSB radix
// select instead.
//
// One CTA per gene.  The gene's row of the gene-major, class-grouped matrix ([control | case]) is read
// from global memory with coalesced loads; after the first pass it sits in L2/L1.  Six selections run at
// once ("tracks": low and high percentile of control, case and pooled); a pass consumes 4 bits of the
// order-preserving key of every element.  Every thread counts the digits of its elements in private
// 16-bit counters (6 tracks x 16 digits, words interleaved by lane: bank == lane, conflict-free, no
// atomics, deterministic); the CTA reduces them (REDUX + one SMEM hop) and one thread per track fixes the
// next digit (the general path, 8 / 16 passes: always exact).  The fast path reads the row twice: pass A drops every
// cell into one of 256 value buckets (a linear map of the range of a 2 x 512 cell sample -- the cells of a class sit
// in a fixed pseudo-random order, so the first 512 are a random sample -- shared-memory atomics, two histograms) and
// gathers the moments in fp64 (deviations from the first cell of the class, fixed-order sums); the bucket holding
// each of the six ranks follows from the histograms; pass B compacts the cells of those buckets (about 1 % of the
// vector each) into a shared-memory list per track (warp-aggregated appends), and the order statistic y[j] and
// y[j+1] (the smallest value above y[j], or y[j] itself when its tie group reaches rank j+1) come from the lists
// by radix select, one warp per track.  The bucket map is monotone in the value, so buckets partition the order
// statistics exactly; a list that overflows (a tie group of thousands of equal values, a heavy-tailed gene whose
// sample range is huge) sends the gene to the general path.
#include <type_traits>

#include "common.cuh"
#include "warp_prims.cuh"

namespace phet {

constexpr int kCsThreads = 512;
constexpr int kCsTracks = 6;   // (class 0, class 1, pooled) x (low, high)

template <typename T> struct CsKey;
template <> struct CsKey<float> {
    typedef uint32_t type;
    static constexpr int bits = 32;
    static __device__ __forceinline__ uint32_t of(float v) {
        const uint32_t u = __float_as_uint(v);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
    static __device__ __forceinline__ float back(uint32_t k) {
        return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
    }
};
template <> struct CsKey<double> {
    typedef unsigned long long type;
    static constexpr int bits = 64;
    static __device__ __forceinline__ unsigned long long of(double v) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(v);
        return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    }
    static __device__ __forceinline__ double back(unsigned long long k) {
        return __longlong_as_double((long long)((k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k));
    }
};

// (1 - g) * y[j] + g * y[j+1] and the difference of the two percentiles, in X's dtype and scipy's operation
// order (scipy.stats._quantile._quantile_hf last line; iqr = pct[1] - pct[0]), no contraction.
__device__ __forceinline__ double cs_iqr(float yjl, float ykl, float yjh, float ykh, double g_lo, double g_hi) {
    const float gl = (float)g_lo, gh = (float)g_hi;
    const float ql = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, gl), yjl), __fmul_rn(gl, ykl));
    const float qh = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, gh), yjh), __fmul_rn(gh, ykh));
    return (double)__fsub_rn(qh, ql);
}
__device__ __forceinline__ double cs_iqr(double yjl, double ykl, double yjh, double ykh, double g_lo, double g_hi) {
    const double ql = __dadd_rn(__dmul_rn(__dsub_rn(1.0, g_lo), yjl), __dmul_rn(g_lo, ykl));
    const double qh = __dadd_rn(__dmul_rn(__dsub_rn(1.0, g_hi), yjh), __dmul_rn(g_hi, ykh));
    return __dsub_rn(qh, ql);
}

// 16-byte loads of the row: four times (float) / twice (double) the bytes in flight per thread -- the two reads of
// the row are bound by L2 / HBM latency at one CTA of 512 threads per SM, not by bandwidth
template <typename T> struct CsVec;
template <> struct CsVec<float> {
    typedef float4 type;
    static constexpr int width = 4;
    static __device__ __forceinline__ void get(const float4 &v, float *o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};
template <> struct CsVec<double> {
    typedef double2 type;
    static constexpr int width = 2;
    static __device__ __forceinline__ void get(const double2 &v, double *o) { o[0] = v.x; o[1] = v.y; }
};

struct ClassStatsArgs {
    const void *xg;
    int64_t n_pad, xg_g0, g_begin, g_end, n0, n1;
    phet_quantile_pos q[3];   // positions for n0, n1 and n0 + n1 elements
    double *out;              // [7][G]: iqr0, iqr1, iqr_all, mean0, mean1, ss0, ss1
    int64_t G;
    // two launches: the FAST instantiation (bucket selection only: small enough for two CTAs per SM) appends the genes
    // whose candidate lists overflow to ovf_list; the general instantiation then takes exactly those (gene_list)
    int *ovf_count, *ovf_list;
    const int *gene_count, *gene_list;
};

// fixed-order block sum of one double per thread; result broadcast to all threads
__device__ __forceinline__ double cs_block_sum(double v, double *red, int tid) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kCsThreads / 32; ++w) t += red[w];
    return t;
}

template <typename T, bool FAST> struct CsList { static constexpr int cap = FAST ? 3072 : 4096; };   // keys per track kept in shared memory
template <bool FAST> struct CsList<double, FAST> { static constexpr int cap = FAST ? 1536 : 2048; };
constexpr int kCsBuckets = 256;    // value buckets of the first pass (a linear map of the range of a 2 x 512 cell sample)

template <typename T, bool FAST>
__global__ void __launch_bounds__(kCsThreads, FAST ? 2 : 1) k_class_stats(const ClassStatsArgs a) {
    typedef typename CsKey<T>::type K;
    constexpr int NPASS = CsKey<T>::bits / 4;
    constexpr int LCAP = CsList<T, FAST>::cap;
    constexpr int CW = FAST ? 8 : kCsTracks * 8;                 // private counter words per thread (FAST: the list select only)
    extern __shared__ __align__(16) uint32_t cnt[];              // [warp][word][lane], two 16-bit counters per word (96 KB; FAST: 16 KB)
    K *lists = reinterpret_cast<K *>(cnt + CW * kCsThreads);   // [track][LCAP] candidate keys
    __shared__ uint32_t wpart[kCsThreads / 32][kCsTracks * 16]; // per-warp digit counts
    __shared__ double red[kCsThreads / 32];
    __shared__ K prefix[kCsTracks];
    __shared__ int64_t rank[kCsTracks];
    __shared__ unsigned long long cle[kCsTracks];   // elements <= y[j] (of the row, or of the candidate list)
    __shared__ K kmin[kCsTracks];                   // smallest key above y[j]
    __shared__ K nmin[kCsTracks];                   // smallest key of the vector above the candidates' 12-bit bucket
    __shared__ int llen[kCsTracks];                 // candidates appended (may exceed LCAP: overflow -> full passes)
    __shared__ unsigned int hist[2][kCsBuckets + 2];   // per class: cells per value bucket (below, 0 .. kCsBuckets-1, above)
    __shared__ int bsel[kCsTracks];                 // bucket (index into hist) that holds the track's rank
    __shared__ T bmap[2];                           // bucket = floor((v - bmap[0]) * bmap[1]), in the storage type
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *mine = cnt + (size_t)warp * (CW * 32) + lane;   // word w of this thread at mine[32 * w]
    const int64_t n0 = a.n0, n = a.n0 + a.n1;
    const int64_t nvec[3] = {a.n0, a.n1, n};

    for (int w = 0; w < CW; ++w) mine[32 * w] = 0u;
    for (int i = tid; i < 2 * (kCsBuckets + 2); i += kCsThreads) (&hist[0][0])[i] = 0u;

    const int64_t n_genes = a.gene_list ? (int64_t)*a.gene_count : a.g_end - a.g_begin;
    for (int64_t gi = blockIdx.x; gi < n_genes; gi += gridDim.x) {
        const int64_t g = a.gene_list ? (int64_t)a.gene_list[gi] : a.g_begin + gi;
        const T *row = static_cast<const T *>(a.xg) + (size_t)(g - a.xg_g0) * (size_t)a.n_pad;
        if (tid < kCsTracks) {
            prefix[tid] = 0;
            const phet_quantile_pos &q = a.q[tid >> 1];
            rank[tid] = (tid & 1) ? q.j_hi : q.j_lo;
            llen[tid] = 0;
            nmin[tid] = ~K(0);
        }
        __syncthreads();

        // CTA totals of the private counters, then one thread per track fixes the next digit
        auto pick_digit = [&](int shift) {
            if (FAST) return;
            // (a thread sees at most n / 512 + 1 elements: no 16-bit overflow for n < 2^25)
            for (int w = 0; w < kCsTracks * 8; ++w) {
                const uint32_t x = mine[32 * w];
                mine[32 * w] = 0u;
                const uint32_t lo = __reduce_add_sync(kFull, x & 0xFFFFu), hi = __reduce_add_sync(kFull, x >> 16);
                if (lane == 0) {
                    wpart[warp][2 * w] = lo;
                    wpart[warp][2 * w + 1] = hi;
                }
            }
            __syncthreads();
            if (tid < kCsTracks) {
                int64_t r = rank[tid], cum = 0;
                int d = 0;
                for (; d < 16; ++d) {
                    int64_t cd = 0;
                    for (int w = 0; w < kCsThreads / 32; ++w) cd += wpart[w][tid * 16 + d];
                    if (cum + cd > r || d == 15) break;
                    cum += cd;
                }
                rank[tid] = r - cum;
                prefix[tid] |= (K)d << shift;
            }
            __syncthreads();
        };
        // one digit from a pass over the whole row; pass 0 also gathers the moments (deviations from the
        // first cell of each class, fp64: exact zero variance for a constant vector, no cancellation trouble)
        const double K0 = (double)__ldg(row), K1 = (double)__ldg(row + n0);
        double s0 = 0.0, s1 = 0.0, d0 = 0.0, d1 = 0.0;
        auto bucket_of = [&](T v) {   // monotone non-decreasing in the value: buckets partition the order statistics
            T tb = floor((v - bmap[0]) * bmap[1]);   // (each rounding step is monotone)
            tb = tmin(tmax(tb, T(-1)), T(kCsBuckets));
            return (int)tb + 1;
        };
        auto full_pass = [&](int pass) {
            if (FAST) return;
            const int shift = CsKey<T>::bits - 4 * (pass + 1);
            K pf[kCsTracks];
#pragma unroll
            for (int t = 0; t < kCsTracks; ++t) pf[t] = prefix[t];
#pragma unroll 4
            for (int64_t i = tid; i < n; i += kCsThreads) {
                const T v = __ldg(row + i);
                const K key = CsKey<T>::of(v);
                const uint32_t dig = (uint32_t)(key >> shift) & 15u;
                const int c = (i < n0) ? 0 : 1;
                if (pass == 0) {
                    // nothing is fixed yet: the low and the high track of a vector see the same digits
                    const uint32_t inc = (dig & 1u) ? 0x10000u : 1u;
                    mine[32 * ((2 * c) * 8 + (dig >> 1))] += inc;
                    mine[32 * (4 * 8 + (dig >> 1))] += inc;
                } else {
                    const K hi = key >> (shift + 4);   // digits already fixed
#pragma unroll
                    for (int t = 0; t < kCsTracks; ++t) {
                        const bool member = (t >> 1) == 2 || (t >> 1) == c;
                        if (member && hi == (pf[t] >> (shift + 4))) mine[32 * (t * 8 + (dig >> 1))] += (dig & 1u) ? 0x10000u : 1u;
                    }
                }
            }
            if (pass == 0) {   // copy the shared histograms to the high tracks
                for (int w = 0; w < 8; ++w) {
                    mine[32 * (1 * 8 + w)] = mine[32 * (0 * 8 + w)];
                    mine[32 * (3 * 8 + w)] = mine[32 * (2 * 8 + w)];
                    mine[32 * (5 * 8 + w)] = mine[32 * (4 * 8 + w)];
                }
            }
            pick_digit(shift);
        };

        // ---- pass A: value buckets.  The cells of a class sit in a fixed pseudo-random order (phet_set_classes), so
        //      the first 512 of each class are a random sample: its range, cut into kCsBuckets equal parts, is the map.
        {
            double mn_s = CUDART_INF, mx_s = -CUDART_INF;
            if (tid < n0) { const double v = (double)__ldg(row + tid); mn_s = fmin(mn_s, v); mx_s = fmax(mx_s, v); }
            if (tid < a.n1) { const double v = (double)__ldg(row + n0 + tid); mn_s = fmin(mn_s, v); mx_s = fmax(mx_s, v); }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                mn_s = fmin(mn_s, __shfl_xor_sync(kFull, mn_s, o));
                mx_s = fmax(mx_s, __shfl_xor_sync(kFull, mx_s, o));
            }
            __syncthreads();
            if (lane == 0) {
                red[warp] = mn_s;
                wpart[warp][0] = __double2hiint(mx_s);
                wpart[warp][1] = __double2loint(mx_s);
            }
            __syncthreads();
            if (tid == 0) {
                double lo = CUDART_INF, hi = -CUDART_INF;
                for (int w = 0; w < kCsThreads / 32; ++w) {
                    lo = fmin(lo, red[w]);
                    hi = fmax(hi, __hiloint2double((int)wpart[w][0], (int)wpart[w][1]));
                }
                const T scale = (hi > lo) ? (T)((double)kCsBuckets / (hi - lo)) : T(0);
                bmap[0] = (T)lo;
                bmap[1] = isfinite(scale) ? scale : T(0);
            }
            __syncthreads();
        }
        constexpr int VW = CsVec<T>::width;
        const typename CsVec<T>::type *rowv = reinterpret_cast<const typename CsVec<T>::type *>(row);   // rows are 16-byte aligned, n_pad covers the tail
        const int64_t nvec16 = (n + VW - 1) / VW;
        // (32-bit cell numbers in the two passes over the row: phet_class_stats admits at most 2^25 - 1 cells)
        const int n_i = (int)n, n0_i = (int)n0, nvec_i = (int)nvec16;
#pragma unroll 4
        for (int iv = tid; iv < nvec_i; iv += kCsThreads) {
            T e[VW];
            CsVec<T>::get(__ldg(rowv + iv), e);
#pragma unroll
            for (int k = 0; k < VW; ++k) {
                const int i = iv * VW + k;
                if (i < n_i) {
                    const T v = e[k];
                    const int c = (i < n0_i) ? 0 : 1;
                    const double d = (double)v - (c ? K1 : K0);
                    if (c) { s1 += d; d1 = fma(d, d, d1); } else { s0 += d; d0 = fma(d, d, d0); }
                    atomicAdd(&hist[c][bucket_of(v)], 1u);
                }
            }
        }
        __syncthreads();
        if (warp < kCsTracks) {   // one warp per track: the bucket that holds rank j, and the rank inside it
            constexpr int PER = (kCsBuckets + 2 + 31) / 32;
            const int vt = warp >> 1;
            const int64_t r = rank[warp];
            unsigned int cb[PER];
            unsigned int local = 0;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                const int b = lane * PER + q;
                unsigned int x = 0;
                if (b < kCsBuckets + 2) x = (vt == 0) ? hist[0][b] : (vt == 1) ? hist[1][b] : hist[0][b] + hist[1][b];
                cb[q] = x;
                local += x;
            }
            unsigned int incl = local;   // inclusive scan over the lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int up = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += up;
            }
            const int64_t before = (int64_t)incl - local;
            const bool mine_has = before <= r && r < (int64_t)incl;
            const unsigned owner = __ballot_sync(kFull, mine_has);
            if (owner == 0u) {   // r beyond the vector (cannot happen for validated positions): last bucket
                if (lane == 0) { bsel[warp] = kCsBuckets + 1; rank[warp] = 0; }
            } else if (mine_has) {
                int64_t cum = before;
                int q = 0;
#pragma unroll
                for (; q < PER - 1; ++q) {
                    if (cum + cb[q] > r) break;
                    cum += cb[q];
                }
                bsel[warp] = lane * PER + q;
                rank[warp] = r - cum;
            }
        }
        __syncthreads();
        for (int i = tid; i < 2 * (kCsBuckets + 2); i += kCsThreads) (&hist[0][0])[i] = 0u;   // for the next gene
        const int next_pass = 0;   // the lists resolve every digit
        const double sum0 = cs_block_sum(s0, red, tid), sum1 = cs_block_sum(s1, red, tid);
        const double sq0 = cs_block_sum(d0, red, tid), sq1 = cs_block_sum(d1, red, tid);
        const double mean0 = K0 + sum0 / (double)a.n0, mean1 = K1 + sum1 / (double)a.n1;
        double ss0 = sq0 - sum0 * sum0 / (double)a.n0, ss1 = sq1 - sum1 * sum1 / (double)a.n1;
        if (ss0 < 0.0) ss0 = 0.0;
        if (ss1 < 0.0) ss1 = 0.0;

        // ---- pass B, compaction: the cells of a track's bucket go to the track's list
        {
            int pf[kCsTracks];
            K nm[kCsTracks];
#pragma unroll
            for (int t = 0; t < kCsTracks; ++t) {
                pf[t] = bsel[t];
                nm[t] = ~K(0);
            }
            // Trips whose cells all belong to one class (all but the one or two around n0) only look at that class's two
            // tracks and the two pooled ones, and a trip in which no lane matches any of its four buckets (most trips: a
            // bucket holds ~0.4 % of the cells) costs one ballot instead of four.  CL = 0 / 1: the class; 2: mixed trip.
            auto trips = [&](auto cl_tag, int bv0, int bv1) {
                constexpr int CL = decltype(cl_tag)::value;
                typename CsVec<T>::type nxt = {};   // the next trip's cells are in flight while this trip's are compacted
                if (bv0 < bv1 && bv0 + tid < nvec_i) nxt = __ldg(rowv + bv0 + tid);
                for (int bv = bv0; bv < bv1; bv += kCsThreads) {   // warp-uniform trip count (ballots inside)
                    const int iv = bv + tid;
                    T e[VW];
                    CsVec<T>::get(nxt, e);
                    if (bv + kCsThreads < bv1 && iv + kCsThreads < nvec_i) nxt = __ldg(rowv + iv + kCsThreads);
#pragma unroll
                    for (int k = 0; k < VW; ++k) {
                        const int i = iv * VW + k;
                        const bool valid = iv < nvec_i && i < n_i;
                        const T v = e[k];
                        const K key = CsKey<T>::of(v);
                        const int kp = bucket_of(v);
                        const int c = (CL == 2) ? ((i < n0_i) ? 0 : 1) : CL;
                        bool any = false;
#pragma unroll
                        for (int t = 0; t < kCsTracks; ++t) {
                            if (CL != 2 && (t >> 1) != 2 && (t >> 1) != CL) continue;   // the other class's tracks
                            const bool member = valid && ((t >> 1) == 2 || (t >> 1) == c);
                            any = any || (member && kp == pf[t]);
                            if (member && kp > pf[t] && key < nm[t]) nm[t] = key;
                        }
                        if (__ballot_sync(kFull, any) == 0u) continue;
#pragma unroll
                        for (int t = 0; t < kCsTracks; ++t) {
                            if (CL != 2 && (t >> 1) != 2 && (t >> 1) != CL) continue;
                            const bool member = valid && ((t >> 1) == 2 || (t >> 1) == c);
                            const bool match = member && kp == pf[t];
                            const unsigned mask = __ballot_sync(kFull, match);
                            if (mask) {
                                int base = 0;
                                if (lane == (__ffs((int)mask) - 1)) base = atomicAdd(&llen[t], __popc(mask));
                                base = __shfl_sync(kFull, base, __ffs((int)mask) - 1);
                                const int pos = base + __popc(mask & ((1u << lane) - 1u));
                                if (match && pos < LCAP) lists[t * LCAP + pos] = key;
                            }
                        }
                    }
                }
            };
            {
                const int per_trip = kCsThreads;                                   // vectors per trip
                const int t0 = (n0_i / VW) / per_trip * per_trip;                  // trips [0, t0): class 0 only
                int t1 = ((n0_i + VW - 1) / VW + per_trip - 1) / per_trip * per_trip;   // trips from t1 on: class 1 only
                if (t1 > nvec_i) t1 = (nvec_i + per_trip - 1) / per_trip * per_trip;
                trips(std::integral_constant<int, 0>{}, 0, t0);
                trips(std::integral_constant<int, 2>{}, t0, t1);
                trips(std::integral_constant<int, 1>{}, t1, nvec_i);
            }
#pragma unroll
            for (int t = 0; t < kCsTracks; ++t) {
                K mk = nm[t];
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    const K other = __shfl_xor_sync(kFull, mk, o);
                    mk = other < mk ? other : mk;
                }
                if (lane == 0) atomicMin(&nmin[t], mk);
            }
            __syncthreads();
        }
        bool overflow = false;
#pragma unroll
        for (int t = 0; t < kCsTracks; ++t) overflow = overflow || llen[t] > LCAP;   // CTA-uniform

        if (tid < kCsTracks) {
            cle[tid] = 0ull;
            kmin[tid] = ~K(0);
        }
        unsigned long long le[kCsTracks];
        K mn[kCsTracks];
#pragma unroll
        for (int t = 0; t < kCsTracks; ++t) {
            le[t] = 0ull;
            mn[t] = ~K(0);
        }
        int64_t r0 = 0;   // (threads < kCsTracks) rank of y[j] inside the candidate list
        if (!overflow) {
            if (tid < kCsTracks) r0 = rank[tid];
            __syncthreads();   // r0 is read before the track warps move rank[] on
            // ---- remaining digits and y[j+1] from the lists: ONE WARP PER TRACK, no CTA-wide step.  A lane counts
            //      the digits of its share of the list in its private counters; 16 REDUX give the warp the digit.
            if (warp < kCsTracks) {
                const int t = warp;
                const int len = llen[t];
                const K *lt = lists + t * LCAP;
                K pk = prefix[t];
                int64_t r = rank[t];
                for (int pass = next_pass; pass < NPASS; ++pass) {
                    const int shift = CsKey<T>::bits - 4 * (pass + 1);
                    const K fixed = (pass == 0) ? K(0) : (~K(0) << (shift + 4));   // digits already chosen
                    for (int e = lane; e < len; e += 32) {
                        const K key = lt[e];
                        const uint32_t dig = (uint32_t)(key >> shift) & 15u;
                        if (((key ^ pk) & fixed) == 0) mine[32 * (dig >> 1)] += (dig & 1u) ? 0x10000u : 1u;
                    }
                    int64_t cum = 0;
                    int d = 0;
                    bool done = false;
#pragma unroll
                    for (int w = 0; w < 8; ++w) {
                        const uint32_t x = mine[32 * w];
                        mine[32 * w] = 0u;
                        const int64_t c0 = __reduce_add_sync(kFull, x & 0xFFFFu), c1 = __reduce_add_sync(kFull, x >> 16);
                        if (!done) {
                            if (cum + c0 > r) { d = 2 * w; done = true; }
                            else {
                                cum += c0;
                                if (cum + c1 > r || w == 7) { d = 2 * w + 1; done = true; }
                                else cum += c1;
                            }
                        }
                    }
                    r -= cum;
                    pk |= (K)d << shift;
                }
                unsigned long long l = 0ull;
                K mk = ~K(0);
                for (int e = lane; e < len; e += 32) {
                    const K key = lt[e];
                    if (key <= pk) ++l;
                    else if (key < mk) mk = key;
                }
#pragma unroll
                for (int tt = 0; tt < kCsTracks; ++tt) {   // (static indices: le[] / mn[] stay in registers)
                    if (tt == t) {
                        le[tt] = l;
                        mn[tt] = mk;
                    }
                }
                if (lane == 0) prefix[t] = pk;
            }
        } else if (FAST) {
            // ---- this instantiation has no general path: the gene goes to the second launch
            if (tid == 0) a.ovf_list[atomicAdd(a.ovf_count, 1)] = (int)g;
            __syncthreads();
            continue;   // (CTA-uniform)
        } else {
            // ---- a tie group (or a very dense bucket) larger than the list: exact radix select over the whole row,
            //      4 bits per pass, from the original ranks
            if (tid < kCsTracks) {
                prefix[tid] = 0;
                const phet_quantile_pos &q = a.q[tid >> 1];
                rank[tid] = (tid & 1) ? q.j_hi : q.j_lo;
            }
            __syncthreads();
            if (!FAST)   // (the full passes use 48 counter words per thread; the FAST instantiation has 8)
                for (int pass = 0; pass < NPASS; ++pass) full_pass(pass);
            K pf[kCsTracks];
#pragma unroll
            for (int t = 0; t < kCsTracks; ++t) pf[t] = prefix[t];
            for (int64_t i = tid; i < n; i += kCsThreads) {
                const K key = CsKey<T>::of(__ldg(row + i));
                const int c = (i < n0) ? 0 : 1;
#pragma unroll
                for (int t = 0; t < kCsTracks; ++t) {
                    const bool member = (t >> 1) == 2 || (t >> 1) == c;
                    if (member) {
                        if (key <= pf[t]) ++le[t];
                        else if (key < mn[t]) mn[t] = key;
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < kCsTracks; ++t) {
            unsigned long long l = le[t];
            K mk = mn[t];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                l += __shfl_xor_sync(kFull, l, o);
                const K other = __shfl_xor_sync(kFull, mk, o);
                mk = other < mk ? other : mk;
            }
            if (lane == 0) {
                atomicAdd(&cle[t], l);       // integer sum / min: order-independent, deterministic
                atomicMin(&kmin[t], mk);
            }
        }
        __syncthreads();
        if (tid < kCsTracks) {
            // y[j+1] = y[j] when the tie group of y[j] reaches rank j+1, else the next larger key
            const phet_quantile_pos &q = a.q[tid >> 1];
            const int64_t j = (tid & 1) ? q.j_hi : q.j_lo, k = (tid & 1) ? q.k_hi : q.k_lo;
            const int64_t need = overflow ? j + 2 : r0 + 2;
            K kk = prefix[tid];
            if (k > j && (int64_t)cle[tid] < need) {
                kk = kmin[tid];
                if (!overflow && nmin[tid] < kk) kk = nmin[tid];
            }
            kmin[tid] = kk;   // key of y[k]
        }
        __syncthreads();
        if (tid < 3) {
            const phet_quantile_pos &q = a.q[tid];
            const T yjl = CsKey<T>::back(prefix[2 * tid]), ykl = CsKey<T>::back(kmin[2 * tid]);
            const T yjh = CsKey<T>::back(prefix[2 * tid + 1]), ykh = CsKey<T>::back(kmin[2 * tid + 1]);
            a.out[(size_t)tid * a.G + g] = (nvec[tid] > 0) ? cs_iqr(yjl, ykl, yjh, ykh, q.g_lo, q.g_hi) : 0.0;
        }
        if (tid == 0) {
            a.out[(size_t)3 * a.G + g] = mean0;
            a.out[(size_t)4 * a.G + g] = mean1;
            a.out[(size_t)5 * a.G + g] = ss0;
            a.out[(size_t)6 * a.G + g] = ss1;
        }
        __syncthreads();
    }
}

template <typename T>
static int launch_class_stats_T(phet_ctx *c, ClassStatsArgs a, int64_t ng) {
    typedef typename CsKey<T>::type K;
    // launch 1: bucket selection only, two CTAs per SM; genes whose lists overflow are listed for launch 2
    if (int e = c->cs_ovf.alloc((size_t)c->G + 1)) return e;
    PHET_CUDA(cudaMemsetAsync(c->cs_ovf.p, 0, sizeof(int), c->stream));
    a.ovf_count = c->cs_ovf.p;
    a.ovf_list = c->cs_ovf.p + 1;
    a.gene_count = nullptr;
    a.gene_list = nullptr;
    const int smem_fast = 8 * kCsThreads * (int)sizeof(uint32_t) + kCsTracks * CsList<T, true>::cap * (int)sizeof(K);
    PHET_CUDA(cudaFuncSetAttribute(k_class_stats<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_fast));
    int per_sm = 0;
    PHET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_class_stats<T, true>, kCsThreads, smem_fast));
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)per_sm * c->num_sms;
    if (grid > ng) grid = ng;
    k_class_stats<T, true><<<(unsigned)grid, kCsThreads, smem_fast, c->stream>>>(a);
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    // launch 2: the general instantiation on the listed genes (it reads their number on the device: no host round trip)
    a.gene_count = c->cs_ovf.p;
    a.gene_list = c->cs_ovf.p + 1;
    const int smem = kCsTracks * 8 * kCsThreads * (int)sizeof(uint32_t) + kCsTracks * CsList<T, false>::cap * (int)sizeof(K);
    PHET_CUDA(cudaFuncSetAttribute(k_class_stats<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    grid = c->num_sms < ng ? c->num_sms : ng;
    k_class_stats<T, false><<<(unsigned)grid, kCsThreads, smem, c->stream>>>(a);
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    return PHET_OK;
}

int launch_class_stats(phet_ctx *c, int64_t g_begin, int64_t g_end, const phet_quantile_pos *q3, double *out_dev) {
    ClassStatsArgs a;
    a.xg = c->xg.p;
    a.n_pad = c->n_pad;
    a.xg_g0 = c->xg_g0;
    a.g_begin = g_begin;
    a.g_end = g_end;
    a.n0 = c->n0;
    a.n1 = c->n1;
    for (int i = 0; i < 3; ++i) a.q[i] = q3[i];
    a.out = out_dev;
    a.G = c->G;
    const int64_t ng = g_end - g_begin;
    if (ng < 1) return PHET_OK;
    return c->storage == PHET_F32 ? launch_class_stats_T<float>(c, a, ng) : launch_class_stats_T<double>(c, a, ng);
}

}  // namespace phet
