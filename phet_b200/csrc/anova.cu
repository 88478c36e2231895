// Multi-class Step 1: one-way ANOVA p-value per (gene, subsample) over K groups of m
// cells, accumulated into Fisher's sum.  Replaces /root/reference/src/model/phet.py:348-382 with
// group_test="anova": scipy.stats.f_oneway(*subsamples) (_stats_py.py f_oneway, equal_var=True: sums of squares
// about the pooled mean, F = (SSB/(K-1)) / (SSW/(K m - K)), p = special.fdtrc(K-1, K m - K, F)), then
// nan_to_num(p) (phet.py:378) and the Fisher term -2 ln p with p == 0 dropped (phet.py:449-451).
//
// Same shape as the two-class Step 1: a CTA owns a gene, stages its whole cell vector in shared memory once
// (coalesced loads) and serves every subsample from there -- one thread per subsample walks the K*m positions of
// that subsample (transposed index table: consecutive threads read consecutive words) and gathers from SMEM.
// No order statistics here, only moments: per group the sum of deviations from the vector's first cell, and the
// pooled sum of squared deviations; fp32 running sums for float storage (m <= 512 terms each, combined in fp64),
// fp64 throughout for double storage.  The per-gene Fisher sum is reduced in a fixed order (deterministic).
#include "common.cuh"
#include "pvalue.cuh"
#include "warp_prims.cuh"

namespace phet {

constexpr int kAnovaThreads = 1024;

struct AnovaArgs {
    const void *xg;
    int64_t n_pad, xg_g0, g_begin, g_end, N;
    const uint32_t *idxT;   // [K * m][S]: grouped position of element e of group k in subsample s
    const uint32_t *idxU;   // [S][K * m]: the same positions, one subsample's elements adjacent (warp-per-subsample kernels) or NULL
    int64_t S;
    int K, m;
    double dfn, dfd, ln_beta, p_flush;
    double *acc_f;          // [G]
    double *dbgP;           // [G][S] or NULL
};

// original rows [S][K][m] -> grouped positions, transposed to [K*m][S]
__global__ void k_anova_index(const int32_t *__restrict__ idx, int64_t S, int64_t KM, const int32_t *__restrict__ pos_of_row,
                              int64_t N, uint32_t *__restrict__ idxT, uint32_t *__restrict__ idxU, int *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over [K*m][S]
    if (i >= S * KM) return;
    const int64_t e = i / S, s = i - e * S;
    const int32_t r = idx[s * KM + e];
    uint32_t pos = 0u;
    if (r < 0 || r >= N) *bad = 1;
    else pos = (uint32_t)pos_of_row[r];
    idxT[i] = pos;
    if (idxU) idxU[s * KM + e] = pos;
}

// Upper tail of the F distribution, cephes fdtrc(a, b, x) = incbet(b/2, a/2, b / (b + a x)); NaN for x < 0
// (cephes' domain error), which the caller turns into "no contribution" like the reference's nan_to_num.
__device__ __forceinline__ double p_f_upper(double f, double dfn, double dfd, double ln_beta) {
    if (!(f == f) || f < 0.0) return __builtin_nan("");
    if (isinf(f)) return 0.0;
    const double den = dfd + dfn * f;
    const double x = dfd / den, y = dfn * f / den;
    if (y <= 0.0) return 1.0;
    if (x <= 0.0) return 0.0;
    const double a = 0.5 * dfd, b = 0.5 * dfn;
    const double lnx = (y < 0.5) ? log1p(-y) : log(x);
    const double lny = (x < 0.5) ? log1p(-x) : log(y);
    const double front = exp(a * lnx + b * lny - ln_beta);
    if (x < (a + 1.0) / (a + b + 2.0)) return front * betacf(a, b, x) / a;
    return 1.0 - front * betacf(b, a, y) / b;
}

template <typename T> struct AnovaAcc { typedef double type; };
template <> struct AnovaAcc<float> { typedef float type; };

template <typename T, bool STAGED>
__global__ void __launch_bounds__(kAnovaThreads, 1) k_anova(const AnovaArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *vec = reinterpret_cast<T *>(smem_raw);
    __shared__ double red[kAnovaThreads / 32];
    typedef typename AnovaAcc<T>::type A;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.K, m = a.m;
    const double bign = (double)K * (double)m;
    for (int64_t g = a.g_begin + blockIdx.x; g < a.g_end; g += gridDim.x) {
        const T *row = static_cast<const T *>(a.xg) + (size_t)(g - a.xg_g0) * (size_t)a.n_pad;
        const T *src = row;
        if (STAGED) {
            __syncthreads();   // the previous gene's gathers are done
            for (int64_t i = tid; i < a.N; i += kAnovaThreads) vec[i] = row[i];
            __syncthreads();
            src = vec;
        }
        const T c0 = src[0];
        double fisher = 0.0;
        for (int64_t s = tid; s < a.S; s += kAnovaThreads) {
            const uint32_t *ix = a.idxT + s;
            double ssb_raw = 0.0, tot = 0.0, q = 0.0;   // sum_k s_k^2 / m, sum of all deviations, sum of squares
            bool all_const = true;                        // every group constant (f_oneway: F = inf -> p = 0)
            T first_all = T(0);
            bool all_same = true;                         // every value identical (F = nan -> p = nan -> 0)
            for (int k = 0; k < K; ++k) {
                A sk = A(0), qk = A(0);
                T first = T(0);
                bool cst = true;
#pragma unroll 4
                for (int e = 0; e < m; ++e) {
                    const uint32_t pos = __ldg(ix + (int64_t)(k * m + e) * a.S);
                    const T v = src[pos];
                    if (e == 0) first = v;
                    cst = cst && (v == first);
                    const A d = (A)(v - c0);
                    sk += d;
                    qk = fma(d, d, qk);
                }
                if (k == 0) first_all = first;
                all_const = all_const && cst;
                all_same = all_same && cst && (first == first_all);
                const double sd = (double)sk;
                ssb_raw += sd * sd / (double)m;
                tot += sd;
                q += (double)qk;
            }
            // about the pooled mean (the shift by c0 cancels): sstot = q - tot^2/n, ssbn = sum s_k^2/m - tot^2/n
            const double nss = tot * tot / bign;
            const double sstot = q - nss;
            const double ssbn = ssb_raw - nss;
            const double sswn = sstot - ssbn;
            double f = (ssbn / a.dfn) / (sswn / a.dfd);
            if (all_const) f = CUDART_INF;
            if (all_same) f = __builtin_nan("");
            double p = p_f_upper(f, a.dfn, a.dfd, a.ln_beta);
            if (p <= a.p_flush) p = 0.0;              // float32 fdtrc underflow (float input: scipy returns float32)
            if (!(p == p) || isinf(p)) p = 0.0;       // phet.py:378 nan_to_num
            if (a.dbgP) a.dbgP[(size_t)g * (size_t)a.S + (size_t)s] = p;
            fisher += fisher_term(p);
        }
        // fixed-order CTA reduction
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) fisher += __shfl_xor_sync(0xFFFFFFFFu, fisher, o);
        __syncthreads();
        if (lane == 0) red[warp] = fisher;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < kAnovaThreads / 32; ++w) t += red[w];
            a.acc_f[g] = t;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------
// Kruskal-Wallis H (phet.py:373-375, group_test="kruskal" with m > 5): scipy.stats.kruskal(*subsamples) --
// average ranks of the pooled K*m values, H = 12/(n(n+1)) * sum_k R_k^2/m - 3(n+1), divided by the tie correction
// 1 - sum(t^3 - t)/(n^3 - n), p = chi2.sf(H, K-1).  One warp per (gene, subsample): the pooled values sit in the
// warp's shared-memory strip and every element is ranked by counting (#{x < v}, #{x <= v}: 128-bit broadcast
// loads, four own elements per pass) -- n^2 comparisons per unit but no sort, no payload, and exact integer
// (doubled) ranks, rank sums and tie sums, so H is the reference's up to its last few floating-point operations.
constexpr int kKwThreads = 512;
constexpr int kKwMaxGroups = 64;

// chi2.sf(x, df) = igamc(df/2, x/2), df a small positive integer: finite sums (no cancellation)
__device__ __forceinline__ double chi2_sf_int(double x, int df) {
    if (!(x == x)) return __builtin_nan("");
    if (x < 0.0) return 1.0;   // cephes chdtrc
    const double y = 0.5 * x;
    if (isinf(y)) return 0.0;
    const double ey = exp(-y);
    if ((df & 1) == 0) {       // a = df/2 integer: e^-y sum_{i<a} y^i / i!
        double term = 1.0, sum = 1.0;
        for (int i = 1; i < df / 2; ++i) {
            term *= y / (double)i;
            sum += term;
        }
        return ey * sum;
    }
    // a = j + 1/2: erfc(sqrt y) + e^-y sum_{i=1..j} y^(i-1/2) / Gamma(i+1/2)
    const double sy = sqrt(y);
    double res = erfc(sy);
    double term = sy / 0.88622692545275801365;   // y^(1/2) / Gamma(3/2)
    for (int i = 1; i <= df / 2; ++i) {
        res += ey * term;
        term *= y / ((double)i + 0.5);
    }
    return res;
}

template <typename T> struct KwVec;
template <> struct KwVec<float> {
    typedef float4 type;
    static constexpr int width = 4;
    static __device__ __forceinline__ void get(const float4 &v, float *o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
};
template <> struct KwVec<double> {
    typedef double2 type;
    static constexpr int width = 2;
    static __device__ __forceinline__ void get(const double2 &v, double *o) { o[0] = v.x; o[1] = v.y; }
};

template <typename T>
__global__ void __launch_bounds__(kKwThreads, 1) k_kruskal(const AnovaArgs a, int n_pad4) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = kKwThreads / 32;
    constexpr int W = KwVec<T>::width;
    T *pool_all = reinterpret_cast<T *>(smem_raw);                                  // [NW][n_pad4]
    int *gsum_all = reinterpret_cast<int *>(smem_raw + (size_t)NW * n_pad4 * sizeof(T));   // [NW][K] doubled rank sums
    __shared__ double red[NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.K, m = a.m, n = K * m;
    T *pool = pool_all + (size_t)warp * n_pad4;
    int *gsum = gsum_all + warp * K;
    const double dn = (double)n;
    for (int64_t g = a.g_begin + blockIdx.x; g < a.g_end; g += gridDim.x) {
        const T *row = static_cast<const T *>(a.xg) + (size_t)(g - a.xg_g0) * (size_t)a.n_pad;
        double fisher = 0.0;
        for (int64_t s = warp; s < a.S; s += NW) {
            for (int e = lane; e < n_pad4; e += 32)
                pool[e] = (e < n) ? __ldg(row + __ldg(a.idxT + (int64_t)e * a.S + s)) : (T)CUDART_INF;
            for (int k = lane; k < K; k += 32) gsum[k] = 0;
            __syncwarp();
            long long tie = 0;   // sum over elements of (t^2 - 1), t = size of the element's tie group
            for (int base = 0; base < n; base += 128) {
                T v[4];
                int lo[4], hi[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = base + lane + 32 * q;
                    v[q] = pool[e < n ? e : 0];
                    lo[q] = 0;
                    hi[q] = 0;
                }
                const typename KwVec<T>::type *pv = reinterpret_cast<const typename KwVec<T>::type *>(pool);
                for (int j = 0; j < n_pad4 / W; ++j) {
                    T x[W];
                    KwVec<T>::get(pv[j], x);
#pragma unroll
                    for (int w = 0; w < W; ++w) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            lo[q] += (x[w] < v[q]) ? 1 : 0;
                            hi[q] += (x[w] <= v[q]) ? 1 : 0;
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = base + lane + 32 * q;
                    if (e < n) {
                        atomicAdd(&gsum[e / m], lo[q] + hi[q] + 1);   // twice the average rank
                        const long long t = hi[q] - lo[q];
                        tie += t * t - 1;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) tie += __shfl_xor_sync(0xFFFFFFFFu, tie, o);
            __syncwarp();
            if (lane == 0) {
                double ssbn = 0.0;
                for (int k = 0; k < K; ++k) {
                    const double r = 0.5 * (double)gsum[k];
                    ssbn += r * r / (double)m;
                }
                double h = 12.0 / (dn * (dn + 1.0)) * ssbn - 3.0 * (dn + 1.0);
                const double ties = 1.0 - (double)tie / (dn * dn * dn - dn);
                h /= ties;
                double p = chi2_sf_int(h, K - 1);
                if (p <= a.p_flush) p = 0.0;
                if (!(p == p) || isinf(p)) p = 0.0;       // phet.py:378 nan_to_num
                if (a.dbgP) a.dbgP[(size_t)g * (size_t)a.S + (size_t)s] = p;
                fisher += fisher_term(p);
            }
            __syncwarp();
        }
        __syncthreads();
        if (lane == 0) red[warp] = fisher;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < NW; ++w) t += red[w];
            a.acc_f[g] = t;
        }
    }
}

// Kruskal-Wallis by SORTING (n log^2 n instead of n^2): one warp per (gene, subsample) sorts the pooled K * m values with
// the register network of warp_prims.cuh, parks the sorted strip in shared memory, and every element finds
// #{x < v} and #{x <= v} by two binary searches -- the same two integers the counting kernel forms, so the doubled
// ranks, the per-group rank sums and the tie sum are the same exact integers (phet.py:372-377 -> scipy.stats.kruskal).
// E = registers per lane (32 E >= K m).
template <typename T, int E, bool STAGED>
__global__ void __launch_bounds__(kKwThreads, (sizeof(T) == 4 && E <= 8) ? 2 : 1) k_kruskal_sort(const AnovaArgs a, int vec_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = kKwThreads / 32;
    constexpr int NP = 32 * E;
    // the gene's cell vector is staged once per gene (the K m gathers of the S subsamples then hit shared memory, not L2)
    T *vec = reinterpret_cast<T *>(smem_raw);                                                      // [N] when STAGED
    T *pool_all = reinterpret_cast<T *>(smem_raw + vec_bytes);                                     // [NW][NP] sorted strips
    int *gsum_all = reinterpret_cast<int *>(smem_raw + vec_bytes + (size_t)NW * NP * sizeof(T));   // [NW][K] doubled rank sums
    __shared__ double red[NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.K, m = a.m, n = K * m;
    T *pool = pool_all + (size_t)warp * NP;
    int *gsum = gsum_all + warp * K;
    const double dn = (double)n;
    for (int64_t g = a.g_begin + blockIdx.x; g < a.g_end; g += gridDim.x) {
        const T *row = static_cast<const T *>(a.xg) + (size_t)(g - a.xg_g0) * (size_t)a.n_pad;
        if (STAGED) {
            __syncthreads();   // the previous gene's gathers are done
            for (int64_t i = tid; i < a.N; i += kKwThreads) vec[i] = row[i];
            __syncthreads();
        }
        const T *src = STAGED ? vec : row;
        double fisher = 0.0, my_ssbn = 0.0, my_tie = 0.0;
        int64_t my_s = -1;
        for (int64_t s = warp; s < a.S; s += NW) {
            const uint32_t *ix = a.idxU + s * (int64_t)n;   // this subsample's positions, adjacent
            T v[E], orig[E];
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int e = lane + 32 * r;
                v[r] = (e < n) ? src[__ldg(ix + e)] : Lim<T>::inf();
                orig[r] = v[r];
            }
            for (int k = lane; k < K; k += 32) gsum[k] = 0;
            warp_sort_asc<T, E>(v, lane);
#pragma unroll
            for (int r = 0; r < E; ++r) pool[lane + 32 * r] = v[r];
            __syncwarp();
            long long tie = 0;   // sum over elements of (t^2 - 1), t = size of the element's tie group
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int e = lane + 32 * r;
                const bool ok = e < n;
                const T x = orig[r];
                // lower / upper bound of x in pool[0, n): branch-free bisection over the padded power-of-two span
                int lo = 0, hi = 0;
#pragma unroll
                for (int step = next_pow2(NP); step >= 1; step >>= 1) {   // (the steps must be able to add up to n = NP)
                    const int pl = lo + step, ph = hi + step;
                    if (pl <= n && pool[pl - 1] < x) lo = pl;
                    if (ph <= n && pool[ph - 1] <= x) hi = ph;
                }
                const int dr = ok ? lo + hi + 1 : 0;   // twice the average rank
                if (ok) {
                    const long long t = hi - lo;
                    tie += t * t - 1;
                }
                if (m >= 32) {   // the 32 elements of this row belong to at most two groups: two warp reductions
                    const int gA = (32 * r) / m;
                    const int bnd = (gA + 1) * m - 32 * r;   // first lane of group gA + 1
                    const int sA = __reduce_add_sync(0xFFFFFFFFu, lane < bnd ? dr : 0);
                    const int sB = __reduce_add_sync(0xFFFFFFFFu, lane < bnd ? 0 : dr);
                    if (lane == 0) {
                        if (gA < K) gsum[gA] += sA;
                        if (gA + 1 < K) gsum[gA + 1] += sB;
                    }
                } else if (ok) {
                    atomicAdd(&gsum[e / m], dr);
                }
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) tie += __shfl_xor_sync(0xFFFFFFFFu, tie, o);
            __syncwarp();
            // the statistic of this unit is parked; every 32 units the warp evaluates 32 tail probabilities at once
            // (one lane each) instead of one lane evaluating exp / erfc / log while 31 wait
            const int slot = (int)(((s - warp) / NW) & 31);
            if (lane == slot) {
                double ssbn = 0.0;
                for (int k = 0; k < K; ++k) {   // (same order as the counting kernel: the two stay bit-identical)
                    const double r = 0.5 * (double)gsum[k];
                    ssbn += r * r / (double)m;
                }
                my_ssbn = ssbn;
                my_tie = (double)tie;
                my_s = s;
            }
            __syncwarp();
            if (slot == 31 || s + NW >= a.S) {
                if (my_s >= 0) {
                    double h = 12.0 / (dn * (dn + 1.0)) * my_ssbn - 3.0 * (dn + 1.0);
                    const double ties = 1.0 - my_tie / (dn * dn * dn - dn);
                    h /= ties;
                    double p = chi2_sf_int(h, K - 1);
                    if (p <= a.p_flush) p = 0.0;
                    if (!(p == p) || isinf(p)) p = 0.0;       // phet.py:378 nan_to_num
                    if (a.dbgP) a.dbgP[(size_t)g * (size_t)a.S + (size_t)my_s] = p;
                    fisher += fisher_term(p);
                    my_s = -1;
                }
            }
        }
        fisher = warp_sum(fisher);
        __syncthreads();
        if (lane == 0) red[warp] = fisher;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < NW; ++w) t += red[w];
            a.acc_f[g] = t;
        }
    }
}

template <typename T, int E>
static int launch_kruskal_sort_E(phet_ctx *c, const AnovaArgs &a, int64_t grid) {
    const size_t strips = (size_t)(kKwThreads / 32) * ((size_t)32 * E * sizeof(T) + (size_t)a.K * sizeof(int));
    if (strips > (size_t)kMaxSmemOptin - 1024) return 1;
    const size_t vec_bytes = ((size_t)a.N * sizeof(T) + 15) & ~(size_t)15;
    if (vec_bytes + strips <= (size_t)kMaxSmemOptin - 1024) {
        PHET_CUDA(cudaFuncSetAttribute(k_kruskal_sort<T, E, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(vec_bytes + strips)));
        int per_sm = 1;   // narrow float instantiations are compiled for two CTAs per SM (64 registers)
        PHET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_kruskal_sort<T, E, true>, kKwThreads, vec_bytes + strips));
        const int64_t ng = a.g_end - a.g_begin;
        if (per_sm > 1) grid = (int64_t)per_sm * c->num_sms < ng ? (int64_t)per_sm * c->num_sms : ng;
        k_kruskal_sort<T, E, true><<<(unsigned)grid, kKwThreads, vec_bytes + strips, c->stream>>>(a, (int)vec_bytes);
    } else {   // a cell vector too long to stage: gathers go to L2
        PHET_CUDA(cudaFuncSetAttribute(k_kruskal_sort<T, E, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)strips));
        k_kruskal_sort<T, E, false><<<(unsigned)grid, kKwThreads, strips, c->stream>>>(a, 0);
    }
    return 0;
}

// Returns 1 when the pooled sample does not fit the sort kernel (more than 1024 values): the counting kernel takes it.
template <typename T>
static int launch_kruskal_sort(phet_ctx *c, const AnovaArgs &a, int64_t grid) {
    const int need = (a.K * a.m + 31) / 32;
    if (sizeof(T) == 8 && need > 12) return 1;   // (two copies of the values per lane: wider double instantiations spill)
    if (need <= 1) return launch_kruskal_sort_E<T, 1>(c, a, grid);
    if (need <= 2) return launch_kruskal_sort_E<T, 2>(c, a, grid);
    if (need <= 3) return launch_kruskal_sort_E<T, 3>(c, a, grid);
    if (need <= 4) return launch_kruskal_sort_E<T, 4>(c, a, grid);
    if (need <= 6) return launch_kruskal_sort_E<T, 6>(c, a, grid);
    if (need <= 8) return launch_kruskal_sort_E<T, 8>(c, a, grid);
    if (need <= 12) return launch_kruskal_sort_E<T, 12>(c, a, grid);
    if (need <= 16) return launch_kruskal_sort_E<T, 16>(c, a, grid);
    if (need <= 24) return launch_kruskal_sort_E<T, 24>(c, a, grid);
    if (need <= 32) return launch_kruskal_sort_E<T, 32>(c, a, grid);
    return 1;
}

int launch_anova(phet_ctx *c, const int32_t *idx_host, int64_t S, int64_t K, int64_t m, double ln_beta, double p_flush,
                 int capture_debug, int kruskal) {
    const int64_t KM = K * m, N = c->n0 + c->n1;
    if (int e = c->anova_idx.alloc((size_t)(S * KM))) return e;
    if (int e = c->anova_idxT.alloc((size_t)(S * KM))) return e;
    PHET_CUDA(cudaMemcpyAsync(c->anova_idx.p, idx_host, sizeof(int32_t) * (size_t)(S * KM), cudaMemcpyHostToDevice, c->stream));
    int *bad = reinterpret_cast<int *>(c->scalars.p + 12);
    PHET_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), c->stream));
    uint32_t *idxU = nullptr;
    if (kruskal) {   // warp-per-subsample kernel: one subsample's positions adjacent
        if (int e = c->anova_idxU.alloc((size_t)(S * KM))) return e;
        idxU = c->anova_idxU.p;
    }
    k_anova_index<<<(unsigned)((S * KM + 255) / 256), 256, 0, c->stream>>>(c->anova_idx.p, S, KM, c->pos_of_row.p, N, c->anova_idxT.p,
                                                                          idxU, bad);
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    int hbad = 0;
    PHET_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));   // also: idx_host is the caller's buffer
    PHET_REQUIRE(hbad == 0, "index sets contain a row number outside [0, N)");

    AnovaArgs a;
    a.xg = c->xg.p;
    a.n_pad = c->n_pad;
    a.xg_g0 = c->xg_g0;
    a.g_begin = c->xg_g0;
    a.g_end = c->xg_g0 + c->xg_rows;
    a.N = N;
    a.idxT = c->anova_idxT.p;
    a.idxU = idxU;
    a.S = S;
    a.K = (int)K;
    a.m = (int)m;
    a.dfn = (double)(K - 1);
    a.dfd = (double)(KM - K);
    a.ln_beta = ln_beta;
    a.p_flush = p_flush;
    a.acc_f = c->acc.p;
    a.dbgP = nullptr;
    if (capture_debug) {
        const size_t need = (size_t)c->G * (size_t)S;
        if (int e = c->dbgP.alloc(need)) return e;
        PHET_CUDA(cudaMemsetAsync(c->dbgP.p, 0, need * sizeof(double), c->stream));
        a.dbgP = c->dbgP.p;
        c->S_local = S;
    }
    const int64_t ng = a.g_end - a.g_begin;
    if (ng < 1) return PHET_OK;
    const size_t e_st = c->storage == PHET_F32 ? 4 : 8;
    if (kruskal) {
        PHET_REQUIRE(K <= kKwMaxGroups, "phet_group_test: Kruskal-Wallis takes at most 64 groups");
        const int n_pad4 = (int)((KM + 3) / 4 * 4);
        const size_t smem_kw = (size_t)(kKwThreads / 32) * ((size_t)n_pad4 * e_st + (size_t)K * sizeof(int));
        PHET_REQUIRE(smem_kw <= (size_t)kMaxSmemOptin - 1024, "phet_group_test: K * m too large for the per-warp rank strips");
        int64_t gridk = c->num_sms < ng ? c->num_sms : ng;
        if (getenv("PHET_KRUSKAL_IMPL") == nullptr || strcmp(getenv("PHET_KRUSKAL_IMPL"), "count") != 0) {
            const int rc = c->storage == PHET_F32 ? launch_kruskal_sort<float>(c, a, gridk) : launch_kruskal_sort<double>(c, a, gridk);
            if (rc != 1) {
                if (rc) return rc;
                c->launches++;
                PHET_CUDA(cudaGetLastError());
                return PHET_OK;
            }
        }
        if (c->storage == PHET_F32) {
            PHET_CUDA(cudaFuncSetAttribute(k_kruskal<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kw));
            k_kruskal<float><<<(unsigned)gridk, kKwThreads, smem_kw, c->stream>>>(a, n_pad4);
        } else {
            PHET_CUDA(cudaFuncSetAttribute(k_kruskal<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kw));
            k_kruskal<double><<<(unsigned)gridk, kKwThreads, smem_kw, c->stream>>>(a, n_pad4);
        }
        c->launches++;
        PHET_CUDA(cudaGetLastError());
        return PHET_OK;
    }
    const size_t smem = (size_t)c->n_pad * e_st;
    const bool staged = smem <= (size_t)kMaxSmemOptin - 1024;
    int64_t grid = c->num_sms < ng ? c->num_sms : ng;
    if (c->storage == PHET_F32) {
        if (staged) {
            PHET_CUDA(cudaFuncSetAttribute(k_anova<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_anova<float, true><<<(unsigned)grid, kAnovaThreads, smem, c->stream>>>(a);
        } else {
            k_anova<float, false><<<(unsigned)grid, kAnovaThreads, 0, c->stream>>>(a);
        }
    } else {
        if (staged) {
            PHET_CUDA(cudaFuncSetAttribute(k_anova<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_anova<double, true><<<(unsigned)grid, kAnovaThreads, smem, c->stream>>>(a);
        } else {
            k_anova<double, false><<<(unsigned)grid, kAnovaThreads, 0, c->stream>>>(a);
        }
    }
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    return PHET_OK;
}

}  // namespace phet
