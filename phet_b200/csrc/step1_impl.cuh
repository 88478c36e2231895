// Step 1: the recurrent-subsampling loop of PHet.fit_predict, one (gene, subsample) unit per
// warp-iteration.  Replaces /root/reference/src/model/phet.py:397-432 (+ the per-gene sums of
// 435 and 449-451): for every subsample s and gene g
//     delta[g,s] = | IQR(X[idx_ctrl[s], g]) - IQR(X[idx_case[s], g]) |      (phet.py:413-417)
//     p[g,s]     = two-sided p of the pooled t (m < 30) or z statistic      (phet.py:425-429)
// and accumulates  acc_f[g] = sum_s -2 ln p[g,s] (p == 0 / NaN dropped),  acc_r[g] = sum_s delta[g,s].
//
// Design (B200): persistent CTAs; each CTA stages ONE gene's full cell vector (class-grouped,
// n_pad elements) into shared memory with 1-D TMA bulk copies signalled on an mbarrier, then
// its warps serve every subsample of the launch from SMEM: coalesced index loads (L2-resident
// index sets), random SMEM gathers, fp64 warp-shuffle reductions for the moments, and a
// register-resident pruned bitonic network (warp_sort_asc) for the order statistics.  p-values
// are evaluated 32 at a time (one per lane) so the fp64 erfc/log cost is amortised.
// When a gene does not fit in SMEM the same code gathers straight from global/L2 (STAGED=false).
//
// Algorithmic bytes per unit: 2*m*sizeof(T) SMEM gather + 2*m*4 index bytes (L2) and
// n_pad*sizeof(T)/S_launch amortised HBM bytes.
#pragma once

#include "common.cuh"
#include "pvalue.cuh"
#include "warp_prims.cuh"

namespace phet {

constexpr int kThreads1 = 512;

struct Step1Args {
    const void *xg;      // row of gene g starts at xg + (g - xg_g0) * n_pad elements
    int64_t n_pad, G, xg_g0;
    int64_t g_begin, g_end;
    const int32_t *idx;  // [S][2][m] grouped positions
    int64_t m;
    int64_t s_begin, s_end;
    double *acc_f, *acc_r;
    double *dbgP, *dbgD;  // [G][dbg_ld] or nullptr
    int64_t dbg_ld;
    phet_quantile_pos q;
    int test;
    double df, ln_beta, p_flush;
    int accumulate;
};

template <typename T, int E>
__device__ __forceinline__ double iqr_of_sorted(const T (&v)[E], const phet_quantile_pos &q) {
    const double y_jl = (double)warp_pick<T, E>(v, q.j_lo);
    const double y_kl = (double)warp_pick<T, E>(v, q.k_lo);
    const double y_jh = (double)warp_pick<T, E>(v, q.j_hi);
    const double y_kh = (double)warp_pick<T, E>(v, q.k_hi);
    // (1-g)*y[j] + g*y[j+1], products and sum rounded separately like numpy (scipy _quantile_hf)
    const double lo = __dadd_rn(__dmul_rn(1.0 - q.g_lo, y_jl), __dmul_rn(q.g_lo, y_kl));
    const double hi = __dadd_rn(__dmul_rn(1.0 - q.g_hi, y_jh), __dmul_rn(q.g_hi, y_kh));
    return hi - lo;
}

template <typename T, int E, bool STAGED>
__global__ void __launch_bounds__(kThreads1, 1) k_step1(const Step1Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    const int m = (int)a.m;

    // shared memory carve-up: [gene row | mbarrier | per-warp partials]
    const size_t gene_bytes = STAGED ? (size_t)a.n_pad * sizeof(T) : 0;
    T *gene_s = reinterpret_cast<T *>(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + ((gene_bytes + 15) & ~(size_t)15));
    double *part = reinterpret_cast<double *>(bar + 2);

    if (STAGED) {
        if (tid == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    unsigned phase = 0;
    const double n = (double)m;
    const double inv_pair = 1.0 / n + 1.0 / n;

    for (int64_t g = a.g_begin + blockIdx.x; g < a.g_end; g += gridDim.x) {
        const T *gene;
        if (STAGED) {
            if (tid == 0) stage_row(gene_s, static_cast<const T *>(a.xg) + (g - a.xg_g0) * a.n_pad, gene_bytes, bar);
            mbar_wait(bar, phase);
            phase ^= 1u;
            gene = gene_s;
        } else {
            gene = static_cast<const T *>(a.xg) + (g - a.xg_g0) * a.n_pad;
        }

        double facc = 0.0;  // per lane: sum of Fisher terms this lane evaluated
        double racc = 0.0;  // warp-uniform: sum of delta
        double stat_slot = 0.0;
        int cnt = 0;

        int ia[E], ib[E];
        int64_t s = a.s_begin + warp;
        if (s < a.s_end) {
            const int32_t *pa = a.idx + (s * 2) * a.m;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int k = lane + 32 * r;
                ia[r] = (k < m) ? __ldg(pa + k) : -1;
                ib[r] = (k < m) ? __ldg(pa + m + k) : -1;
            }
        }
        for (; s < a.s_end; s += nw) {
            T va[E], vb[E];
            double sa = 0.0, qa = 0.0, sb = 0.0, qb = 0.0;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                if (ia[r] >= 0) {
                    const T x = STAGED ? gene[ia[r]] : __ldg(gene + ia[r]);
                    const T y = STAGED ? gene[ib[r]] : __ldg(gene + ib[r]);
                    va[r] = x;
                    vb[r] = y;
                    const double xd = (double)x, yd = (double)y;
                    sa += xd;
                    qa = fma(xd, xd, qa);
                    sb += yd;
                    qb = fma(yd, yd, qb);
                } else {
                    va[r] = Lim<T>::inf();
                    vb[r] = Lim<T>::inf();
                }
            }
            // prefetch the next subsample's index sets while this one is processed
            {
                const int64_t sn = s + nw;
                if (sn < a.s_end) {
                    const int32_t *pa = a.idx + (sn * 2) * a.m;
#pragma unroll
                    for (int r = 0; r < E; ++r) {
                        const int k = lane + 32 * r;
                        ia[r] = (k < m) ? __ldg(pa + k) : -1;
                        ib[r] = (k < m) ? __ldg(pa + m + k) : -1;
                    }
                }
            }
            sa = warp_sum(sa);
            qa = warp_sum(qa);
            sb = warp_sum(sb);
            qb = warp_sum(qb);
            // pooled statistic (scipy _equal_var_ttest_denom / statsmodels ztest 'pooled'):
            // svar = (SS_a + SS_b)/(n1+n2-2), denom = sqrt(svar*(1/n1+1/n2)), stat = (mean_a-mean_b)/denom
            const double ssa = qa - sa * sa / n;
            const double ssb = qb - sb * sb / n;
            const double svar = (ssa + ssb) / a.df;
            const double denom = sqrt(svar * inv_pair);
            const double stat = (sa / n - sb / n) / denom;

            warp_sort_asc<T, E>(va, lane);
            warp_sort_asc<T, E>(vb, lane);
            const double iq_a = iqr_of_sorted<T, E>(va, a.q);
            const double iq_b = iqr_of_sorted<T, E>(vb, a.q);
            double delta = fabs(iq_a - iq_b);
            if (!isfinite(delta)) delta = 0.0;  // phet.py:416
            racc += delta;
            if (a.dbgD != nullptr && lane == 0) a.dbgD[g * a.dbg_ld + (s - a.s_begin)] = delta;

            if (lane == (cnt & 31)) stat_slot = stat;
            ++cnt;
            if ((cnt & 31) == 0) {
                double p = (a.test == PHET_TEST_Z) ? p_two_sided_z(stat_slot)
                                                   : p_two_sided_t(stat_slot, a.df, a.ln_beta);
                if (p <= a.p_flush) p = 0.0;
                facc += fisher_term(p);
                if (a.dbgP != nullptr) {
                    const int64_t sl = a.s_begin + warp + (int64_t)(cnt - 32 + lane) * nw;
                    a.dbgP[g * a.dbg_ld + (sl - a.s_begin)] = (p == p && !isinf(p)) ? p : 0.0;
                }
            }
        }
        {
            const int rem = cnt & 31;
            if (lane < rem) {
                double p = (a.test == PHET_TEST_Z) ? p_two_sided_z(stat_slot)
                                                   : p_two_sided_t(stat_slot, a.df, a.ln_beta);
                if (p <= a.p_flush) p = 0.0;
                facc += fisher_term(p);
                if (a.dbgP != nullptr) {
                    const int64_t sl = a.s_begin + warp + (int64_t)(cnt - rem + lane) * nw;
                    a.dbgP[g * a.dbg_ld + (sl - a.s_begin)] = (p == p && !isinf(p)) ? p : 0.0;
                }
            }
        }
        facc = warp_sum(facc);
        if (lane == 0) {
            part[warp * 2 + 0] = facc;
            part[warp * 2 + 1] = racc;
        }
        __syncthreads();
        if (tid == 0) {
            double tf = 0.0, tr = 0.0;
            for (int w = 0; w < nw; ++w) {  // fixed order: reproducible for a given launch shape
                tf += part[w * 2 + 0];
                tr += part[w * 2 + 1];
            }
            if (a.accumulate) {
                tf += a.acc_f[g];
                tr += a.acc_r[g];
            }
            a.acc_f[g] = tf;
            a.acc_r[g] = tr;
        }
        __syncthreads();  // all SMEM reads of this gene done before the next bulk copy lands
    }
}

template <typename T, int E, bool STAGED>
static int launch_step1_inst(phet_ctx *c, const Step1Args &args, size_t smem) {
    auto kern = k_step1<T, E, STAGED>;
    PHET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PHET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads1, smem));
    if (per_sm < 1) {
        set_error("step1: kernel does not fit on an SM");
        return PHET_ERR_UNSUPPORTED;
    }
    int64_t grid = (int64_t)per_sm * c->num_sms;
    if (grid > args.g_end - args.g_begin) grid = args.g_end - args.g_begin;
    if (grid < 1) return 0;
    c->step1_geom[0] = (int32_t)grid;
    c->step1_geom[1] = kThreads1;
    c->step1_geom[2] = (int32_t)smem;
    c->step1_geom[3] = E;
    c->step1_geom[4] = STAGED ? 1 : 0;
    kern<<<(unsigned)grid, kThreads1, smem, c->stream>>>(args);
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, bool STAGED>
static int dispatch_step1_E(phet_ctx *c, const Step1Args &args, size_t smem) {
    const int need = (int)((args.m + 31) / 32);
    // m <= 256 is served by the half-warp path (step1h_impl.cuh); this network covers 257..512
    if (need <= 10) return launch_step1_inst<T, 10, STAGED>(c, args, smem);
    if (need <= 12) return launch_step1_inst<T, 12, STAGED>(c, args, smem);
    if (need <= 16) return launch_step1_inst<T, 16, STAGED>(c, args, smem);
    set_error("step1: subsample size m > 512 is not supported by the register-resident sort");
    return PHET_ERR_UNSUPPORTED;
}

template <typename T>
static int run_step1_T(phet_ctx *c, const Step1Args &args) {
    const size_t tail = 16 + 16 + (size_t)(kThreads1 / 32) * 2 * sizeof(double);
    const size_t staged = (((size_t)c->n_pad * sizeof(T) + 15) & ~(size_t)15) + tail;
    if (staged <= (size_t)kMaxSmemOptin) return dispatch_step1_E<T, true>(c, args, staged);
    return dispatch_step1_E<T, false>(c, args, tail);
}

}  // namespace phet
