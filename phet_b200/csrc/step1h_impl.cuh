// Step 1, fast path (m <= 256): same contract as k_step1 (step1_impl.cuh) but the two groups of
// a (gene, subsample) unit are handled by the two halves of one warp: lanes 0-15 gather, reduce
// and sort the control draw, lanes 16-31 the case draw, with the half-warp merge-split sorter
// (warp_prims.cuh).  E = ceil(m/16) values per lane.  Replaces phet.py:397-432 + 435, 449-451.
#pragma once

#include "step1_impl.cuh"
#include "step1p_impl.cuh"
#include <cstdlib>

namespace phet {

// Registers r < EV of a lane are valid for every m this instantiation serves (dispatch picks the
// smallest E with 16*E >= m; E = 12 / 16 serve need 11..16), so only the tail needs bound checks.
template <int E>
struct StaticValid {
    static constexpr int value = (E <= 10) ? (E - 1) : 10;
};

template <typename T, int E, bool STAGED, int THREADS = kThreads1>
__global__ void __launch_bounds__(THREADS, 1) k_step1h(const Step1Args a) {
    constexpr int EV = StaticValid<E>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    const int grp = lane >> 4;  // 0: control draw, 1: case draw
    const int h = lane & 15;
    const int m = (int)a.m;

    const size_t gene_bytes = STAGED ? (size_t)a.n_pad * sizeof(T) : 0;
    T *gene_s = reinterpret_cast<T *>(smem_raw);
    const uint32_t gene_sa = smem_u32(gene_s);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + ((gene_bytes + 15) & ~(size_t)15));
    double *part = reinterpret_cast<double *>(bar + 2);
    // per-warp scratch for the sorted blocks: the order statistics are read back from SMEM
    // (4 broadcast LDS) instead of a register-select tree; the second half is skewed by one
    // word so the two half-warps do not hit the same banks
    T *sorted_s = reinterpret_cast<T *>(part + 2 * (THREADS / 32)) + (size_t)warp * (32 * E + 2) + grp * (16 * E + 1);

    if (STAGED) {
        if (tid == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    unsigned phase = 0;
    const double inv_n = 1.0 / (double)m;
    const double inv_df = 1.0 / a.df;  // df == 0 (m == 1) -> inf -> NaN statistic, like the reference
    const double inv_pair = inv_n + inv_n;
    const HalfSortCoef<T> coef(h);
    const double w_lo0 = 1.0 - a.q.g_lo, w_hi0 = 1.0 - a.q.g_hi;

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

        double facc = 0.0, racc = 0.0, stat_slot = 0.0;
        int cnt = 0;
        int id[E];
        int64_t s = a.s_begin + warp;
        if (s < a.s_end) {
            const int32_t *pa = a.idx + (s * 2 + grp) * a.m + h;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                if (r < EV) id[r] = __ldg(pa + 16 * r);
                else id[r] = (h + 16 * r < m) ? __ldg(pa + 16 * r) : -1;
            }
        }
        for (; s < a.s_end; s += nw) {
            T x[E];
            double sx = 0.0, qx = 0.0;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                if (r < EV || id[r] >= 0) {
                    const T v = STAGED ? lds_at(gene_sa, id[r], T(0)) : __ldg(gene + id[r]);
                    x[r] = v;
                    const double vd = (double)v;
                    sx += vd;
                    qx = fma(vd, vd, qx);
                } else {
                    x[r] = Big<T>::v();
                }
            }
            {   // prefetch the next subsample's index set of this half-warp's group
                const int64_t sn = s + nw;
                if (sn < a.s_end) {
                    const int32_t *pa = a.idx + (sn * 2 + grp) * a.m + h;
#pragma unroll
                    for (int r = 0; r < E; ++r) {
                        if (r < EV) id[r] = __ldg(pa + 16 * r);
                        else id[r] = (h + 16 * r < m) ? __ldg(pa + 16 * r) : -1;
                    }
                }
            }
            // moments of my group over the half-warp, then the other group's from lane ^ 16
            sx = half_sum(sx);
            qx = half_sum(qx);
            double ss = qx - sx * sx * inv_n;                 // sum of squared deviations
            if (fabs(ss) <= 1e-15 * qx) ss = 0.0;             // identical values: exactly 0 like the two-pass variance
            const double ss_o = __shfl_xor_sync(kFull, ss, 16);
            const double sx_o = __shfl_xor_sync(kFull, sx, 16);
            // pooled statistic: (mean_a - mean_b) / sqrt((SS_a + SS_b)/df * (1/n + 1/n))
            const double svar = (ss + ss_o) * inv_df;
            const double dmean = (grp == 0 ? (sx - sx_o) : (sx_o - sx)) * inv_n;
            const double stat = dmean * rsqrt(svar * inv_pair);   // 0 * inf = NaN and x * inf = +-inf like x / 0

            halfwarp_sort_asc<T, E>(x, coef);
#pragma unroll
            for (int r = 0; r < E; ++r) sorted_s[h * E + r] = x[r];  // blocked layout == ascending array
            __syncwarp();
            const double y_jl = (double)sorted_s[a.q.j_lo];
            const double y_kl = (double)sorted_s[a.q.k_lo];
            const double y_jh = (double)sorted_s[a.q.j_hi];
            const double y_kh = (double)sorted_s[a.q.k_hi];
            __syncwarp();
            const double q_lo = __dadd_rn(__dmul_rn(w_lo0, y_jl), __dmul_rn(a.q.g_lo, y_kl));
            const double q_hi = __dadd_rn(__dmul_rn(w_hi0, y_jh), __dmul_rn(a.q.g_hi, y_kh));
            const double iq = q_hi - q_lo;
            const double iq_o = __shfl_xor_sync(kFull, iq, 16);
            double delta = fabs(iq - iq_o);
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
            for (int w = 0; w < nw; ++w) {
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
        __syncthreads();
    }
}

template <typename T, int E, bool STAGED, int THREADS = kThreads1>
static int launch_step1h_inst(phet_ctx *c, const Step1Args &args, size_t smem_in) {
    auto kern = k_step1h<T, E, STAGED, THREADS>;
    const size_t smem = smem_in + (size_t)(THREADS - kThreads1) / 32 * 2 * sizeof(double) +
                        (size_t)(THREADS / 32) * (32 * E + 2) * sizeof(T);
    PHET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PHET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) {
        set_error("step1: kernel does not fit on an SM");
        return PHET_ERR_UNSUPPORTED;
    }
    int64_t grid = (int64_t)per_sm * c->num_sms;
    if (grid > args.g_end - args.g_begin) grid = args.g_end - args.g_begin;
    if (grid < 1) return 0;
    c->step1_geom[0] = (int32_t)grid;
    c->step1_geom[1] = THREADS;
    c->step1_geom[2] = (int32_t)smem;
    c->step1_geom[3] = E;
    c->step1_geom[4] = (STAGED ? 1 : 0) | 2;  // bit 1: half-warp merge-split path
    kern<<<(unsigned)grid, THREADS, smem, c->stream>>>(args);
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, bool STAGED>
static int dispatch_step1h_E(phet_ctx *c, const Step1Args &args, size_t smem) {
    const int need = (int)((args.m + 15) / 16);
    if (need <= 1) return launch_step1h_inst<T, 1, STAGED>(c, args, smem);
    if (need <= 2) return launch_step1h_inst<T, 2, STAGED>(c, args, smem);
    if (need <= 3) return launch_step1h_inst<T, 3, STAGED>(c, args, smem);
    if (need <= 4) return launch_step1h_inst<T, 4, STAGED>(c, args, smem);
    if (need <= 5) return launch_step1h_inst<T, 5, STAGED>(c, args, smem);
    if (need <= 6) return launch_step1h_inst<T, 6, STAGED>(c, args, smem);
    if (need <= 7) return launch_step1h_inst<T, 7, STAGED>(c, args, smem);
    if (need <= 8) return launch_step1h_inst<T, 8, STAGED>(c, args, smem);
    if (need <= 9) return launch_step1h_inst<T, 9, STAGED>(c, args, smem);
    if (need <= 10) {
#ifdef PHET_EXPERIMENT_THREADS
        if (sizeof(T) == 4 && STAGED) {
            const char *e = getenv("PHET_STEP1_THREADS");
            const int t = e ? atoi(e) : 512;
            if (t == 640) return launch_step1h_inst<T, 10, STAGED, 640>(c, args, smem);
            if (t == 768) return launch_step1h_inst<T, 10, STAGED, 768>(c, args, smem);
            if (t == 1024) return launch_step1h_inst<T, 10, STAGED, 1024>(c, args, smem);
        }
#endif
        return launch_step1h_inst<T, 10, STAGED>(c, args, smem);
    }
    if (need <= 12) return launch_step1h_inst<T, 12, STAGED>(c, args, smem);
    return launch_step1h_inst<T, 16, STAGED>(c, args, smem);
}

// Pair-per-thread bucket selection when it applies (step1p_impl.cuh); else m <= 256: half-warp
// sorter; larger m: full-warp network (step1_impl.cuh, E = 10, 12, 16 per lane).
template <typename T>
static int run_step1_any(phet_ctx *c, const Step1Args &args) {
    {
        const int rc = run_step1_pair<T>(c, args);
        if (rc != 1) return rc;
    }
    const size_t tail = 16 + 16 + (size_t)(kThreads1 / 32) * 2 * sizeof(double);
    const size_t staged = (((size_t)c->n_pad * sizeof(T) + 15) & ~(size_t)15) + tail;
    const int need = (int)((args.m + 15) / 16);
    const int e_pick = need <= 10 ? need : (need <= 12 ? 12 : 16);
    const size_t scratch = (size_t)(kThreads1 / 32) * (32 * e_pick + 2) * sizeof(T);
    const bool fits = staged + scratch <= (size_t)kMaxSmemOptin;
    if (args.m <= 256) {
        if (fits) return dispatch_step1h_E<T, true>(c, args, staged);
        return dispatch_step1h_E<T, false>(c, args, tail);
    }
    return run_step1_T<T>(c, args);
}

}  // namespace phet
