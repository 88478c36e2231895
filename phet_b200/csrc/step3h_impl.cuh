// Step 3, fast path (slice length m <= 256): same contract as k_step3 (step3_impl.cuh); the control
// slice is sorted by lanes 0-15 and the case slice by lanes 16-31 with the half-warp merge-split
// sorter, both land in a per-warp SMEM scratch (blocked layout = plain ascending arrays), then all
// 32 lanes evaluate the ECDF difference at the a-values by binary search in b.
// Replaces /root/reference/src/model/phet.py:473-498.
#pragma once

#include "step3_impl.cuh"

namespace phet {

template <typename T> int run_step3_scatter(phet_ctx *c, const Step3Args &args);   // step3s_impl.cuh

template <typename T, int E, bool STAGED>
__global__ void __launch_bounds__(kThreads3, 1) k_step3h(const Step3Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    const int grp = lane >> 4;
    const int h = lane & 15;

    const size_t gene_bytes = STAGED ? (size_t)a.n_pad * sizeof(T) : 0;
    T *gene_s = reinterpret_cast<T *>(smem_raw);
    const uint32_t gene_sa = smem_u32(gene_s);
    unsigned char *after = smem_raw + ((gene_bytes + 15) & ~(size_t)15);
    uint64_t *bar = reinterpret_cast<uint64_t *>(after);
    double *part = reinterpret_cast<double *>(after + 16);
    T *scratch = reinterpret_cast<T *>(after + 16 + (size_t)nw * sizeof(double));
    T *sa = scratch + (size_t)warp * (2 * 16 * E);
    T *sb = sa + 16 * E;
    T *mine = grp ? sb : sa;

    if (STAGED) {
        if (tid == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    unsigned phase = 0;
    const HalfSortCoef<T> coef(h);
    const int base = grp ? (int)a.n0 : 0;  // class-grouped row: controls first, then cases
    const uint32_t ncls = (uint32_t)(grp ? a.n1 : a.n0);
    constexpr int RA = (16 * E + 31) / 32;  // a-values per lane in the evaluation phase

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
        const uint32_t *perm = nullptr;
        uint32_t x0 = 0, step_r = 0, step_s = 0;   // device mode: affine walk over class-relative positions
        if (a.perm_mode == PHET_PERM_DEVICE) {
            const AffinePerm ap = make_affine(a.seed, (uint64_t)g, (uint32_t)grp, ncls, grp ? a.coprime1 : a.coprime0);
            x0 = affine_at(ap, (uint64_t)warp * (uint64_t)a.m + (uint64_t)h);          // first slice of this warp
            step_r = (uint32_t)((16ull * ap.a) % ncls);                                  // next register: t += 16
            step_s = (uint32_t)(((uint64_t)ap.a * ((uint64_t)nw * (uint64_t)a.m)) % ncls);  // next slice of this warp
        } else {
            perm = (grp ? a.perm1 : a.perm0) + (g - a.perm_g0) * a.min_c;
        }
        double omin = CUDART_INF;
        for (int64_t k = warp; k < a.n_slices; k += nw) {
            const int n = (k == a.n_slices - 1) ? (int)a.n_last : (int)a.m;
            const int64_t start = k * a.m;
            T x[E];
            uint32_t xr = x0;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int p = h + 16 * r;
                if (p < n) {
                    const uint32_t q = (a.perm_mode == PHET_PERM_DEVICE) ? xr : __ldg(perm + start + p);
                    x[r] = STAGED ? lds_at(gene_sa, base + (int)q, T(0)) : __ldg(gene + base + q);
                } else {
                    x[r] = Big<T>::v();
                }
                xr = add_mod(xr, step_r, ncls);
            }
            x0 = add_mod(x0, step_s, ncls);
            halfwarp_sort_asc<T, E>(x, coef);
#pragma unroll
            for (int r = 0; r < E; ++r) mine[h * E + r] = x[r];
            __syncwarp();
            int hmax = 0;
#pragma unroll
            for (int r = 0; r < RA; ++r) {
                const int p = lane + 32 * r;
                if (p < n) {
                    const T v = sa[p];
                    const bool first = (p == 0) || (sa[p - 1] != v);
                    const bool last = (p == n - 1) || (sa[p + 1] != v);
                    // lower bound of v in b: #{b < v}.  b is padded with +BIG up to 16*E entries, so probing
                    // with power-of-two steps never needs a bounds check (pads compare >= v).
                    int lo1 = 0;
#pragma unroll
                    for (int step = next_pow2(16 * E + 1) / 2; step >= 1; step >>= 1) {
                        const int mid = lo1 + step;
                        if (mid <= 16 * E && sb[mid - 1] < v) lo1 = mid;
                    }
                    // upper bound #{b <= v}: equals the lower bound unless b holds v itself (ties)
                    int lo2 = lo1;
                    if (lo1 < n && sb[lo1] == v) {
                        lo2 = 0;
#pragma unroll
                        for (int step = next_pow2(16 * E + 1) / 2; step >= 1; step >>= 1) {
                            const int mid = lo2 + step;
                            if (mid <= 16 * E && sb[mid - 1] <= v) lo2 = mid;
                        }
                        lo2 = min(lo2, n);
                    }
                    if (last) hmax = max(hmax, abs(p + 1 - lo2));
                    if (first) hmax = max(hmax, abs(p - lo1));
                }
            }
            hmax = __reduce_max_sync(kFull, hmax);
            const double pv = (n == (int)a.m) ? __ldg(a.tab_full + hmax) : __ldg(a.tab_last + hmax);
            omin = fmin(omin, pv);
            if (a.H != nullptr && lane == 0) a.H[g * a.h_ld + k] = hmax;
            __syncwarp();
        }
        if (lane == 0) part[warp] = omin;
        __syncthreads();
        if (tid == 0) {
            double t = CUDART_INF;
            for (int w = 0; w < nw; ++w) t = fmin(t, part[w]);
            a.o[g] = t;
        }
        __syncthreads();
    }
}

template <typename T, int E, bool STAGED>
static int launch_step3h_inst(phet_ctx *c, const Step3Args &args, size_t smem_gene) {
    auto kern = k_step3h<T, E, STAGED>;
    const size_t smem = smem_gene + 16 + (size_t)(kThreads3 / 32) * sizeof(double) +
                        (size_t)(kThreads3 / 32) * 2 * 16 * E * sizeof(T);
    if (smem > (size_t)kMaxSmemOptin) return 1;  // caller retries without staging
    PHET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PHET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads3, smem));
    if (per_sm < 1) {
        set_error("step3: kernel does not fit on an SM");
        return PHET_ERR_UNSUPPORTED;
    }
    int64_t grid = (int64_t)per_sm * c->num_sms;
    const int64_t ng = args.g_end - args.g_begin;
    if (grid > ng) grid = ng;
    if (grid < 1) return 0;
    kern<<<(unsigned)grid, kThreads3, smem, c->stream>>>(args);
    c->launches++;
    c->step3_path = 2;
    PHET_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, int E>
static int launch_step3h_E(phet_ctx *c, const Step3Args &args) {
    const size_t gene = ((size_t)c->n_pad * sizeof(T) + 15) & ~(size_t)15;
    int rc = launch_step3h_inst<T, E, true>(c, args, gene);
    if (rc == 1) rc = launch_step3h_inst<T, E, false>(c, args, 0);
    if (rc == 1) {
        set_error("step3: scratch does not fit in shared memory");
        return PHET_ERR_UNSUPPORTED;
    }
    return rc;
}

template <typename T>
int run_step3_prune(phet_ctx *c, const Step3Args &args);  // step3p_impl.cuh, instantiated in step3p_f32.cu / step3p_f64.cu

// Bound-and-prune path when it applies (step3p_impl.cuh); else m <= 256: half-warp path; larger m:
// full-warp network (step3_impl.cuh).
template <typename T>
static int run_step3_any(phet_ctx *c, const Step3Args &args) {
    {
        const int rc = run_step3_prune<T>(c, args);     // step3p_impl.cuh: class vector staged, private counters
        if (rc != 1) return rc;
    }
    {
        const int rc = run_step3_scatter<T>(c, args);   // step3s_impl.cuh: streams the vectors, any class size (returns 1 when k_step3p applies)
        if (rc != 1) return rc;
    }
    if (args.cap > 256) return run_step3_T<T>(c, args);
    const int need = (int)((args.cap + 15) / 16);
    if (need <= 1) return launch_step3h_E<T, 1>(c, args);
    if (need <= 2) return launch_step3h_E<T, 2>(c, args);
    if (need <= 3) return launch_step3h_E<T, 3>(c, args);
    if (need <= 4) return launch_step3h_E<T, 4>(c, args);
    if (need <= 5) return launch_step3h_E<T, 5>(c, args);
    if (need <= 6) return launch_step3h_E<T, 6>(c, args);
    if (need <= 7) return launch_step3h_E<T, 7>(c, args);
    if (need <= 8) return launch_step3h_E<T, 8>(c, args);
    if (need <= 9) return launch_step3h_E<T, 9>(c, args);
    if (need <= 10) return launch_step3h_E<T, 10>(c, args);
    if (need <= 12) return launch_step3h_E<T, 12>(c, args);
    return launch_step3h_E<T, 16>(c, args);
}

}  // namespace phet
