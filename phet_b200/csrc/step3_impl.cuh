// Step 3: "discriminative power".  Replaces /root/reference/src/model/phet.py:473-498: for each
// gene, permute the control column and the case column, cut slices of length m over
// [0, min(n0, n1)), run a two-sample Kolmogorov-Smirnov test per slice, keep the smallest p.
//
// For equal slice lengths n the exact two-sided p-value of scipy.stats.ks_2samp depends only on
// the INTEGER statistic h = n*D = max_v |#{a <= v} - #{b <= v}| (scipy _stats_py.py:7834-7836),
// so the device computes h exactly (integer arithmetic on sorted samples, ties evaluated at
// the end of each tie group like the ECDF) and looks p up in a host-built table.
//
// One CTA stages one gene (same SMEM layout as step 1); each warp takes slices: gather the two
// samples through the permutation, sort both in registers, spill them to a per-warp SMEM
// scratch, and for every a-value binary-search the b sample: D is non-increasing between
// consecutive a-values, so its extremes are at a_i (counting a <= a_i, b <= a_i) and just
// below a_i (counting a < a_i, b < a_i).
#pragma once

#include "common.cuh"
#include "warp_prims.cuh"

namespace phet {

constexpr int kThreads3 = 512;

struct Step3Args {
    const void *xg;  // row of gene g starts at xg + (g - xg_g0) * n_pad elements
    int64_t n_pad, G, xg_g0;
    int64_t n0, n1, min_c, m, n_slices, n_last;   // min_c: positions covered by the slices = (n_slices - 1) m + n_last
    int64_t h_ld;  // row length of H (slices of the fit; a launch may cover fewer: see step3big.cu)
    int64_t cap;   // max(m, n_last): what a slice sorter must hold (n_last > m only in multi-class fits, phet.py:488-492)
    int64_t g_begin, g_end;
    int perm_mode;
    const uint32_t *perm0, *perm1;  // [genes from perm_g0][min_c] class-relative grouped positions (host mode, remapped at upload)
    int64_t perm_g0;                // gene of row 0 of the permutation tables (0 for uploaded tables, a block's first gene in replay mode)
    const uint32_t *coprime0, *coprime1;  // [kCoprimeCount] multipliers coprime to n0 / n1 (device mode)
    const uint32_t *coprime_inv0, *coprime_inv1;  // their inverses mod n0 / n1 (scatter kernel: cell -> permuted index)
    const uint16_t *slice_of;   // [genes from perm_g0][slice_ld]: slice of every cell in storage order, 0xFFFF = in no slice
    int64_t slice_ld;           // n0 + n1 (table modes; the scatter kernel's view of the permutation)
    uint64_t seed;
    const double *tab_full, *tab_last;
    double *o;
    int32_t *H;
};

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Device-mode permutation.  phet_set_classes lays the cells of each class out in a fixed
// pseudo-random order (a seeded Fisher-Yates shuffle on the host), so a per-gene AFFINE bijection
// of the class-relative positions,  t -> (a_g * t + b_g) mod n  with gcd(a_g, n) = 1, already
// yields a pseudo-random slice membership per gene, and consecutive t cost one add and one
// unsigned min.  a_g is drawn from a host-built list of multipliers coprime to n, b_g from a hash.
// Mirrored on the host in phet_b200/perm.py.
struct AffinePerm {
    uint32_t a, b, n;
};

__host__ __device__ __forceinline__ AffinePerm make_affine(uint64_t seed, uint64_t gene, uint32_t cls, uint32_t n,
                                                           const uint32_t *coprime) {
    const uint64_t z = splitmix64(seed ^ (0xD1B54A32D192ED03ull * (2 * gene + cls + 1)));
    AffinePerm p;
    p.a = coprime[(uint32_t)z & (uint32_t)(kCoprimeCount - 1)];
    p.b = (uint32_t)(z >> 32) % n;
    p.n = n;
    return p;
}

__host__ __device__ __forceinline__ uint32_t affine_at(const AffinePerm &p, uint64_t t) {
    return (uint32_t)(((uint64_t)p.a * t + p.b) % p.n);
}

__device__ __forceinline__ uint32_t add_mod(uint32_t x, uint32_t step, uint32_t n) {  // x, step < n < 2^31
    x += step;
    return min(x, x - n);
}

template <typename T, int E, bool STAGED>
__global__ void __launch_bounds__(kThreads3, 1) k_step3(const Step3Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nw = blockDim.x >> 5;

    const size_t gene_bytes = STAGED ? (size_t)a.n_pad * sizeof(T) : 0;
    T *gene_s = reinterpret_cast<T *>(smem_raw);
    unsigned char *after = smem_raw + ((gene_bytes + 15) & ~(size_t)15);
    uint64_t *bar = reinterpret_cast<uint64_t *>(after);
    double *part = reinterpret_cast<double *>(after + 16);
    T *scratch = reinterpret_cast<T *>(after + 16 + (size_t)nw * sizeof(double));
    T *sa = scratch + (size_t)warp * (2 * 32 * E);
    T *sb = sa + 32 * E;

    if (STAGED) {
        if (tid == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    unsigned phase = 0;
    const int n0 = (int)a.n0;

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
        AffinePerm ap0 = {1u, 0u, 1u}, ap1 = {1u, 0u, 1u};
        if (a.perm_mode == PHET_PERM_DEVICE) {
            ap0 = make_affine(a.seed, (uint64_t)g, 0u, (uint32_t)a.n0, a.coprime0);
            ap1 = make_affine(a.seed, (uint64_t)g, 1u, (uint32_t)a.n1, a.coprime1);
        }
        double omin = CUDART_INF;
        for (int64_t k = warp; k < a.n_slices; k += nw) {
            const int n = (k == a.n_slices - 1) ? (int)a.n_last : (int)a.m;
            const int64_t start = k * a.m;
            T va[E], vb[E];
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int p = lane + 32 * r;
                if (p < n) {
                    uint32_t qa, qb;
                    if (a.perm_mode == PHET_PERM_DEVICE) {
                        qa = affine_at(ap0, (uint64_t)(start + p));
                        qb = affine_at(ap1, (uint64_t)(start + p));
                    } else {
                        qa = __ldg(a.perm0 + (g - a.perm_g0) * a.min_c + start + p);
                        qb = __ldg(a.perm1 + (g - a.perm_g0) * a.min_c + start + p);
                    }
                    va[r] = STAGED ? gene[qa] : __ldg(gene + qa);
                    vb[r] = STAGED ? gene[n0 + qb] : __ldg(gene + n0 + qb);
                } else {
                    va[r] = Lim<T>::inf();
                    vb[r] = Lim<T>::inf();
                }
            }
            warp_sort_asc<T, E>(va, lane);
            warp_sort_asc<T, E>(vb, lane);
#pragma unroll
            for (int r = 0; r < E; ++r) {
                sa[lane + 32 * r] = va[r];
                sb[lane + 32 * r] = vb[r];
            }
            __syncwarp();
            int h = 0;
#pragma unroll
            for (int r = 0; r < E; ++r) {
                const int p = lane + 32 * r;
                if (p < n) {
                    const T v = va[r];
                    const bool first = (p == 0) || (sa[p - 1] != v);
                    const bool last = (p == n - 1) || (sa[p + 1] != v);
                    // lower and upper bound of v in the sorted b sample, searched together
                    int lo1 = 0, len1 = n, lo2 = 0, len2 = n;
                    while ((len1 | len2) > 0) {
                        const int h1 = len1 >> 1, h2 = len2 >> 1;
                        const T x1 = sb[min(lo1 + h1, n - 1)];
                        const T x2 = sb[min(lo2 + h2, n - 1)];
                        const bool g1 = (len1 > 0) && (x1 < v);
                        const bool g2 = (len2 > 0) && (x2 <= v);
                        lo1 = g1 ? lo1 + h1 + 1 : lo1;
                        len1 = g1 ? len1 - h1 - 1 : h1;
                        lo2 = g2 ? lo2 + h2 + 1 : lo2;
                        len2 = g2 ? len2 - h2 - 1 : h2;
                    }
                    if (last) h = max(h, abs(p + 1 - lo2));
                    if (first) h = max(h, abs(p - lo1));
                }
            }
            h = __reduce_max_sync(kFull, h);
            const double pv = (n == (int)a.m) ? __ldg(a.tab_full + h) : __ldg(a.tab_last + h);
            omin = fmin(omin, pv);
            if (a.H != nullptr && lane == 0) a.H[g * a.h_ld + k] = h;
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
static int launch_step3_inst(phet_ctx *c, const Step3Args &args, size_t smem_gene) {
    auto kern = k_step3<T, E, STAGED>;
    const size_t smem = smem_gene + 16 + (size_t)(kThreads3 / 32) * sizeof(double) +
                        (size_t)(kThreads3 / 32) * 2 * 32 * E * sizeof(T);
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
    c->step3_path = 1;
    PHET_CUDA(cudaGetLastError());
    return 0;
}

template <typename T, int E>
static int launch_step3_E(phet_ctx *c, const Step3Args &args) {
    const size_t gene = ((size_t)c->n_pad * sizeof(T) + 15) & ~(size_t)15;
    int rc = launch_step3_inst<T, E, true>(c, args, gene);
    if (rc == 1) rc = launch_step3_inst<T, E, false>(c, args, 0);
    if (rc == 1) {
        set_error("step3: scratch does not fit in shared memory");
        return PHET_ERR_UNSUPPORTED;
    }
    return rc;
}

template <typename T>
static int run_step3_T(phet_ctx *c, const Step3Args &args) {
    const int need = (int)((args.cap + 31) / 32);
    // slices of length <= 256 are served by the half-warp path (step3h_impl.cuh)
    if (need <= 10) return launch_step3_E<T, 10>(c, args);
    if (need <= 12) return launch_step3_E<T, 12>(c, args);
    if (need <= 16) return launch_step3_E<T, 16>(c, args);
    set_error("step3: slice length m > 512 is not supported by the register-resident sort");
    return PHET_ERR_UNSUPPORTED;
}

}  // namespace phet
