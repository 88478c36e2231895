// Step 3, a last slice longer than the warp sorters hold (multi-class fits only).
// With K >= 3 classes the reference cuts slice starts over [0, min_size) of the SMALLEST class but lets the last
// slice run to the end of the shorter column of the pair (/root/reference/src/model/phet.py:488-492), so for
// unbalanced classes that one slice holds hundreds to thousands of cells.  One CTA per gene: both columns of the
// slice are gathered into shared memory, sorted by a block-wide bitonic network, and the integer KS statistic
// h = max |#{a <= v} - #{b <= v}| is read off with paired binary searches exactly as in k_step3; the p-value comes
// from the host table for n_last and is folded into o[g] (min over slices, phet.py:496).
#include "step3_impl.cuh"

namespace phet {

constexpr int kBigThreads = 512;

template <typename T>
__device__ __forceinline__ void block_bitonic_pair(T *sa, T *sb, int P, int tid) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < P; t += kBigThreads) {   // t < P/2: array a, else array b
                T *s = (t < (P >> 1)) ? sa : sb;
                const int u = (t < (P >> 1)) ? t : t - (P >> 1);
                const int i = ((u & ~(j - 1)) << 1) | (u & (j - 1));   // insert a 0 at bit log2(j)
                const int l = i | j;
                const bool asc = (i & k) == 0;
                const T x = s[i], y = s[l];
                if ((x > y) == asc) {
                    s[i] = y;
                    s[l] = x;
                }
            }
            __syncthreads();
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kBigThreads, 1) k_ks_big(const Step3Args a, int P, int fold) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sa = reinterpret_cast<T *>(smem_raw);
    T *sb = sa + P;
    __shared__ int hred[kBigThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = (int)a.n_last;
    const int64_t start = (a.n_slices - 1) * a.m;
    const int n0 = (int)a.n0;
    for (int64_t g = a.g_begin + blockIdx.x; g < a.g_end; g += gridDim.x) {
        const T *gene = static_cast<const T *>(a.xg) + (g - a.xg_g0) * a.n_pad;
        AffinePerm ap0 = {1u, 0u, 1u}, ap1 = {1u, 0u, 1u};
        if (a.perm_mode == PHET_PERM_DEVICE) {
            ap0 = make_affine(a.seed, (uint64_t)g, 0u, (uint32_t)a.n0, a.coprime0);
            ap1 = make_affine(a.seed, (uint64_t)g, 1u, (uint32_t)a.n1, a.coprime1);
        }
        __syncthreads();
        for (int p = tid; p < P; p += kBigThreads) {
            T va = Lim<T>::inf(), vb = Lim<T>::inf();
            if (p < n) {
                uint32_t qa, qb;
                if (a.perm_mode == PHET_PERM_DEVICE) {
                    qa = affine_at(ap0, (uint64_t)(start + p));
                    qb = affine_at(ap1, (uint64_t)(start + p));
                } else {
                    qa = __ldg(a.perm0 + (g - a.perm_g0) * a.min_c + start + p);
                    qb = __ldg(a.perm1 + (g - a.perm_g0) * a.min_c + start + p);
                }
                va = __ldg(gene + qa);
                vb = __ldg(gene + n0 + qb);
            }
            sa[p] = va;
            sb[p] = vb;
        }
        __syncthreads();
        block_bitonic_pair<T>(sa, sb, P, tid);
        int h = 0;
        for (int p = tid; p < n; p += kBigThreads) {
            const T v = sa[p];
            const bool first = (p == 0) || (sa[p - 1] != v);
            const bool last = (p == n - 1) || (sa[p + 1] != v);
            int lo1 = 0, len1 = n, lo2 = 0, len2 = n;   // lower / upper bound of v in b
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
        h = __reduce_max_sync(kFull, h);
        if (lane == 0) hred[warp] = h;
        __syncthreads();
        if (tid == 0) {
            int hm = 0;
            for (int w = 0; w < kBigThreads / 32; ++w) hm = max(hm, hred[w]);
            const double pv = __ldg(a.tab_last + hm);
            a.o[g] = fold ? fmin(a.o[g], pv) : pv;
            if (a.H != nullptr) a.H[g * a.h_ld + (a.n_slices - 1)] = hm;
        }
    }
}

// last slice of every gene in [g_begin, g_end); fold = 1: o[g] already holds the minimum over the earlier slices
int launch_ks_big(phet_ctx *c, const Step3Args &a, int fold) {
    int P = 1;
    while (P < a.n_last) P <<= 1;
    if (P < 64) P = 64;
    const size_t e_st = c->storage == PHET_F32 ? 4 : 8;
    const size_t smem = 2 * (size_t)P * e_st;
    if (smem > (size_t)kMaxSmemOptin - 256) {
        set_error("unsupported: last Step-3 slice of " + std::to_string(a.n_last) + " cells does not fit the block sorter (" +
                  std::to_string((kMaxSmemOptin - 256) / (2 * e_st)) + " cells, rounded down to a power of two)");
        return PHET_ERR_UNSUPPORTED;
    }
    const int64_t ng = a.g_end - a.g_begin;
    if (ng < 1) return PHET_OK;
    int64_t grid = (int64_t)c->num_sms * (smem <= 100 * 1024 ? 2 : 1);
    if (grid > ng) grid = ng;
    if (c->storage == PHET_F32) {
        PHET_CUDA(cudaFuncSetAttribute(k_ks_big<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_ks_big<float><<<(unsigned)grid, kBigThreads, smem, c->stream>>>(a, P, fold);
    } else {
        PHET_CUDA(cudaFuncSetAttribute(k_ks_big<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_ks_big<double><<<(unsigned)grid, kBigThreads, smem, c->stream>>>(a, P, fold);
    }
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    return PHET_OK;
}

}  // namespace phet
