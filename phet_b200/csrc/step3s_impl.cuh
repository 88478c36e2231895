// Step 3, scatter formulation of bound-and-prune.  Same contract as k_step3 for o[g]; replaces phet.py:473-498.
//
// k_step3p walks every slice's members through the permutation and GATHERS their values -- which needs the class
// vector in shared memory (impossible for config 5's 100,000-cell classes: 400 KB) and pays random-access bank
// conflicts.  The slice histograms do not care in which order the cells arrive, so this kernel turns the loop inside
// out: it STREAMS each class vector once in storage order (coalesced loads straight from L2 / HBM, nothing staged),
// asks for every cell which slice it belongs to -- the INVERSE permutation: a 16-bit table written by the
// Fisher-Yates kernel / the upload, or the inverse of the per-gene affine walk in device mode (one add per cell) --
// and bumps that slice's bucket counter with a shared-memory atomic.  Per gene:
//
//   scatter  both classes: cell -> (slice s, bucket b) -> ATOMS.ADD on hist[class][s][b]  (byte counters packed four
//            to a word, slice pitch 33 words so that equal buckets of different slices sit in different banks)
//   bounds   one thread per slice: L = max_k |D_k| <= h <= U = max_k max(D_{k-1} + ha_k, hb_k - D_{k-1})
//   prune    L* = max L over full-length slices; a slice needs its exact h only if U > L* (table non-increasing
//            from h0 on), plus a shorter/longer last slice, plus (m > 255) a slice whose byte counter wrapped
//   exact    the cells of the flagged slices (typically ~4 of 159) are collected by a second streaming pass into
//            per-slice lists (shared memory, on top of the histograms), each list is sorted by one warp and the exact
//            integer h evaluated as in k_step3h.  o[g] = min(table[L*], exact p-values).
//
// Cost per cell: one coalesced load, the bucket map (FFMA + FMNMX + F2I), one slice lookup, one shared atomic.
// 85 KB of shared memory at config 5 (317 slices) -> two CTAs per SM; any class size; slices up to 512 cells.
#pragma once

#include "smem_prims.cuh"
#include "step3p_impl.cuh"

namespace phet {

constexpr int kS3sThreads = 512;
constexpr int kS3sPitchWords = 33;   // words per (class, slice) histogram: 32 used + 1 of skew
constexpr int kS3sMaxSlots = 16;     // flagged slices evaluated per round

struct Step3ScatterArgs {
    Step3Args a;
    const float2 *win;       // [G]: (scale, offset) of the common bucket map of the gene (k_window3)
    int32_t ns_pad;          // n_slices rounded up to 4
    int32_t h0_full;
    int32_t slots;           // lists per round
    int32_t mcap;            // list capacity (max(m, n_last) rounded up to 32)
    uint32_t magic_m;        // floor(2^32 / m) + 1: t / m == umulhi(t, magic_m) for t * m < 2^32
};

// Slice of the cell at class-relative storage position q: the inverse of the permutation.
struct SliceOfTable {
    const uint16_t *tab;   // [ncls] for this gene and class
    __device__ __forceinline__ void init(uint32_t) {}
    __device__ __forceinline__ uint32_t at(uint32_t q) { return (uint32_t)__ldg(tab + q); }
};

struct SliceOfAffine {   // q = (a t + b) mod n  ->  t = ainv (q - b) mod n, walked with one modular add per step
    uint32_t ainv, b, n, step, t, min_c, magic_m, ns_last;
    __device__ __forceinline__ void init(uint32_t q0) {
        const uint32_t d = q0 >= b ? q0 - b : q0 + n - b;
        t = (uint32_t)(((uint64_t)ainv * (uint64_t)(d % n)) % n);
    }
    __device__ __forceinline__ uint32_t at(uint32_t) {  // cells are visited with a fixed stride: advance the walk
        const uint32_t tt = t;
        t = add_mod(t, step, n);
        return tt < min_c ? min(__umulhi(tt, magic_m), ns_last) : 0xFFFFu;
    }
};

template <typename T, int R, bool OVF, bool DEVICE>
__global__ void __launch_bounds__(kS3sThreads, 2) k_step3s(const Step3ScatterArgs pa) {
    extern __shared__ __align__(16) unsigned char s3s_smem[];
    const Step3Args &a = pa.a;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int m = (int)a.m, ns = (int)a.n_slices, n_last = (int)a.n_last;
    constexpr int NPAD = 32 * R;
    constexpr float kTop = 4.0f * 32 - 1.5f;

    // shared memory: [control | flagged list | slot map | overflow bitmap | list counters | histograms / lists]
    int *ctl = reinterpret_cast<int *>(s3s_smem);                       // [0] n flagged, [1] L*, [2..3] p_min bits
    unsigned long long *pmin_bits = reinterpret_cast<unsigned long long *>(ctl + 2);
    int *red = ctl + 4;                                                 // [16] per-warp maxima
    uint16_t *flagged = reinterpret_cast<uint16_t *>(red + 16);         // [ns_pad]
    unsigned char *slotmap = reinterpret_cast<unsigned char *>(flagged + pa.ns_pad);   // [ns_pad]
    uint32_t *ovf = reinterpret_cast<uint32_t *>(slotmap + pa.ns_pad);  // [(ns_pad + 31) / 32]
    int *cnt = reinterpret_cast<int *>(ovf + (pa.ns_pad + 31) / 32);    // [2][slots]
    uint32_t *hist = reinterpret_cast<uint32_t *>((reinterpret_cast<uintptr_t>(cnt + 2 * pa.slots) + 15) & ~(uintptr_t)15);
    T *lists = reinterpret_cast<T *>(hist);                             // [slots][2][mcap], after the bounds are taken
    const int hist_words = 2 * pa.ns_pad * kS3sPitchWords;
    const uint32_t hist_sa = smem_u32(hist);

    const int64_t ncl[2] = {a.n0, a.n1};
    const uint32_t ns_last = (uint32_t)(ns - 1);

    for (int64_t g = a.g_begin + blockIdx.x; g < a.g_end; g += gridDim.x) {
        const T *row = static_cast<const T *>(a.xg) + (size_t)(g - a.xg_g0) * (size_t)a.n_pad;
        const float2 wn = __ldg(pa.win + g);
        // inverse walks of this gene (device mode)
        SliceOfAffine wa[2];
        if (DEVICE) {
#pragma unroll
            for (int cls = 0; cls < 2; ++cls) {
                const uint32_t n = (uint32_t)ncl[cls];
                const uint64_t z = splitmix64(a.seed ^ (0xD1B54A32D192ED03ull * (2 * (uint64_t)g + (uint64_t)cls + 1)));
                const uint32_t ci = (uint32_t)z & (uint32_t)(kCoprimeCount - 1);
                wa[cls].ainv = __ldg((cls ? a.coprime_inv1 : a.coprime_inv0) + ci);
                wa[cls].b = (uint32_t)(z >> 32) % n;
                wa[cls].n = n;
                wa[cls].step = (uint32_t)(((uint64_t)wa[cls].ainv * (uint64_t)(kS3sThreads % n)) % n);
                wa[cls].min_c = (uint32_t)((ns - 1) * m + n_last);   // positions covered by the slices of THIS launch
                wa[cls].magic_m = pa.magic_m;
                wa[cls].ns_last = ns_last;
            }
        }
        const uint16_t *tabg = DEVICE ? nullptr : a.slice_of + (size_t)(g - a.perm_g0) * (size_t)a.slice_ld;

        // ---- clear
        for (int i = tid; i < hist_words; i += kS3sThreads) hist[i] = 0u;
        if (tid < (pa.ns_pad + 31) / 32) ovf[tid] = 0u;
        if (tid == 0) {
            ctl[0] = 0;
            *pmin_bits = 0x7FF0000000000000ull;  // +inf
        }
        __syncthreads();

        // ---- scatter: every cell of both classes into the histogram of its slice
#pragma unroll
        for (int cls = 0; cls < 2; ++cls) {
            const T *xc = row + (cls ? a.n0 : 0);
            const uint32_t n = (uint32_t)ncl[cls];
            SliceOfTable st{DEVICE ? nullptr : tabg + (cls ? a.n0 : 0)};
            if (DEVICE) wa[cls].init((uint32_t)tid % n);
            const uint32_t hb = hist_sa + 4u * (uint32_t)(cls * pa.ns_pad * kS3sPitchWords);
            auto one = [&](T v, uint32_t s) {
                if (s < (uint32_t)ns) {
                    const float t = fminf(fmaf((float)v, wn.x, wn.y), kTop);
                    uint32_t b;
                    asm("cvt.rzi.u32.f32 %0, %1;" : "=r"(b) : "f"(t));
                    const uint32_t addr = hb + 4u * (s * (uint32_t)kS3sPitchWords + (b >> 2));
                    const uint32_t sh = 8u * (b & 3u);
                    uint32_t old;
                    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(1u << sh) : "memory");
                    if (OVF && ((old >> sh) & 255u) == 255u) atomicOr(&ovf[s >> 5], 1u << (s & 31));  // this byte just wrapped
                }
            };
            uint32_t q = (uint32_t)tid;
            if (DEVICE && n < (uint32_t)kS3sThreads) {   // fewer cells than threads: the strided walk does not apply
                if (q < n) {
                    SliceOfAffine w1 = wa[cls];
                    w1.init(q);
                    one(__ldg(xc + q), w1.at(q));
                }
                continue;
            }
            for (; q + 3u * kS3sThreads < n; q += 4u * kS3sThreads) {   // four independent cells in flight
                const T v0 = __ldg(xc + q), v1 = __ldg(xc + q + kS3sThreads), v2 = __ldg(xc + q + 2 * kS3sThreads),
                        v3 = __ldg(xc + q + 3 * kS3sThreads);
                uint32_t s0, s1, s2, s3;
                if (DEVICE) {
                    s0 = wa[cls].at(q), s1 = wa[cls].at(q), s2 = wa[cls].at(q), s3 = wa[cls].at(q);
                } else {
                    s0 = st.at(q), s1 = st.at(q + kS3sThreads), s2 = st.at(q + 2 * kS3sThreads), s3 = st.at(q + 3 * kS3sThreads);
                }
                one(v0, s0);
                one(v1, s1);
                one(v2, s2);
                one(v3, s3);
            }
            for (; q < n; q += kS3sThreads) one(__ldg(xc + q), DEVICE ? wa[cls].at(q) : st.at(q));
        }
        __syncthreads();

        // ---- bounds on h of every slice from its two histograms (one thread per slice)
        int L = 0, U = 0;
        bool full_len = false, mine = false;
        for (int s0 = 0; s0 < ns; s0 += kS3sThreads) {   // (ns <= 512 in practice: one trip)
            const int s = s0 + tid;
            if (s < ns) {
                mine = true;
                const uint32_t ha = hist_sa + 4u * (uint32_t)(s * kS3sPitchWords);
                const uint32_t hbb = ha + 4u * (uint32_t)(pa.ns_pad * kS3sPitchWords);
                int D = 0;
#pragma unroll 4
                for (int w = 0; w < 32; ++w) {
                    const uint32_t xa = lds_u32(ha + 4u * (uint32_t)w), xb = lds_u32(hbb + 4u * (uint32_t)w);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int ca = (int)((xa >> (8 * k)) & 255u), cb = (int)((xb >> (8 * k)) & 255u);
                        U = max(U, max(D + ca, cb - D));
                        D += ca - cb;
                        L = max(L, abs(D));
                    }
                }
                full_len = !(s == ns - 1 && n_last != m);
                if (OVF && ((ovf[s >> 5] >> (s & 31)) & 1u)) {   // a wrapped counter: nothing read from it is trusted
                    L = 0;
                    U = 0x7fffffff;
                }
            }
        }
        int lmax = __reduce_max_sync(kFull, (mine && full_len) ? L : 0);
        if (lane == 0) red[warp] = lmax;
        __syncthreads();
        int Lstar = 0;
        for (int w = 0; w < nw; ++w) Lstar = max(Lstar, red[w]);
        if (mine) {
            const bool need = !full_len || U > Lstar || Lstar < pa.h0_full;
            if (need) flagged[atomicAdd(&ctl[0], 1)] = (uint16_t)tid;   // (ns <= 512: slice == tid)
        }
        __syncthreads();
        const int nflag = ctl[0];

        // ---- exact h of the flagged slices, `slots` of them per round: collect their cells, sort, evaluate
        for (int f0 = 0; f0 < nflag; f0 += pa.slots) {
            const int nf = min(pa.slots, nflag - f0);
            for (int i = tid; i < pa.ns_pad; i += kS3sThreads) slotmap[i] = 0xFF;
            if (tid < 2 * pa.slots) cnt[tid] = 0;
            __syncthreads();
            if (tid < nf) slotmap[flagged[f0 + tid]] = (unsigned char)tid;
            __syncthreads();
#pragma unroll
            for (int cls = 0; cls < 2; ++cls) {
                const T *xc = row + (cls ? a.n0 : 0);
                const uint32_t n = (uint32_t)ncl[cls];
                SliceOfTable st{DEVICE ? nullptr : tabg + (cls ? a.n0 : 0)};
                if (DEVICE) wa[cls].init((uint32_t)tid % n);
                auto take = [&](uint32_t q, uint32_t s) {
                    if (s < (uint32_t)ns) {
                        const uint32_t sl = slotmap[s];
                        if (sl != 0xFFu) {
                            const int p = atomicAdd(&cnt[cls * pa.slots + (int)sl], 1);
                            if (p < pa.mcap) lists[((size_t)sl * 2 + cls) * pa.mcap + p] = __ldg(xc + q);
                        }
                    }
                };
                if (DEVICE && n < (uint32_t)kS3sThreads) {
                    if ((uint32_t)tid < n) {
                        SliceOfAffine w1 = wa[cls];
                        w1.init((uint32_t)tid);
                        take((uint32_t)tid, w1.at(0));
                    }
                    continue;
                }
                for (uint32_t q = (uint32_t)tid; q < n; q += kS3sThreads) take(q, DEVICE ? wa[cls].at(q) : st.at(q));
            }
            __syncthreads();
            for (int j = warp; j < nf; j += nw) {
                const int k = (int)flagged[f0 + j];
                const int n = (k == ns - 1) ? n_last : m;
                T *la = lists + ((size_t)j * 2) * pa.mcap, *lb = la + pa.mcap;
#pragma unroll 1
                for (int cls = 0; cls < 2; ++cls) {   // one warp sorts the control list, then the case list, in place
                    T *l = cls ? lb : la;
                    T v[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int p = lane + 32 * r;
                        v[r] = (p < n) ? l[p] : Lim<T>::inf();
                    }
                    __syncwarp();
                    warp_sort_asc<T, R>(v, lane);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int p = lane + 32 * r;
                        if (p < pa.mcap) l[p] = (p < n) ? v[r] : Big<T>::v();
                    }
                }
                __syncwarp();
                const int h = ks_h_sorted<T, NPAD>(la, lb, n, lane);
                if (lane == 0) {
                    const double pv = (n == m) ? __ldg(a.tab_full + h) : __ldg(a.tab_last + h);
                    atomicMin(pmin_bits, (unsigned long long)__double_as_longlong(pv));
                }
            }
            __syncthreads();
        }
        if (tid == 0) {
            double o = __longlong_as_double((long long)*pmin_bits);
            if (Lstar >= pa.h0_full) o = fmin(o, __ldg(a.tab_full + Lstar));  // some full slice reaches at least L*
            a.o[g] = o;
        }
        __syncthreads();
    }
}

struct S3sGeometry {
    bool ok = false;
    int ns_pad = 0, slots = 0, mcap = 0;
    size_t smem = 0;
};

template <typename T>
inline S3sGeometry s3s_geometry(const Step3Args &a) {
    S3sGeometry ge;
    if (a.m < 1 || a.cap > 512 || a.n_slices < 1 || a.n_slices > kS3sThreads) return ge;
    if ((uint64_t)a.min_c * (uint64_t)a.m >= ((uint64_t)1 << 32)) return ge;   // magic division by m
    ge.ns_pad = (int)((a.n_slices + 3) / 4 * 4);
    {   // lists hold 32 R cells, R one of the sort widths instantiated below (ks_h_sorted reads 32 R padded entries)
        const int need = (int)((a.cap + 31) / 32);
        const int widths[10] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16};
        int R = 16;
        for (int i = 9; i >= 0; --i)
            if (widths[i] >= need) R = widths[i];
        ge.mcap = 32 * R;
    }
    const size_t hist_bytes = (size_t)2 * ge.ns_pad * kS3sPitchWords * 4;
    const size_t per_slot = (size_t)2 * ge.mcap * sizeof(T);
    int slots = (int)(hist_bytes / per_slot);
    if (slots > kS3sMaxSlots) slots = kS3sMaxSlots;
    if (slots < 2) slots = 2;
    ge.slots = slots;
    const size_t body = hist_bytes > per_slot * slots ? hist_bytes : per_slot * slots;
    ge.smem = 4 * sizeof(int) + 16 * sizeof(int) + (size_t)ge.ns_pad * 2 + (size_t)ge.ns_pad + ((size_t)(ge.ns_pad + 31) / 32) * 4 +
              (size_t)2 * slots * sizeof(int) + 16 + body;
    ge.ok = ge.smem <= (size_t)kMaxSmemOptin;
    return ge;
}

// True when the scatter kernel will serve this launch (the caller then provides inverse tables, not forward ones).
template <typename T>
inline bool step3_scatter_applies(const phet_ctx *c, const Step3Args &args) {
    if (c->keep_H || c->step3_impl == 1 || c->step3_impl == 2 || c->tab_h0 < 0) return false;
    // a shared atomic costs ~1 cycle per cell and SM (measured: 5.4 ms at config 4 against 3.3 ms for the gather
    // formulation, whose counters are private), so by default this kernel takes what k_step3p cannot stage
    if (c->step3_impl != 3 && step3_prune_applies<T>(c, args)) return false;
    return s3s_geometry<T>(args).ok;
}

template <typename T, int R>
static int launch_step3s_R(phet_ctx *c, const Step3Args &args, const S3sGeometry &ge) {
    const bool ovf = args.cap > 255;
    const bool dev = args.perm_mode == PHET_PERM_DEVICE;
    auto kern = ovf ? (dev ? k_step3s<T, R, true, true> : k_step3s<T, R, true, false>)
                    : (dev ? k_step3s<T, R, false, true> : k_step3s<T, R, false, false>);
    PHET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ge.smem));
    int per_sm = 0;
    PHET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kS3sThreads, ge.smem));
    if (per_sm < 1) return 1;
    const int64_t ng = args.g_end - args.g_begin;
    int64_t grid = (int64_t)per_sm * c->num_sms;
    if (grid > ng) grid = ng;
    if (grid < 1) return 0;
    if (int e = c->win3.alloc((size_t)c->G)) return e;
    k_window3<T><<<(unsigned)((ng + 7) / 8), 256, 0, c->stream>>>(args.xg, args.n_pad, args.xg_g0, args.g_begin, args.g_end,
                                                              args.n0, args.n1, 125.0f, c->win3.p);
    c->launches++;
    Step3ScatterArgs pa;
    pa.a = args;
    pa.a.H = nullptr;
    pa.win = c->win3.p;
    pa.ns_pad = ge.ns_pad;
    pa.h0_full = c->tab_h0;
    pa.slots = ge.slots;
    pa.mcap = ge.mcap;
    pa.magic_m = (uint32_t)(((uint64_t)1 << 32) / (uint64_t)args.m) + 1u;
    kern<<<(unsigned)grid, kS3sThreads, ge.smem, c->stream>>>(pa);
    c->launches++;
    c->step3_path = 4;
    PHET_CUDA(cudaGetLastError());
    return 0;
}

// Returns 1 when the path does not apply (caller falls back to the gather kernels).
template <typename T>
int run_step3_scatter(phet_ctx *c, const Step3Args &args) {
    if (!step3_scatter_applies<T>(c, args)) return 1;
    if (args.perm_mode != PHET_PERM_DEVICE && args.slice_of == nullptr) return 1;
    if (args.perm_mode == PHET_PERM_DEVICE && (args.coprime_inv0 == nullptr || args.coprime_inv1 == nullptr)) return 1;
    const S3sGeometry ge = s3s_geometry<T>(args);
    const int need = ge.mcap / 32;
    if (need <= 1) return launch_step3s_R<T, 1>(c, args, ge);
    if (need <= 2) return launch_step3s_R<T, 2>(c, args, ge);
    if (need <= 3) return launch_step3s_R<T, 3>(c, args, ge);
    if (need <= 4) return launch_step3s_R<T, 4>(c, args, ge);
    if (need <= 5) return launch_step3s_R<T, 5>(c, args, ge);
    if (need <= 6) return launch_step3s_R<T, 6>(c, args, ge);
    if (need <= 8) return launch_step3s_R<T, 8>(c, args, ge);
    if (need <= 10) return launch_step3s_R<T, 10>(c, args, ge);
    if (need <= 12) return launch_step3s_R<T, 12>(c, args, ge);
    return launch_step3s_R<T, 16>(c, args, ge);
}

}  // namespace phet
