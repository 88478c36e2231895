// Step 1, pair-per-thread path.  Same contract as k_step1 (step1_impl.cuh); replaces phet.py:397-432
// + 435, 449-451.  Two instantiations of one kernel:
//   BIG = false  m <= 255, both classes <= 65536 cells, one class vector fits in shared memory: the
//                vector is staged with TMA and gathered from SMEM through 16-bit positions (configs 1-4);
//   BIG = true   m <= 496 or classes too large to stage (config 5: 100k-cell classes, m = 316): the
//                values are gathered from global memory through 32-bit positions.  A class vector
//                (400 KB) lives in L2 while its gene is worked on, and every subsample's positions are
//                stored in ASCENDING order, so the 16 warps of a CTA sweep the vector front to back
//                together and most gathers hit lines another lane has just pulled into L1.
//
// Why a third formulation: the warp-cooperative sorters are bound by the ALU pipe (FMNMX at one
// warp-instruction per two cycles per scheduler, ~470 of them per unit) although only four order
// statistics per sample are needed.  Here ONE THREAD owns one (subsample, class) sample and finds
// those order statistics by EXACT bucket selection instead of sorting:
//
//   pass A   gather the m values from the class vector staged in SMEM; accumulate sum / sum of
//            squares in fp64; map every value through a per-(gene, class) monotone affine map
//            to one of 127 buckets (FFMA + saturating F2I: no ALU-pipe work); count the bucket in
//            a thread-private byte histogram (bank = lane: conflict-free) and remember the bucket
//            id in a thread-private list.
//   locate   walk the 127 counters once: for each of the four sorted positions (y[j], y[j+1] of
//            the low and the high percentile) the bucket that holds it and its rank inside.
//   select   per percentile, the few elements (typically 2-5) of the bucket range that holds
//            its two positions are re-gathered (SIMD-in-register range test over the bucket
//            list, 4 ids per word), sorted with a tiny register network, and the two ranks read
//            off.  Because the map is monotone, bucket order is value order: the result is the
//            exact order statistic, bit-identical to a full sort.
//   rescue   a range with more than 16 (fp64: 8) elements is either one big tie group (all
//            equal: answer known) or is handed to a warp-cooperative full sort of that sample.
//
// The bucket map comes from k_windows: it brackets the two percentiles of the class with a
// margin of z = 4 standard errors of a subsample quantile (estimated on 512 cells of the class;
// the class-grouped layout is a fixed pseudo-random shuffle, so the first cells are a sample), so
// the interior buckets see ~1 element each.  A bad map can only cost time, never exactness.
//
// One class vector is resident at a time (class-at-a-time staging with 1-D TMA bulk copies):
// control results (sum, sum of squares, IQR) of a thread's subsamples are parked in a
// thread-private global scratch line and combined when the same thread finishes the case sample.
#pragma once

#include "smem_prims.cuh"
#include "step1_impl.cuh"

namespace phet {

constexpr int kPairMaxThreads = 512;
// Per thread: NBW histogram words = 4*NBW one-byte counters; buckets 0 .. 4*NBW-2 are used.  NBW = 32 when
// shared memory allows, 16 when that buys more resident threads (latency hiding matters more than
// bucket resolution: coarser buckets only add a few candidates per range).
constexpr uint32_t kPairPadId = 127u; // bucket-list filler of the last chunk (never a target)
constexpr int kWinSample = 512;       // cells per class sampled by k_windows

struct Step1PairArgs {
    Step1Args a;
    const void *idxT;      // [2][nblk][m][32]: class-local grouped position of element j of subsample 32*blk + lane
                           // (uint16_t, bank-spread order; BIG: uint32_t, ascending order)
    int64_t nblk;          // ceil(S_total / 32)
    const float2 *win;     // [G][2]: (scale, offset) of the bucket map of (gene, class)
    double *scratch;       // [grid][3][slots]
    int32_t slots;         // rounds * blockDim.x
    int32_t rounds;
    int32_t bl_stride;     // bytes per thread of the bucket list (16 * odd)
    int32_t vals_bytes;    // SMEM bytes reserved for one class vector (multiple of 128)
    int64_t n0, n1;
    unsigned long long *prof;  // optional [16] cycle counters (slots 8..15), env PHET_PROFILE=1, else NULL
};

// Bytes of w are bucket ids < 128.  Bit 7 of byte i of the result is set iff lo <= byte i <= hi
// (lo4 = lo * 0x01010101, hi4h = hi * 0x01010101 | 0x80808080; no borrow crosses a byte).
__device__ __forceinline__ uint32_t range_flags(uint32_t w, uint32_t lo4, uint32_t hi4h) {
    return ((w | 0x80808080u) - lo4) & (hi4h - w) & 0x80808080u;
}

// Indices of one 16-element chunk.  The index table is zero-padded to a multiple of 16 rows, so the
// loads (and the gathers they feed) need no bounds predicate.
template <typename IT>
__device__ __forceinline__ void pair_load_pos(const IT *__restrict__ ip, int j0, uint32_t (&pos)[16]) {
    const IT *ipc = ip + j0 * 32;  // element j of this thread's subsample sits at ip[32 * j]
#pragma unroll
    for (int u = 0; u < 16; ++u) pos[u] = (uint32_t)__ldg(ipc + u * 32);
}

// One 16-element chunk of pass A on gathered values.  TAIL: elements j >= m are fillers: they add 0 to
// the moments, carry the filler id in the bucket list and count in a dummy byte (the last one of the
// histogram, beyond every real bucket) -- selects, not branches.
// FASTMOM (float storage of z-scored data): moments of the deviations d = v - c from the sample's first
// element c; the chunk's sum of d and of d*d are formed in fp32 (two interleaved partial sums each) and
// added to the fp64 accumulators once per chunk: 2 conversions per chunk instead of 16 on the
// quarter-rate conversion pipe.  A sample of identical values gives exactly zero sums (so the zero-variance
// cases of the reference stay exact); otherwise the relative error of the moments is ~1e-7, against the
// 1e-3 tolerance of float storage.  fp64 storage and un-normalised data keep exact fp64 accumulation.
template <typename T, bool TAIL, bool FASTMOM>
__device__ __forceinline__ void pair_chunk(const T (&v)[16], int j0, int m, uint32_t hbase, uint32_t blbase, float scale,
                                           float off, float top, uint32_t dummy, float cshift, double (&sx)[2],
                                           double (&qx)[2]) {
    uint32_t b[16], ad[16];
    float fs[2] = {0.0f, 0.0f}, fq[2] = {0.0f, 0.0f};
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        const bool ok = !TAIL || j0 + u < m;
        if (FASTMOM) {
            const float d = ok ? (float)v[u] - cshift : 0.0f;
            fs[u & 1] += d;
            fq[u & 1] = fmaf(d, d, fq[u & 1]);
        } else {
            const T vm = ok ? v[u] : T(0);
            const double vd = (double)vm;
            sx[u & 1] += vd;
            qx[u & 1] = fma(vd, vd, qx[u & 1]);
        }
        const uint32_t bb = pair_bucket((float)v[u], scale, off, top, hbase, ad[u]);
        b[u] = ok ? bb : kPairPadId;
        if (TAIL) ad[u] = ok ? ad[u] : dummy;
    }
    if (FASTMOM) {
        sx[0] += (double)(fs[0] + fs[1]);
        qx[0] += (double)(fq[0] + fq[1]);
    }
#ifdef PHET_S1P_SERIAL_RMW
#pragma unroll
    for (int u = 0; u < 16; ++u) sts_u8(ad[u], lds_u8(ad[u]) + 1u);   // one update at a time (the LSU keeps a thread's accesses in order)
#else
#pragma unroll
    for (int u = 0; u < 16; u += 2) {
        // two histogram updates in flight: both counters are read before either is written, and the second
        // write accounts for the first when the two elements share a counter
        const uint32_t c0 = lds_u8(ad[u]), c1 = lds_u8(ad[u + 1]);
        sts_u8(ad[u], c0 + 1u);
        sts_u8(ad[u + 1], c1 + 1u + (ad[u] == ad[u + 1] ? 1u : 0u));
    }
#endif
    uint32_t pk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)  // IMAD chain: ids are < 128, bytes do not overlap
        pk[k] = ((b[4 * k + 3] * 256u + b[4 * k + 2]) * 256u + b[4 * k + 1]) * 256u + b[4 * k];
    sts_v4(blbase + (uint32_t)j0, pk[0], pk[1], pk[2], pk[3]);
}

// Ascending bitonic network on N = 2^k registers: every index is a compile-time constant.
template <typename T, int N>
__device__ __forceinline__ void sort_pow2_asc(T (&x)[N]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j >= 1; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const T lo = tmin(x[i], x[l]), hi = tmax(x[i], x[l]);
                    const bool up = (i & k) == 0;
                    x[i] = up ? lo : hi;
                    x[l] = up ? hi : lo;
                }
            }
        }
    }
}

// Ascending sort of 8 registers with the 19-comparator network (bitonic needs 24).
template <typename T>
__device__ __forceinline__ void sort8_asc(T (&x)[8]) {
#define PHET_CE(i, j) { const T lo_ = tmin(x[i], x[j]), hi_ = tmax(x[i], x[j]); x[i] = lo_; x[j] = hi_; }
    PHET_CE(0, 2) PHET_CE(1, 3) PHET_CE(4, 6) PHET_CE(5, 7)
    PHET_CE(0, 4) PHET_CE(1, 5) PHET_CE(2, 6) PHET_CE(3, 7)
    PHET_CE(0, 1) PHET_CE(2, 3) PHET_CE(4, 5) PHET_CE(6, 7)
    PHET_CE(2, 4) PHET_CE(3, 5)
    PHET_CE(1, 4) PHET_CE(3, 6)
    PHET_CE(1, 2) PHET_CE(3, 4) PHET_CE(5, 6)
#undef PHET_CE
}

// Candidate element numbers (EB = 8 or 16 bits each, newest in the lowest bits) shifted through W 32-bit registers.
template <int W, int EB>
__device__ __forceinline__ void push_entry(uint32_t (&p)[W], uint32_t j) {
#pragma unroll
    for (int i = W - 1; i >= 1; --i) p[i] = __funnelshift_l(p[i - 1], p[i], EB);
    p[0] = p[0] * (1u << EB) + j;   // (IMAD: the FMA pipe has room, the ALU pipe does not)
}
template <int W, int EB>
__device__ __forceinline__ uint32_t pack_entry(const uint32_t (&p)[W], int k) {  // k: compile-time constant after unrolling
    return (p[(k * EB) >> 5] >> ((k * EB) & 31)) & ((1u << EB) - 1u);
}

// x[r] for a runtime r < N without dynamic register indexing
template <typename T, int N>
__device__ __forceinline__ T pick_reg(const T (&x)[N], int r) {
    T y = x[0];
#pragma unroll
    for (int i = 1; i < N; ++i)
        if (r == i) y = x[i];
    return y;
}

// Where the values of the current class vector are read from: the staged copy in shared memory, or
// (BIG) the gene-major matrix itself (L2 / L1).
template <typename T, bool BIG>
struct PairSrc {
    uint32_t sa;   // shared address of element 0 (staged)
    const T *gp;   // global address of element 0 (BIG)
    __device__ __forceinline__ T at(uint32_t pos) const {
        if (BIG) return __ldg(gp + pos);
        return lds_pos(sa, pos, T(0));
    }
};

template <bool BIG> struct PairIdx { typedef uint16_t type; };
template <> struct PairIdx<true> { typedef uint32_t type; };

template <typename T>
struct PairCand {
    static constexpr int kMax = 64 / (int)sizeof(T);  // candidates per bucket range (64 bytes of the bucket-list area each)
};

// A bucket range with more than kMax elements: exact only if it is one tie group.
template <typename T, typename IT, bool BIG>
__device__ __noinline__ bool pair_big_range(const IT *__restrict__ ip, int m, const PairSrc<T, BIG> src, uint32_t blbase,
                                            uint32_t blo, uint32_t bhi, T &y) {
    const uint32_t lo4 = blo * 0x01010101u;
    const uint32_t hi4h = (bhi * 0x01010101u) | 0x80808080u;
    const int nchunk = (m + 15) >> 4;
    T vmin = Big<T>::v(), vmax = -Big<T>::v();
    for (int c = 0; c < nchunk; ++c) {
        const uint4 q = lds_v4(blbase + 16u * (uint32_t)c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t f = range_flags(k == 0 ? q.x : (k == 1 ? q.y : (k == 2 ? q.z : q.w)), lo4, hi4h);
            while (f) {
                const int bit = __ffs((int)f) - 1;
                f &= f - 1u;
                const int j = c * 16 + k * 4 + (bit >> 3);
                if (j < m) {
                    const T v = src.at((uint32_t)__ldg(ip + j * 32));
                    vmin = tmin(vmin, v);
                    vmax = tmax(vmax, v);
                }
            }
        }
    }
    y = vmin;
    return vmin == vmax;
}

template <typename T, int NBW, bool FASTMOM, bool BIG>
__global__ void __launch_bounds__(kPairMaxThreads, 1) k_step1p(const Step1PairArgs pa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    typedef typename PairIdx<BIG>::type IT;
    constexpr int CM = PairCand<T>::kMax;
    constexpr int RR = BIG ? 16 : 8;  // values per lane of the rescue sort (32 * RR >= m)
    const Step1Args &a = pa.a;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    const int m = (int)a.m;
    constexpr uint32_t kHistWarpBytes = 32u * 4u * NBW;
    constexpr float kTop = 4.0f * NBW - 1.5f;
    constexpr bool kFast = FASTMOM && sizeof(T) == 4;

    // shared memory: [class vector | mbarrier | per-warp partials | histograms (128*NBW bytes per warp) | bucket lists]
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + pa.vals_bytes);
    double *part = reinterpret_cast<double *>(bar + 2);
    unsigned char *hist_raw = reinterpret_cast<unsigned char *>(part + 2 * (kPairMaxThreads / 32));
    const uint32_t hbase = smem_u32(hist_raw) + (uint32_t)warp * kHistWarpBytes + (uint32_t)lane * 4u;
    const uint32_t blbase = smem_u32(hist_raw) + (uint32_t)nw * kHistWarpBytes + (uint32_t)tid * (uint32_t)pa.bl_stride;
    const uint32_t hdummy = hist_addr(hbase, 4u * NBW - 1u);  // counter of the chunk fillers (beyond every real bucket)

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
#pragma unroll
    for (int i = 0; i < NBW; ++i) sts_u32(hbase + 128u * (uint32_t)i, 0u);
    __syncthreads();

    unsigned phase = 0;
    const double inv_n = 1.0 / (double)m;
    const double inv_df = 1.0 / a.df;
    const double inv_pair = inv_n + inv_n;
    const double w_lo0 = 1.0 - a.q.g_lo, w_hi0 = 1.0 - a.q.g_hi;
    const int P0 = a.q.j_lo, P1 = a.q.k_lo, P2 = a.q.j_hi, P3 = a.q.k_hi;
    double *scr = pa.scratch + (size_t)blockIdx.x * 3u * (size_t)pa.slots;
    const int m_full = m & ~15;
    const int nchunk = (m + 15) >> 4;
    const size_t idx_rows = (size_t)nchunk * 16u;  // rows per block of the zero-padded index table

    for (int64_t g = a.g_begin + blockIdx.x; g < a.g_end; g += gridDim.x) {
        double facc = 0.0, racc = 0.0;
        const unsigned char *row = static_cast<const unsigned char *>(a.xg) + (size_t)(g - a.xg_g0) * (size_t)a.n_pad * sizeof(T);
        for (int cls = 0; cls < 2; ++cls) {
            const int64_t e0 = cls ? pa.n0 : 0;
            const int64_t ncls = cls ? pa.n1 : pa.n0;
            const size_t b0 = ((size_t)e0 * sizeof(T)) & ~(size_t)15;
            const size_t b1 = (((size_t)(e0 + ncls) * sizeof(T)) + 15) & ~(size_t)15;
            __syncthreads();  // every read of the previous class vector (and of part[]) is done
            if (tid == 0) {
                if (!BIG) stage_row(smem_raw, row + b0, b1 - b0, bar);
                // the vector staged next (this gene's case cells, or the control cells of the CTA's next gene)
                // starts its trip from HBM to L2 now, under the work on this one
                if (cls == 0) {
                    const size_t c0 = ((size_t)pa.n0 * sizeof(T)) & ~(size_t)15;
                    prefetch_row_l2(row + c0, ((((size_t)(pa.n0 + pa.n1) * sizeof(T)) + 15) & ~(size_t)15) - c0);
                } else if (g + gridDim.x < a.g_end) {
                    prefetch_row_l2(row + (size_t)gridDim.x * (size_t)a.n_pad * sizeof(T), (((size_t)pa.n0 * sizeof(T)) + 15) & ~(size_t)15);
                }
            }
            PairSrc<T, BIG> src;
            src.sa = smem_u32(smem_raw) + (uint32_t)((size_t)e0 * sizeof(T) - b0);
            src.gp = reinterpret_cast<const T *>(row) + e0;
            const float2 wn = __ldg(pa.win + (size_t)g * 2 + cls);
            if (!BIG) {
                mbar_wait(bar, phase);
                phase ^= 1u;
            }
            if (pa.prof != nullptr && tid == 0) atomicAdd(pa.prof + 15, 1ull);  // class passes (for averaging)

            long long t_mark = (pa.prof != nullptr && tid == 0) ? clock64() : 0;
            auto mark = [&](int slot_) {  // thread 0: cycles since the previous mark -> prof[slot_]
                if (pa.prof != nullptr && tid == 0) {
                    const long long t = clock64();
                    atomicAdd(pa.prof + slot_, (unsigned long long)(t - t_mark));
                    t_mark = t;
                }
            };
            for (int r = 0; r < pa.rounds; ++r) {
                const int slot = r * (int)blockDim.x + tid;
                const int64_t s = a.s_begin + slot;
                const bool active = s < a.s_end;
                const int64_t sa = active ? s : a.s_begin;
                const IT *ip = static_cast<const IT *>(pa.idxT) + (((size_t)cls * (size_t)pa.nblk + (size_t)(sa >> 5)) * idx_rows) * 32u + (size_t)(sa & 31);
                double sx = 0.0, qx = 0.0, iq = 0.0;
                bool rescue = false;
                if (active) {
                    double sxp[2] = {0.0, 0.0}, qxp[2] = {0.0, 0.0};
                    // ---- pass A, software-pipelined: the index loads of chunk c+1 are in flight while chunk c is processed
                    {
                        uint32_t pos[16], nxt[16];
                        pair_load_pos(ip, 0, pos);
                        const float cshift = kFast ? (float)src.at(pos[0]) : 0.0f;
                        int j0 = 0;
                        for (; j0 + 32 <= m_full; j0 += 32) {  // two chunks per trip: the index buffers swap roles
                            pair_load_pos(ip, j0 + 16, nxt);
                            {
                                T v[16];
#pragma unroll
                                for (int u = 0; u < 16; ++u) v[u] = src.at(pos[u]);
                                pair_chunk<T, false, kFast>(v, j0, m, hbase, blbase, wn.x, wn.y, kTop, hdummy, cshift, sxp, qxp);
                            }
                            if (j0 + 32 < m) pair_load_pos(ip, j0 + 32, pos);
                            {
                                T v[16];
#pragma unroll
                                for (int u = 0; u < 16; ++u) v[u] = src.at(nxt[u]);
                                pair_chunk<T, false, kFast>(v, j0 + 16, m, hbase, blbase, wn.x, wn.y, kTop, hdummy, cshift, sxp, qxp);
                            }
                        }
                        if (j0 < m_full) {  // one more full chunk
                            if (j0 + 16 < m) pair_load_pos(ip, j0 + 16, nxt);
                            T v[16];
#pragma unroll
                            for (int u = 0; u < 16; ++u) v[u] = src.at(pos[u]);
                            pair_chunk<T, false, kFast>(v, j0, m, hbase, blbase, wn.x, wn.y, kTop, hdummy, cshift, sxp, qxp);
#pragma unroll
                            for (int u = 0; u < 16; ++u) pos[u] = nxt[u];
                        }
                        if (m_full < m) {
                            T v[16];
#pragma unroll
                            for (int u = 0; u < 16; ++u) v[u] = src.at(pos[u]);
                            pair_chunk<T, true, kFast>(v, m_full, m, hbase, blbase, wn.x, wn.y, kTop, hdummy, cshift, sxp, qxp);
                        }
                        // sx: sum of the sample; qx: sum of squared deviations from its mean
                        const double sd = sxp[0] + sxp[1], qd = qxp[0] + qxp[1];
                        double ss = qd - sd * sd * inv_n;
                        if (!kFast && fabs(ss) <= 1e-15 * qd) ss = 0.0;  // identical values: exactly 0 like the two-pass variance
                        if (kFast && ss < 0.0) ss = 0.0;  // fp32 rounding of a (near-)zero variance
                        sx = sd + (double)m * (double)cshift;
                        qx = ss;
                    }
                    mark(8);
                    // ---- locate the four sorted positions in the histogram
                    // sum of the four counters of a word: m <= 255 keeps it below 256 (one IMAD); BIG needs the exact sum
                    auto sum4 = [](uint32_t w) -> uint32_t { return BIG ? __dp4a(w, 0x01010101u, 0u) : (w * 0x01010101u) >> 24; };
                    int B[4], E[4], C[4];
                    const int P[4] = {P0, P1, P2, P3};
                    uint32_t wi = 0, x = 0u, s4 = 0u;
                    int cum = 0;
                    if (!BIG) {
                        // staged instantiation: all words are loaded up front and the four searches are straight-line code
                        // (the word-by-word walk below is a chain of dependent shared-memory loads)
                        uint32_t s4v[NBW];
                        int ps[NBW];
#pragma unroll
                        for (int i = 0; i < NBW; ++i) s4v[i] = sum4(lds_u32(hbase + 128u * (uint32_t)i));
                        int run0 = 0;
#pragma unroll
                        for (int i = 0; i < NBW; ++i) {
                            run0 += (int)s4v[i];
                            ps[i] = run0;
                        }
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            uint32_t wt = 0;
                            int ct = 0;
#pragma unroll
                            for (int i = 0; i < NBW - 1; ++i) {
                                const bool below = ps[i] <= P[t];
                                wt += below ? 1u : 0u;
                                ct += below ? (int)s4v[i] : 0;
                            }
                            const uint32_t xw = lds_u32(hbase + 128u * wt);
                            const int c0 = (int)(xw & 255u), c1 = (int)((xw >> 8) & 255u), c2 = (int)((xw >> 16) & 255u), c3 = (int)(xw >> 24);
                            const int r0 = ct + c0, r1 = r0 + c1, r2 = r1 + c2;
                            const int q = (r0 <= P[t] ? 1 : 0) + (r1 <= P[t] ? 1 : 0) + (r2 <= P[t] ? 1 : 0);
                            B[t] = (int)(wt * 4u) + q;
                            E[t] = q == 0 ? ct : (q == 1 ? r0 : (q == 2 ? r1 : r2));
                            C[t] = q == 0 ? c0 : (q == 1 ? c1 : (q == 2 ? c2 : c3));
                        }
                    } else {
                        x = lds_u32(hbase);
                        s4 = sum4(x);
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            while (cum + (int)s4 <= P[t] && wi < (uint32_t)(NBW - 1)) {
                                cum += (int)s4;
                                ++wi;
                                x = lds_u32(hbase + 128u * wi);
                                s4 = sum4(x);
                            }
                            int run = cum;
                            uint32_t q = 0, c = x & 255u;
                            while (run + (int)c <= P[t] && q < 3u) {
                                run += (int)c;
                                ++q;
                                c = (x >> (8u * q)) & 255u;
                            }
                            B[t] = (int)(wi * 4u + q);
                            E[t] = run;
                            C[t] = (int)c;
                        }
                    }
                    // BIG (m > 255 possible): a one-byte counter wraps when 256 or more elements share a bucket (a big tie
                    // group); the counters then no longer add up to the number of list entries and the sample goes
                    // to the full sort below.  Nothing read from a wrapped histogram is trusted (all uses are bounded).
                    bool wrapped = false;
                    if (BIG) {
                        int tot = cum + (int)s4;
                        for (uint32_t w2 = wi + 1u; w2 < (uint32_t)NBW; ++w2) tot += (int)sum4(lds_u32(hbase + 128u * w2));
                        wrapped = tot != nchunk * 16;
                    }
                    const int cntL = E[1] + C[1] - E[0], cntH = E[3] + C[3] - E[2];
                    const bool smallL = cntL <= CM, smallH = cntH <= CM;
                    // ---- a range too big to enumerate is exact only if it is one tie group (needs the intact bucket list)
                    T tieL = T(0), tieH = T(0);
                    if (!smallL) rescue = !pair_big_range<T, IT, BIG>(ip, m, src, blbase, (uint32_t)B[0], (uint32_t)B[1], tieL);
                    if (!smallH) rescue = !pair_big_range<T, IT, BIG>(ip, m, src, blbase, (uint32_t)B[2], (uint32_t)B[3], tieH) || rescue;
                    rescue = rescue || wrapped;
                    mark(9);
                    T y0 = T(0), y1 = T(0), y2 = T(0), y3 = T(0);
                    {
                        // ---- one pass over the bucket list turns the words holding a candidate of either bucket range into
                        //      records; the records are decoded into element numbers that are shifted into packs held in
                        //      REGISTERS (8 bits per element number, 16 when m can exceed 255; newest in the lowest bits) -- no
                        //      element lists in shared memory.  Up to CM candidates per range (16 floats / 8 doubles).
                        constexpr int EB = BIG ? 16 : 8;
                        constexpr int PW = CM * EB / 32;
                        uint32_t pkL[PW], pkH[PW];
#pragma unroll
                        for (int i = 0; i < PW; ++i) pkL[i] = pkH[i] = 0u;
                        int nL = 0, nH = 0;
                        {
                            // a range that is too big is matched by nothing here (lo > hi) and was handled above
                            const uint32_t lo4L = (smallL ? (uint32_t)B[0] : 1u) * 0x01010101u;
                            const uint32_t hi4L = ((smallL ? (uint32_t)B[1] : 0u) * 0x01010101u) | 0x80808080u;
                            const uint32_t lo4H = (smallH ? (uint32_t)B[2] : 1u) * 0x01010101u;
                            const uint32_t hi4H = ((smallH ? (uint32_t)B[3] : 0u) * 0x01010101u) | 0x80808080u;
                            // records: low-range flags in bit 7 of each byte, high-range flags in bit 6, word number in bits 0-5
                            // (BIG: < 128 words, bit 6 of the number goes to bit 8), compacted in place over the part of the
                            // list already consumed
                            int nrec = 0;
                            for (int c = 0; c < nchunk; ++c) {
                                const uint4 q = lds_v4(blbase + 16u * (uint32_t)c);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint32_t w = (k == 0 ? q.x : (k == 1 ? q.y : (k == 2 ? q.z : q.w)));
                                    const uint32_t wh = w | 0x80808080u;
                                    const uint32_t fl = (wh - lo4L) & (hi4L - w) & 0x80808080u;
                                    const uint32_t fh = (wh - lo4H) & (hi4H - w) & 0x80808080u;
                                    const uint32_t f = fl | (fh >> 1);
                                    if (f) {
                                        const uint32_t wno = (uint32_t)(4 * c + k);
                                        sts_u32(blbase + 4u * (uint32_t)nrec, f | (BIG ? ((wno & 63u) | ((wno >> 6) << 8)) : wno));
                                        ++nrec;
                                    }
                                }
                            }
                            mark(10);
                            // records -> element numbers, shifted into the packs of their range
                            for (int i = 0; i < nrec; ++i) {
                                const uint32_t rec = lds_u32(blbase + 4u * (uint32_t)i);
                                const uint32_t jb = (BIG ? ((rec & 63u) | ((rec >> 2) & 0x40u)) : (rec & 63u)) * 4u;
                                uint32_t f = rec & 0xC0C0C0C0u;
                                do {
                                    const int bit = __ffs((int)f) - 1;
                                    f &= f - 1u;
                                    const uint32_t j = jb + (uint32_t)(bit >> 3);
                                    if (bit & 1) {   // bit 7 of a byte: low range
                                        push_entry<PW, EB>(pkL, j);
                                        ++nL;
                                    } else {
                                        push_entry<PW, EB>(pkH, j);
                                        ++nH;
                                    }
                                } while (f);
                            }
                        }
                        rescue = rescue || (smallL && nL != cntL) || (smallH && nH != cntH);  // second part never expected: guard
                        mark(11);
                        // ---- the first 8 candidates of each range: index loads, then gathers, all in flight together
                        T cL[8], cH[8];
                        {
                            uint32_t posL[8], posH[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                posL[k] = (k < nL) ? (uint32_t)__ldg(ip + pack_entry<PW, EB>(pkL, k) * 32u) : 0u;
                                posH[k] = (k < nH) ? (uint32_t)__ldg(ip + pack_entry<PW, EB>(pkH, k) * 32u) : 0u;
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const T vl = src.at(posL[k]), vh = src.at(posH[k]);
                                cL[k] = (k < nL) ? vl : Big<T>::v();
                                cH[k] = (k < nH) ? vh : Big<T>::v();
                            }
                        }
                        sort8_asc<T>(cL);
                        sort8_asc<T>(cH);
                        const int raL = P0 - E[0], rbL = P1 - E[0], raH = P2 - E[2], rbH = P3 - E[2];
                        y0 = smallL ? pick_reg<T, 8>(cL, raL) : tieL;
                        y1 = smallL ? pick_reg<T, 8>(cL, rbL) : tieL;
                        y2 = smallH ? pick_reg<T, 8>(cH, raH) : tieH;
                        y3 = smallH ? pick_reg<T, 8>(cH, rbH) : tieH;
                        mark(12);
                        if (CM > 8 && (nL > 8 || nH > 8)) {   // 9..16 candidates in a range (a few per cent of the samples)
                            // the 8 newest candidates are sorted already: the others (slots 8 .. nc - 1 of the pack, usually one
                            // or two) are inserted into the sorted run one at a time
#pragma unroll 1
                            for (int h = 0; h < 2; ++h) {
                                const int nc = h ? nH : nL;
                                if (nc <= 8 || nc > CM) continue;
                                T c16[16];
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    c16[k] = h ? cH[k] : cL[k];
                                    c16[8 + k] = Big<T>::v();
                                }
                                uint32_t pk[PW];
#pragma unroll
                                for (int i = 0; i < PW; ++i) pk[i] = h ? pkH[i] : pkL[i];
                                // all their index loads in flight together; the values wait in the (consumed) histogram words
                                uint32_t pe[8];
#pragma unroll
                                for (int u = 0; u < 8; ++u)
                                    pe[u] = (8 + u < nc) ? (uint32_t)__ldg(ip + pack_entry<PW, EB>(pk, 8 + u < CM ? 8 + u : 0) * 32u) : 0u;
#pragma unroll
                                for (int u = 0; u < 8; ++u)
                                    if (8 + u < nc) sts_val(hbase + 128u * (uint32_t)(u * (int)(sizeof(T) / 4)), src.at(pe[u]));
#pragma unroll 1
                                for (int e = 8; e < nc; ++e) {
                                    T x = lds_val(hbase + 128u * (uint32_t)((e - 8) * (int)(sizeof(T) / 4)), T(0));
#pragma unroll
                                    for (int k = 0; k < 16; ++k) {   // x takes its place, the larger values move up
                                        const T lo_ = tmin(c16[k], x);
                                        x = tmax(c16[k], x);
                                        c16[k] = lo_;
                                    }
                                }
                                const T ya = pick_reg<T, 16>(c16, h ? raH : raL), yb = pick_reg<T, 16>(c16, h ? rbH : rbL);
                                if (h) {
                                    y2 = ya;
                                    y3 = yb;
                                } else {
                                    y0 = ya;
                                    y1 = yb;
                                }
                            }
                        }
                    }
                    const double q_lo = __dadd_rn(__dmul_rn(w_lo0, (double)y0), __dmul_rn(a.q.g_lo, (double)y1));
                    const double q_hi = __dadd_rn(__dmul_rn(w_hi0, (double)y2), __dmul_rn(a.q.g_hi, (double)y3));
                    iq = q_hi - q_lo;
#pragma unroll
                    for (int i = 0; i < NBW; ++i) sts_u32(hbase + 128u * (uint32_t)i, 0u);
                }
                mark(13);
                // ---- rescue: warp-cooperative full sort of the samples that could not be bucket-selected
                unsigned todo = __ballot_sync(kFull, rescue);
                while (todo) {
                    const int L = __ffs((int)todo) - 1;
                    todo &= todo - 1u;
                    const int64_t sL = __shfl_sync(kFull, s, L);
                    const IT *ipL = static_cast<const IT *>(pa.idxT) + (((size_t)cls * (size_t)pa.nblk + (size_t)(sL >> 5)) * idx_rows) * 32u + (size_t)(sL & 31);
                    T v[RR];
#pragma unroll
                    for (int rr = 0; rr < RR; ++rr) {
                        const int j = lane + 32 * rr;
                        v[rr] = (j < m) ? src.at((uint32_t)__ldg(ipL + j * 32)) : Lim<T>::inf();
                    }
                    warp_sort_asc<T, RR>(v, lane);
                    const double iqL = iqr_of_sorted<T, RR>(v, a.q);
                    if (lane == L) iq = iqL;
                }
                if (active) {
                    if (cls == 0) {
                        scr[slot] = sx;
                        scr[pa.slots + slot] = qx;
                        scr[2 * pa.slots + slot] = iq;
                    } else {
                        const double sx0 = scr[slot], ss0 = scr[pa.slots + slot], iq0 = scr[2 * pa.slots + slot];
                        const double svar = (ss0 + qx) * inv_df;
                        const double dmean = (sx0 - sx) * inv_n;
                        const double stat = dmean * rsqrt(svar * inv_pair);
                        double delta = fabs(iq0 - iq);
                        if (!isfinite(delta)) delta = 0.0;  // phet.py:416
                        racc += delta;
                        double p = (a.test == PHET_TEST_Z) ? p_two_sided_z(stat) : p_two_sided_t(stat, a.df, a.ln_beta);
                        if (p <= a.p_flush) p = 0.0;
                        facc += fisher_term(p);
                        if (a.dbgD != nullptr) a.dbgD[g * a.dbg_ld + (s - a.s_begin)] = delta;
                        if (a.dbgP != nullptr) a.dbgP[g * a.dbg_ld + (s - a.s_begin)] = (p == p && !isinf(p)) ? p : 0.0;
                    }
                }
                mark(14);
            }
        }
        facc = warp_sum(facc);
        racc = warp_sum(racc);
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
    }
}

// Bucket map per (gene, class): one warp sorts a sample of the class and brackets the two
// percentiles.  fa / fb: lower / upper probability positions before the sampling margin.
template <typename T>
__global__ void __launch_bounds__(256) k_windows(const void *__restrict__ xg, int64_t n_pad, int64_t xg_g0, int64_t g_begin,
                                                 int64_t g_end, int64_t n0, int64_t n1, float p_lo, float p_hi, float m,
                                                 float z, float interior, float2 *__restrict__ win) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t g = g_begin + (wid >> 1);
    const int cls = (int)(wid & 1);
    if (g >= g_end) return;
    const int64_t ncls = cls ? n1 : n0;
    const int ns = (int)(ncls < kWinSample ? ncls : kWinSample);
    const T *gene = static_cast<const T *>(xg) + (g - xg_g0) * n_pad + (cls ? n0 : 0);
    T v[kWinSample / 32];
#pragma unroll
    for (int r = 0; r < kWinSample / 32; ++r) {
        const int k = lane + 32 * r;
        v[r] = (k < ns) ? __ldg(gene + k) : Lim<T>::inf();
    }
    warp_sort_asc<T, kWinSample / 32>(v, lane);
    const float fns = (float)ns;
    const float var = 1.0f / m + 1.0f / fns;
    const float fa = p_lo - z * sqrtf(fmaxf(p_lo * (1.0f - p_lo), 0.0f) * var) - 1.0f / fns;
    const float fb = p_hi + z * sqrtf(fmaxf(p_hi * (1.0f - p_hi), 0.0f) * var) + 1.0f / fns;
    int ra = (int)floorf(fa * fns), rb = (int)ceilf(fb * fns);
    ra = max(0, min(ra, ns - 1));
    rb = max(0, min(rb, ns - 1));
    float lo = (float)warp_pick<T, kWinSample / 32>(v, ra);
    float hi = (float)warp_pick<T, kWinSample / 32>(v, rb);
    if (!(hi > lo)) {  // the bracket collapsed onto a tie group: fall back to the sample's range
        lo = (float)warp_pick<T, kWinSample / 32>(v, 0);
        hi = (float)warp_pick<T, kWinSample / 32>(v, ns - 1);
    }
    float scale = 0.0f, off = 1.0f;
    if (hi > lo) {
        scale = interior / (hi - lo);  // [lo, hi) -> buckets 1 .. interior; below -> 0, above -> interior + 1
        off = 1.0f - lo * scale;
        if (!isfinite(scale) || !isfinite(off)) {
            scale = 0.0f;
            off = 1.0f;
        }
    }
    if (lane == 0) win[(size_t)g * 2 + cls] = make_float2(scale, off);
}

struct PairGeometry {
    int threads = 0, rounds = 0, bl_stride = 0, vals_bytes = 0, nbw = 0;
    size_t smem = 0;
    bool big = false;
};
constexpr int kPairBigMaxM = 496;  // 31 chunks: word numbers < 128, element numbers < 512, rescue sort of 32 * 16

// Shared-memory budget and thread count for S_local subsamples.  threads == 0: not eligible.
// Among the two histogram widths, the one that needs fewer rounds wins (more resident threads);
// on a tie the finer histogram (fewer candidates per range).
template <typename T>
static PairGeometry pair_geometry(int64_t n0, int64_t n1, int64_t m, int64_t S_local, int force_nbw = 0, bool big = false,
                                  int max_threads = kPairMaxThreads) {
    PairGeometry best;
    best.big = big;
    if (m < 1 || S_local < 1) return best;
    if (big ? (m > kPairBigMaxM) : (m > 255 || n0 > 65536 || n1 > 65536)) return best;
    const int64_t nmax = n0 > n1 ? n0 : n1;
    const int vals_bytes = big ? 0 : (int)((nmax * (int64_t)sizeof(T) + 32 + 127) / 128 * 128);
    int chunks = (int)((m + 15) / 16);
    if (chunks < 9) chunks = 9;       // the first 128 bytes double as the candidate-value area
    if ((chunks & 1) == 0) ++chunks;  // stride = 16 bytes * odd: 128-bit bucket-list accesses are conflict-free
    const int bl_stride = 16 * chunks;
    const size_t fixed = (size_t)vals_bytes + 16 + 2 * (kPairMaxThreads / 32) * sizeof(double);
    for (int nbw = 32; nbw >= 16; nbw -= 16) {
        if (force_nbw && nbw != force_nbw) continue;
        const size_t per_thread = (size_t)(4 * nbw) + (size_t)bl_stride;
        if (fixed + 64 * per_thread > (size_t)kMaxSmemOptin) continue;
        int tmax = (int)(((size_t)kMaxSmemOptin - fixed) / per_thread) / 32 * 32;
        if (tmax > max_threads) tmax = max_threads;
        const int rounds = (int)((S_local + tmax - 1) / tmax);
        int t = (int)((S_local + rounds - 1) / rounds);
        t = (t + 31) / 32 * 32;
        if (best.threads == 0 || rounds < best.rounds) {
            best.threads = t;
            best.rounds = rounds;
            best.bl_stride = bl_stride;
            best.vals_bytes = vals_bytes;
            best.nbw = nbw;
            best.smem = fixed + (size_t)t * per_thread;
        }
    }
    return best;
}

// Returns 1 when the pair-per-thread path does not apply (caller falls back to the warp sorters).
template <typename T>
static int run_step1_pair(phet_ctx *c, const Step1Args &args) {
    if (c->step1_impl == 1) return 1;
    const int64_t S_local = args.s_end - args.s_begin;
    PairGeometry ge;
    if (c->idxT_ok && !c->pair_force_big) ge = pair_geometry<T>(c->n0, c->n1, args.m, S_local, c->pair_nbw);
    if (ge.threads == 0 && c->idxT32_ok)  // no staged geometry: gather from global memory (config-5 regime)
        ge = pair_geometry<T>(c->n0, c->n1, args.m, S_local, c->pair_nbw, true, c->pair_big_threads);
    if (ge.threads == 0) return 1;
    const bool fast = c->fast_moments && sizeof(T) == 4;
    auto kern = ge.big ? (ge.nbw == 32 ? (fast ? k_step1p<T, 32, true, true> : k_step1p<T, 32, false, true>)
                                       : (fast ? k_step1p<T, 16, true, true> : k_step1p<T, 16, false, true>))
                       : (ge.nbw == 32 ? (fast ? k_step1p<T, 32, true, false> : k_step1p<T, 32, false, false>)
                                       : (fast ? k_step1p<T, 16, true, false> : k_step1p<T, 16, false, false>));
    PHET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ge.smem));
    int per_sm = 0;
    PHET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ge.threads, ge.smem));
    if (per_sm < 1) return 1;
    const int64_t ng = args.g_end - args.g_begin;
    if (ge.big) per_sm = 1;  // a second CTA would bring a second class vector into the same L1 (measured: 4x slower)
    int64_t grid = (int64_t)per_sm * c->num_sms;
    if (grid > ng) grid = ng;
    if (grid < 1) return 0;

    if (int e = c->win.alloc((size_t)c->G * 2)) return e;
    const int32_t slots = ge.rounds * ge.threads;
    if (int e = c->pair_scratch.alloc((size_t)grid * 3 * (size_t)slots)) return e;

    {   // bucket maps of the genes of this launch
        const int64_t warps = ng * 2;
        const unsigned blocks = (unsigned)((warps + 7) / 8);
        const float mf = (float)args.m;
        k_windows<T><<<blocks, 256, 0, c->stream>>>(args.xg, args.n_pad, args.xg_g0, args.g_begin, args.g_end, c->n0, c->n1,
                                                   (float)args.q.j_lo / mf, (float)(args.q.k_hi + 1) / mf, mf, 4.0f,
                                                   (float)(4 * ge.nbw - 3), c->win.p);
        c->launches++;
    }
    Step1PairArgs pa;
    pa.a = args;
    pa.idxT = ge.big ? static_cast<const void *>(c->idxT32.p) : static_cast<const void *>(c->idxT.p);
    pa.nblk = (c->S + 31) / 32;
    pa.win = c->win.p;
    pa.scratch = c->pair_scratch.p;
    pa.slots = slots;
    pa.rounds = ge.rounds;
    pa.bl_stride = ge.bl_stride;
    pa.vals_bytes = ge.vals_bytes;
    pa.n0 = c->n0;
    pa.n1 = c->n1;
    pa.prof = c->prof.p;
    c->step1_geom[0] = (int32_t)grid;
    c->step1_geom[1] = ge.threads;
    c->step1_geom[2] = (int32_t)ge.smem;
    c->step1_geom[3] = 4 * ge.nbw - 1;  // buckets
    c->step1_geom[4] = (ge.big ? 0 : 1) | 4;  // bit 0: staged; bit 2: pair-per-thread bucket-select path
    kern<<<(unsigned)grid, ge.threads, ge.smem, c->stream>>>(pa);
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace phet
