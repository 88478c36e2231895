// Step 3, bound-and-prune path (slice length m <= 255, one class vector fits in shared memory, H not
// requested).  Same contract as k_step3 (step3_impl.cuh) for o[g]; replaces phet.py:473-498.
//
// o[g] = min over slices of the KS p-value, and p is (essentially) non-increasing in the integer
// statistic h, so only the slices that can reach the gene's largest h need an exact h.  Per gene:
//
//   phase 1  (control vector staged)  P = 3 lanes per slice (every P-th element each) walk its permutation
//            slice, map every value through a per-gene monotone affine map to one of 127 buckets and
//            count it in a lane-private byte histogram (bank == lane: conflict-free), as in k_step1p;
//            the P partial histograms are summed into ha and parked.
//   phase 2  (case vector staged)  same for hb, then one pass over the 127 buckets gives rigorous
//            bounds on h: with D_k = sum_{i<=k} (ha_i - hb_i) the ECDF difference at the bucket edges,
//            L = max_k |D_k| <= h <= U = max_k max(D_{k-1} + ha_k, hb_k - D_{k-1})
//            (inside a bucket the difference moves by at most +ha_k / -hb_k from D_{k-1}).
//   prune    L* = max over full-length slices of L.  A slice with U <= L* cannot beat L*; since the
//            table is non-increasing from h0 on (host-checked) its p is >= table[L*].  The rest
//            (typically a few of ~160), and always a shorter last slice, are flagged.
//   exact    flagged slices only, one per warp trip: the half-warp merge-split sorter sorts the control
//            sample (lanes 0-15; values gathered from L2, where the row just staged still sits) and the
//            case sample (lanes 16-31, from SMEM) at once, and the exact integer h is evaluated as in
//            k_step3h.  o[g] = min(table[L*], exact p-values).
//
// Bucket maps come from k_window3 (one map per gene for both classes: the two ECDFs are compared
// on common edges).  A poor map only flags more slices; the result is always the exact minimum.
#pragma once

#include "smem_prims.cuh"
#include "step3h_impl.cuh"

namespace phet {

constexpr int kS3pThreads = 512;
constexpr int kS3pHistBytes = 128;  // per slice and class: 127 one-byte counters

struct Step3PairArgs {
    Step3Args a;
    const float2 *win;  // [G]: (scale, offset) of the common bucket map of the gene
    int32_t vals_bytes; // SMEM bytes reserved for one class vector (multiple of 128)
    int32_t ns_pad;     // n_slices rounded up to 32
    int32_t h0_full;    // table_full is non-increasing on [h0, m] and table[h0] <= table[h] for h < h0
    int32_t nwx;        // warps taking part in the exact phase (each owns a 4-sample scratch)
    unsigned long long *prof;  // optional [8] cycle counters summed over CTAs (env PHET_PROFILE=1), else NULL
};

// Exact integer KS statistic of two ascending samples of length n held in shared memory (sb padded
// with +BIG up to npad entries): all 32 lanes take a-values; see k_step3h for the derivation.
template <typename T, int NPAD>
__device__ __forceinline__ int ks_h_sorted(const T *__restrict__ sa, const T *__restrict__ sb, int n, int lane) {
    int hmax = 0;
    constexpr int RA = (NPAD + 31) / 32;
#pragma unroll
    for (int r = 0; r < RA; ++r) {
        const int p = lane + 32 * r;
        if (p < n) {
            const T v = sa[p];
            const bool first = (p == 0) || (sa[p - 1] != v);
            const bool last = (p == n - 1) || (sa[p + 1] != v);
            int lo1 = 0;
#pragma unroll
            for (int step = next_pow2(NPAD + 1) / 2; step >= 1; step >>= 1) {
                const int mid = lo1 + step;
                if (mid <= NPAD && sb[mid - 1] < v) lo1 = mid;
            }
            int lo2 = lo1;
            if (lo1 < n && sb[lo1] == v) {
                lo2 = 0;
#pragma unroll
                for (int step = next_pow2(NPAD + 1) / 2; step >= 1; step >>= 1) {
                    const int mid = lo2 + step;
                    if (mid <= NPAD && sb[mid - 1] <= v) lo2 = mid;
                }
                lo2 = min(lo2, n);
            }
            if (last) hmax = max(hmax, abs(p + 1 - lo2));
            if (first) hmax = max(hmax, abs(p - lo1));
        }
    }
    return __reduce_max_sync(kFull, hmax);
}

// Histogram of one slice of one class (thread-private).  pos(j) is supplied by the caller.
template <typename T, typename PosFn>
__device__ __forceinline__ void s3p_hist(PosFn pos_of, int n, uint32_t vals_sa, uint32_t hbase, float scale, float off) {
    constexpr float kTop = 4.0f * 32 - 1.5f;
    int j = 0;
    for (; j + 8 <= n; j += 8) {
        uint32_t ad[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const T v = lds_pos(vals_sa, pos_of(j + u), T(0));
            pair_bucket((float)v, scale, off, kTop, hbase, ad[u]);
        }
#pragma unroll
        for (int u = 0; u < 8; u += 4) {
            // four histogram updates in flight: all counters are read before any is written; a later element
            // of the group adds the number of earlier ones that share its counter (stores land in order)
            const uint32_t c0 = lds_u8(ad[u]), c1 = lds_u8(ad[u + 1]), c2 = lds_u8(ad[u + 2]), c3 = lds_u8(ad[u + 3]);
            const uint32_t e1 = (ad[u + 1] == ad[u]) ? 1u : 0u;
            const uint32_t e2 = ((ad[u + 2] == ad[u]) ? 1u : 0u) + ((ad[u + 2] == ad[u + 1]) ? 1u : 0u);
            const uint32_t e3 = ((ad[u + 3] == ad[u]) ? 1u : 0u) + ((ad[u + 3] == ad[u + 1]) ? 1u : 0u) +
                                ((ad[u + 3] == ad[u + 2]) ? 1u : 0u);
            sts_u8(ad[u], c0 + 1u);
            sts_u8(ad[u + 1], c1 + 1u + e1);
            sts_u8(ad[u + 2], c2 + 1u + e2);
            sts_u8(ad[u + 3], c3 + 1u + e3);
        }
    }
    for (; j < n; ++j) {
        uint32_t ad;
        const T v = lds_pos(vals_sa, pos_of(j), T(0));
        pair_bucket((float)v, scale, off, kTop, hbase, ad);
        sts_u8(ad, lds_u8(ad) + 1u);
    }
}

// The same histogram with the positions read from a permutation table (stride `st` entries between this lane's
// elements): the table entries of trip t + 1 are in flight while trip t is gathered and counted -- the entries come
// from HBM / L2, and without the second buffer every trip of 8 pays that latency in full.
template <typename T>
__device__ __forceinline__ void s3p_hist_table(const uint32_t *__restrict__ pp, int st, int n, uint32_t vals_sa, uint32_t hbase,
                                               float scale, float off) {
    constexpr float kTop = 4.0f * 32 - 1.5f;
    uint32_t cur[8], nxt[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) cur[u] = (u < n) ? __ldg(pp + u * st) : 0u;
    int j = 0;
    for (; j + 8 <= n; j += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) nxt[u] = (j + 8 + u < n) ? __ldg(pp + (j + 8 + u) * st) : 0u;
        uint32_t ad[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const T v = lds_pos(vals_sa, cur[u], T(0));
            pair_bucket((float)v, scale, off, kTop, hbase, ad[u]);
        }
#pragma unroll
        for (int u = 0; u < 8; u += 4) {
            const uint32_t c0 = lds_u8(ad[u]), c1 = lds_u8(ad[u + 1]), c2 = lds_u8(ad[u + 2]), c3 = lds_u8(ad[u + 3]);
            const uint32_t e1 = (ad[u + 1] == ad[u]) ? 1u : 0u;
            const uint32_t e2 = ((ad[u + 2] == ad[u]) ? 1u : 0u) + ((ad[u + 2] == ad[u + 1]) ? 1u : 0u);
            const uint32_t e3 = ((ad[u + 3] == ad[u]) ? 1u : 0u) + ((ad[u + 3] == ad[u + 1]) ? 1u : 0u) +
                                ((ad[u + 3] == ad[u + 2]) ? 1u : 0u);
            sts_u8(ad[u], c0 + 1u);
            sts_u8(ad[u + 1], c1 + 1u + e1);
            sts_u8(ad[u + 2], c2 + 1u + e2);
            sts_u8(ad[u + 3], c3 + 1u + e3);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {   // the last, partial trip (its entries are already here)
        if (j + u < n) {
            uint32_t ad;
            const T v = lds_pos(vals_sa, cur[u], T(0));
            pair_bucket((float)v, scale, off, kTop, hbase, ad);
            sts_u8(ad, lds_u8(ad) + 1u);
        }
    }
}

// P threads (lanes of one warp) share a slice in phases 1 and 2; each takes every P-th element.
template <typename T, int E, int P>
__global__ void __launch_bounds__(kS3pThreads, 1) k_step3p(const Step3PairArgs pa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Step3Args &a = pa.a;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nw = blockDim.x >> 5;
    const int grp = lane >> 4;
    const int hl = lane & 15;
    const int m = (int)a.m;
    const int n_slices = (int)a.n_slices;
    const int n_last = (int)a.n_last;
    constexpr int NPAD = 16 * E;

    // shared memory: [class vector | mbarrier | control words | flagged slice list | per-lane histograms (4 KB per
    //                 warp) | merged control histograms (132 bytes per slice) | per-warp scratch for two sorted samples]
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + pa.vals_bytes);
    int *ctl = reinterpret_cast<int *>(bar + 2);                  // [0] n_flagged, [1] L*, [2..3] p_min bits
    unsigned long long *pmin_bits = reinterpret_cast<unsigned long long *>(ctl + 2);
    int *red = ctl + 4;                                           // [nw] per-warp maxima
    uint16_t *flagged = reinterpret_cast<uint16_t *>(red + 32);
    unsigned char *hist_raw = reinterpret_cast<unsigned char *>(flagged + pa.ns_pad);
    hist_raw = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(hist_raw) + 127) & ~(uintptr_t)127);
    // per-lane histogram (words interleaved by lane inside the warp: bank == lane), then the merged control
    // histograms of all slices (pitch 132 bytes: consecutive slices start in consecutive banks)
    constexpr int SL = 32 / P;  // slices per warp in phases 1 and 2
    constexpr uint32_t kMergedPitch = kS3pHistBytes + 4;
    const uint32_t hwarp = smem_u32(hist_raw) + (uint32_t)warp * (32u * kS3pHistBytes);
    const uint32_t hmine = hwarp + (uint32_t)lane * 4u;
    const int nwh = (n_slices + SL - 1) / SL;  // warps that build histograms
    const uint32_t hmerged = smem_u32(hist_raw) + (uint32_t)nwh * (32u * kS3pHistBytes);
    T *scratch = reinterpret_cast<T *>(hist_raw + (size_t)nwh * 32 * kS3pHistBytes + (size_t)pa.ns_pad * kMergedPitch);
    T *sa0 = scratch + (size_t)warp * (2 * NPAD);  // sorted control sample of a slice, then its case sample
    const int my_slice = warp * SL + lane / P;  // phases 1 and 2
    const int my_part = lane % P;
    const bool hist_lane = lane < SL * P && my_slice < n_slices;

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp < nwh) {
#pragma unroll
        for (int i = 0; i < 32; ++i) sts_u32(hmine + 128u * (uint32_t)i, 0u);
    }
    __syncthreads();

    unsigned phase = 0;
    const HalfSortCoef<T> coef(hl);
    const size_t b0c[2] = {0, ((size_t)a.n0 * sizeof(T)) & ~(size_t)15};
    const size_t b1c[2] = {(((size_t)a.n0 * sizeof(T)) + 15) & ~(size_t)15, (((size_t)(a.n0 + a.n1) * sizeof(T)) + 15) & ~(size_t)15};
    const uint32_t vsa[2] = {smem_u32(smem_raw), smem_u32(smem_raw) + (uint32_t)((size_t)a.n0 * sizeof(T) - b0c[1])};

    for (int64_t g = a.g_begin + blockIdx.x; g < a.g_end; g += gridDim.x) {
        const unsigned char *row = static_cast<const unsigned char *>(a.xg) + (size_t)(g - a.xg_g0) * (size_t)a.n_pad * sizeof(T);
        const float2 wn = __ldg(pa.win + g);
        const uint32_t *permc[2] = {nullptr, nullptr};
        AffinePerm ap[2] = {{1u, 0u, 1u}, {1u, 0u, 1u}};
        if (a.perm_mode == PHET_PERM_DEVICE) {
            ap[0] = make_affine(a.seed, (uint64_t)g, 0u, (uint32_t)a.n0, a.coprime0);
            ap[1] = make_affine(a.seed, (uint64_t)g, 1u, (uint32_t)a.n1, a.coprime1);
        } else {
            permc[0] = a.perm0 + (g - a.perm_g0) * a.min_c;
            permc[1] = a.perm1 + (g - a.perm_g0) * a.min_c;
        }
        auto stage = [&](int cls) {
            __syncthreads();  // every read of the vector in residence is done
            if (tid == 0) {
                stage_row(smem_raw, row + b0c[cls], b1c[cls] - b0c[cls], bar);
                // the vector staged next (this gene's case cells, or the control cells of the CTA's next gene)
                // starts its trip from HBM to L2 now
                if (cls == 0) prefetch_row_l2(row + b0c[1], b1c[1] - b0c[1]);
                else if (g + gridDim.x < a.g_end)
                    prefetch_row_l2(row + (size_t)gridDim.x * (size_t)a.n_pad * sizeof(T), b1c[0] - b0c[0]);
                if (a.perm_mode != PHET_PERM_DEVICE) {
                    // ... and so do the table rows read next: the case row of this gene, the control row of the next
                    const size_t row_b = (size_t)a.min_c * sizeof(uint32_t);
                    const size_t pbytes = row_b > 32 ? ((row_b - 16) & ~(size_t)15) : 0;   // (start rounded up: stay inside the row)
                    if (cls == 0) {
                        prefetch_row_l2(reinterpret_cast<const void *>((reinterpret_cast<uintptr_t>(permc[1]) + 15) & ~(uintptr_t)15), pbytes);
                    } else if (g + gridDim.x < a.g_end) {
                        const uint32_t *nx = permc[0] + (size_t)gridDim.x * (size_t)a.min_c;
                        prefetch_row_l2(reinterpret_cast<const void *>((reinterpret_cast<uintptr_t>(nx) + 15) & ~(uintptr_t)15), pbytes);
                    }
                }
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
        };
        // gather + sort (ascending, blocked layout, +BIG pads) of slice k of class cls by this half-warp
        // (control values are gathered from global memory -- the row was just staged, so it sits in L2 -- because
        // the case vector is the one in residence when the flagged slices are known)
        const T *gene0 = reinterpret_cast<const T *>(row);
        auto sort_slice = [&](int cls, int k, T *dst, bool valid) {  // whole warp must call (full-mask shuffles)
            const int n = (k == n_slices - 1) ? n_last : m;
            const uint32_t ncls = (uint32_t)(cls ? a.n1 : a.n0);
            T x[E];
            if (a.perm_mode == PHET_PERM_DEVICE) {
                uint32_t xr = affine_at(ap[cls], (uint64_t)k * (uint64_t)m + (uint64_t)hl);
                const uint32_t step_r = (uint32_t)((16ull * ap[cls].a) % ncls);
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    x[r] = (hl + 16 * r < n) ? (cls ? lds_pos(vsa[1], xr, T(0)) : __ldg(gene0 + xr)) : Big<T>::v();
                    xr = add_mod(xr, step_r, ncls);
                }
            } else {
                const uint32_t *pp = permc[cls] + (size_t)k * (size_t)m;
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    const uint32_t q = (hl + 16 * r < n) ? __ldg(pp + hl + 16 * r) : 0u;
                    x[r] = (hl + 16 * r < n) ? (cls ? lds_pos(vsa[1], q, T(0)) : __ldg(gene0 + q)) : Big<T>::v();
                }
            }
            halfwarp_sort_asc<T, E>(x, coef);
            if (valid) {
#pragma unroll
                for (int r = 0; r < E; ++r) dst[hl * E + r] = x[r];
            }
        };

        if (tid == 0) {
            ctl[0] = 0;
            *pmin_bits = 0x7FF0000000000000ull;  // +inf
        }
        long long t_mark = (pa.prof != nullptr && tid == 0) ? clock64() : 0;
        auto mark = [&](int slot) {  // thread 0: cycles since the previous mark -> prof[slot]
            if (pa.prof != nullptr && tid == 0) {
                const long long t = clock64();
                atomicAdd(pa.prof + slot, (unsigned long long)(t - t_mark));
                t_mark = t;
            }
        };
        // ---- phases 1 and 2: histograms of the control and the case slices, P lanes per slice
        for (int cls = 0; cls < 2; ++cls) {
            stage(cls);
            mark(cls ? 2 : 0);
            if (hist_lane) {
                const int k = my_slice;
                const int n = (k == n_slices - 1) ? n_last : m;
                const int npart = (n - my_part + P - 1) / P;  // elements my_part, my_part + P, ...
                if (a.perm_mode == PHET_PERM_DEVICE) {
                    const uint32_t ncls = (uint32_t)(cls ? a.n1 : a.n0);
                    const uint32_t stepj = (uint32_t)(((uint64_t)P * ap[cls].a) % ncls);
                    uint32_t xr = affine_at(ap[cls], (uint64_t)k * (uint64_t)m + (uint64_t)my_part);
                    s3p_hist<T>([&](int) { const uint32_t q = xr; xr = add_mod(xr, stepj, ncls); return q; }, npart, vsa[cls], hmine, wn.x, wn.y);
                } else {
                    const uint32_t *pp = permc[cls] + (size_t)k * (size_t)m + my_part;
                    s3p_hist_table<T>(pp, P, npart, vsa[cls], hmine, wn.x, wn.y);
                }
            }
            if (cls == 0) {
                // control counts of a slice: sum of its P lanes' histograms (each lane sums every P-th word),
                // parked in the merged area so that the lanes' own histograms can take the case counts
                __syncwarp();
                if (hist_lane) {
                    const uint32_t h0l = hmine - 4u * (uint32_t)my_part;  // lane of part 0 of my slice
                    for (int i = my_part; i < 32; i += P) {
                        uint32_t w = 0u;
#pragma unroll
                        for (int q = 0; q < P; ++q)  // (lane p starts with region p: the P lanes of a slice hit P banks)
                            w += lds_u32(h0l + 128u * (uint32_t)i + 4u * (uint32_t)((q + my_part) % P));  // bytes <= 255 in total
                        sts_u32(hmerged + (uint32_t)my_slice * kMergedPitch + 4u * (uint32_t)i, w);
                    }
                }
                __syncwarp();
                if (warp < nwh) {  // the warp's 4 KB of histograms are contiguous: each lane clears 128 of them
#pragma unroll
                    for (int i = 0; i < 8; ++i) sts_v4(hwarp + 512u * (uint32_t)i + 16u * (uint32_t)lane, 0u, 0u, 0u, 0u);
                }
                __syncwarp();
                mark(1);
            }
        }
        mark(3);  // (phase 2 of thread 0's own warp)
        // ---- bounds on h from the two histograms, by the P lanes of the slice: lane p takes the words
        //      [w0, w1) of the 32, after an exclusive prefix of the per-lane totals of (ha - hb)
        __syncwarp();
        int L = 0, U = 0;
        {
            const int w0 = (32 * my_part) / P, w1 = (32 * (my_part + 1)) / P;
            const uint32_t h0l = hmine - 4u * (uint32_t)my_part;
            const uint32_t hm = hmerged + (uint32_t)(hist_lane ? my_slice : 0) * kMergedPitch;
            int dsum = 0;
            if (hist_lane) {
                for (int i = w0; i < w1; ++i) {
                    const uint32_t wa = lds_u32(hm + 4u * (uint32_t)i);
                    uint32_t wb = 0u;
#pragma unroll
                    for (int q = 0; q < P; ++q) wb += lds_u32(h0l + 128u * (uint32_t)i + 4u * (uint32_t)((q + my_part) % P));
                    dsum += (int)((wa * 0x01010101u) >> 24) - (int)((wb * 0x01010101u) >> 24);
                }
            }
            int D = 0;  // ECDF difference at the lower edge of my first bucket
#pragma unroll
            for (int q = 1; q < P; ++q) {
                const int dq = __shfl_up_sync(kFull, dsum, q);
                if (my_part >= q) D += dq;
            }
            if (hist_lane) {
                for (int i = w0; i < w1; ++i) {
                    const uint32_t wa = lds_u32(hm + 4u * (uint32_t)i);
                    uint32_t wb = 0u;
#pragma unroll
                    for (int q = 0; q < P; ++q) wb += lds_u32(h0l + 128u * (uint32_t)i + 4u * (uint32_t)((q + my_part) % P));
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int ca = (int)((wa >> (8 * q)) & 255u), cb = (int)((wb >> (8 * q)) & 255u);
                        U = max(U, max(D + ca, cb - D));
                        D += ca - cb;
                        L = max(L, abs(D));
                    }
                }
            }
            // max over the P lanes of the slice (lanes part 0 .. P-1 are adjacent)
            int Lm = L, Um = U;
#pragma unroll
            for (int q = 1; q < P; ++q) {
                const int ld = __shfl_down_sync(kFull, L, q), ud = __shfl_down_sync(kFull, U, q);
                if (my_part == 0) {
                    Lm = max(Lm, ld);
                    Um = max(Um, ud);
                }
            }
            L = Lm;
            U = Um;
            __syncwarp();
            if (warp < nwh) {  // clear the case counts for the next gene
#pragma unroll
                for (int i = 0; i < 8; ++i) sts_v4(hwarp + 512u * (uint32_t)i + 16u * (uint32_t)lane, 0u, 0u, 0u, 0u);
            }
            __syncwarp();
        }
        mark(4);
        const bool leader = hist_lane && my_part == 0;  // holds the slice's L and U
        const bool full_len = leader && !(my_slice == n_slices - 1 && n_last != m);
        int lmax = __reduce_max_sync(kFull, full_len ? L : 0);
        if (lane == 0) red[warp] = lmax;
        __syncthreads();
        int Lstar = 0;
        for (int w = 0; w < nw; ++w) Lstar = max(Lstar, red[w]);
        if (leader) {
            const bool need = !full_len || U > Lstar || Lstar < pa.h0_full;
            if (need) flagged[atomicAdd(&ctl[0], 1)] = (uint16_t)my_slice;
        }
        __syncthreads();
        const int nflag = ctl[0];
        mark(5);
        // ---- exact h of the flagged slices: one per warp trip; the half-warp sorter sorts the control sample
        //      (lanes 0-15, values gathered from L2) and the case sample (lanes 16-31, from SMEM) at once
        for (int i = warp; i < nflag && warp < pa.nwx; i += pa.nwx) {
            const int k = (int)flagged[i];
            const int n = (k == n_slices - 1) ? n_last : m;
            sort_slice(grp, k, sa0 + (size_t)grp * NPAD, true);
            __syncwarp();
            const int h = ks_h_sorted<T, NPAD>(sa0, sa0 + NPAD, n, lane);
            if (lane == 0) {
                const double pv = (n == m) ? __ldg(a.tab_full + h) : __ldg(a.tab_last + h);
                atomicMin(pmin_bits, (unsigned long long)__double_as_longlong(pv));
            }
            __syncwarp();
        }
        mark(6);
        __syncthreads();
        mark(7);
        if (tid == 0) {
            double o = __longlong_as_double((long long)*pmin_bits);
            if (Lstar >= pa.h0_full) o = fmin(o, __ldg(a.tab_full + Lstar));  // some full slice reaches at least L*
            a.o[g] = o;
        }
    }
}

// Common bucket map of a gene for Step 3: one warp sorts 256 + 256 cells of the two classes and
// maps the central 99 % of the pooled sample onto buckets 1 .. interior.
template <typename T>
__global__ void __launch_bounds__(256) k_window3(const void *__restrict__ xg, int64_t n_pad, int64_t xg_g0, int64_t g_begin,
                                                 int64_t g_end, int64_t n0, int64_t n1, float interior,
                                                 float2 *__restrict__ win) {
    const int lane = threadIdx.x & 31;
    const int64_t g = g_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= g_end) return;
    const int h0 = (int)(n0 < 256 ? n0 : 256), h1 = (int)(n1 < 256 ? n1 : 256);
    const int ns = h0 + h1;
    const T *gene = static_cast<const T *>(xg) + (g - xg_g0) * n_pad;
    T v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const int k = lane + 32 * r;
        v[r] = (k < h0) ? __ldg(gene + k) : ((k < ns) ? __ldg(gene + n0 + (k - h0)) : Lim<T>::inf());
    }
    warp_sort_asc<T, 16>(v, lane);
    const int cut = ns / 200;
    float lo = (float)warp_pick<T, 16>(v, cut);
    float hi = (float)warp_pick<T, 16>(v, ns - 1 - cut);
    if (!(hi > lo)) {
        lo = (float)warp_pick<T, 16>(v, 0);
        hi = (float)warp_pick<T, 16>(v, ns - 1);
    }
    float scale = 0.0f, off = 1.0f;
    if (hi > lo) {
        scale = interior / (hi - lo);
        off = 1.0f - lo * scale;
        if (!isfinite(scale) || !isfinite(off)) {
            scale = 0.0f;
            off = 1.0f;
        }
    }
    if (lane == 0) win[g] = make_float2(scale, off);
}

struct S3pGeometry {
    bool ok = false;
    int vals_bytes = 0, ns_pad = 0, P = 1, nwx = 0;
    size_t smem = 0;
};

// Largest (lanes per slice, exact-phase warps) combination whose shared memory fits.
template <typename T, int E>
static S3pGeometry s3p_geometry(const Step3Args &a) {
    S3pGeometry ge;
    const int64_t nmax = a.n0 > a.n1 ? a.n0 : a.n1;
    if (a.m < 1 || a.m > 255 || a.n_slices < 1) return ge;
    ge.vals_bytes = (int)((nmax * (int64_t)sizeof(T) + 32 + 127) / 128 * 128);
    ge.ns_pad = (int)((a.n_slices + 31) / 32 * 32);
    const size_t npad_b = (size_t)16 * E * sizeof(T);
    for (int P = 3; P >= 1; P -= 2) {
        const int SL = 32 / P;
        if (a.n_slices > (kS3pThreads / 32) * SL) continue;
        const size_t nwh = (size_t)((a.n_slices + SL - 1) / SL);
        for (int nwx = kS3pThreads / 32; nwx >= 2; nwx >>= 1) {
            const size_t smem = (size_t)ge.vals_bytes + 16 + 16 + 32 * sizeof(int) + (size_t)ge.ns_pad * 2 + 128 +
                                nwh * 32 * kS3pHistBytes + (size_t)ge.ns_pad * (kS3pHistBytes + 4) + (size_t)nwx * 2 * npad_b;
            if (smem <= (size_t)kMaxSmemOptin) {
                ge.P = P;
                ge.nwx = nwx;
                ge.smem = smem;
                ge.ok = true;
                return ge;
            }
        }
    }
    return ge;
}

// Returns 1 when the path does not apply (caller falls back to the sorting kernels).
template <typename T, int E>
static int launch_step3p_E(phet_ctx *c, const Step3Args &args) {
    // as many lanes per slice as the slice count allows (16 warps x 32/P slices)
    const S3pGeometry ge = s3p_geometry<T, E>(args);
    if (!ge.ok) return 1;
    auto kern = ge.P == 3 ? k_step3p<T, E, 3> : k_step3p<T, E, 1>;
    PHET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ge.smem));
    int per_sm = 0;
    PHET_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kS3pThreads, ge.smem));
    if (per_sm < 1) return 1;
    const int64_t ng = args.g_end - args.g_begin;
    int64_t grid = (int64_t)per_sm * c->num_sms;
    if (grid > ng) grid = ng;
    if (grid < 1) return 0;
    if (int e = c->win3.alloc((size_t)c->G)) return e;
    k_window3<T><<<(unsigned)((ng + 7) / 8), 256, 0, c->stream>>>(args.xg, args.n_pad, args.xg_g0, args.g_begin, args.g_end,
                                                              args.n0, args.n1, 125.0f, c->win3.p);
    c->launches++;
    Step3PairArgs pa;
    pa.a = args;
    pa.a.H = nullptr;
    pa.win = c->win3.p;
    pa.vals_bytes = ge.vals_bytes;
    pa.ns_pad = ge.ns_pad;
    pa.nwx = ge.nwx;
    pa.prof = c->prof.p;
    pa.h0_full = c->tab_h0;
    kern<<<(unsigned)grid, kS3pThreads, ge.smem, c->stream>>>(pa);
    c->launches++;
    c->step3_path = 3;
    PHET_CUDA(cudaGetLastError());
    return 0;
}

// The gather formulation serves this launch (its histograms and the class vector fit in shared memory).
template <typename T>
inline bool step3_prune_applies(const phet_ctx *c, const Step3Args &args) {
    if (c->keep_H || c->step3_impl == 1 || c->step3_impl == 3 || c->tab_h0 < 0 || args.n_last > args.m) return false;
    const int need = (int)((args.m + 15) / 16);
    if (need <= 1) return s3p_geometry<T, 1>(args).ok;
    if (need <= 2) return s3p_geometry<T, 2>(args).ok;
    if (need <= 3) return s3p_geometry<T, 3>(args).ok;
    if (need <= 4) return s3p_geometry<T, 4>(args).ok;
    if (need <= 5) return s3p_geometry<T, 5>(args).ok;
    if (need <= 6) return s3p_geometry<T, 6>(args).ok;
    if (need <= 7) return s3p_geometry<T, 7>(args).ok;
    if (need <= 8) return s3p_geometry<T, 8>(args).ok;
    if (need <= 9) return s3p_geometry<T, 9>(args).ok;
    if (need <= 10) return s3p_geometry<T, 10>(args).ok;
    if (need <= 12) return s3p_geometry<T, 12>(args).ok;
    return s3p_geometry<T, 16>(args).ok;
}

template <typename T>
int run_step3_prune(phet_ctx *c, const Step3Args &args) {
    if (!step3_prune_applies<T>(c, args)) return 1;
    const int need = (int)((args.m + 15) / 16);
    if (need <= 1) return launch_step3p_E<T, 1>(c, args);
    if (need <= 2) return launch_step3p_E<T, 2>(c, args);
    if (need <= 3) return launch_step3p_E<T, 3>(c, args);
    if (need <= 4) return launch_step3p_E<T, 4>(c, args);
    if (need <= 5) return launch_step3p_E<T, 5>(c, args);
    if (need <= 6) return launch_step3p_E<T, 6>(c, args);
    if (need <= 7) return launch_step3p_E<T, 7>(c, args);
    if (need <= 8) return launch_step3p_E<T, 8>(c, args);
    if (need <= 9) return launch_step3p_E<T, 9>(c, args);
    if (need <= 10) return launch_step3p_E<T, 10>(c, args);
    if (need <= 12) return launch_step3p_E<T, 12>(c, args);
    return launch_step3p_E<T, 16>(c, args);
}

}  // namespace phet
