// Fisher-Yates on the device: turns the swap targets drawn by csrc/mt_replay.cpp (the sequential half of
// np.random.permutation) into permutations -- the random-access half, independent per permutation.
//
//   x = arange(n);  for t = 0 .. n-2:  i = n-1-t;  swap(x[i], x[draws[t]])          (numpy _shuffle_raw)
//
// One WARP per permutation, the array in shared memory (16-bit entries, n <= 65536) or, for longer classes, in
// a global scratch row.  The warp takes 32 consecutive steps at a time, lane l the step i = i0 - l.  Steps of a
// batch commute unless an earlier lane's target equals a later lane's target or a later lane's position i, so:
// peers with equal targets come from match.any, targets that fall on a later lane's i from one OR-reduction, and
// the batch is executed as maximal conflict-free runs of lanes (usually the whole batch at once: a conflict has
// probability ~ 1500 / i per batch; the last few hundred steps degrade to short runs).
//
// The first `take` entries are written out through an optional map (class-local index -> whatever the consumer
// addresses: original row numbers for the Step-1 index sets, class-relative grouped positions for Step 3).
#include "common.cuh"

namespace phet {

constexpr unsigned kFullMask = 0xffffffffu;

template <typename X>
struct FyShared {  // x in shared memory
    X *x;
    __device__ __forceinline__ uint32_t get(uint32_t k) const { return (uint32_t)x[k]; }
    __device__ __forceinline__ void put(uint32_t k, uint32_t v) const { x[k] = (X)v; }
};

struct FyGlobal {  // x in a global scratch row (L2): volatile keeps the accesses in program order around __syncwarp
    volatile uint32_t *x;
    __device__ __forceinline__ uint32_t get(uint32_t k) const { return x[k]; }
    __device__ __forceinline__ void put(uint32_t k, uint32_t v) const { x[k] = v; }
};

struct FyInv {        // inverse output: slice of every cell instead of the cell at every permuted position
    uint16_t *out;    // [count][out_ld]; nullptr = forward output
    uint32_t m, last_slice, magic_m;
};

template <typename J, typename Store>
__device__ __forceinline__ void fy_one(const Store &st, const J *__restrict__ draws, uint32_t n, uint32_t take,
                                       const int32_t *__restrict__ map, uint32_t *__restrict__ out, int lane,
                                       const FyInv &inv, uint16_t *__restrict__ inv_out) {
    for (uint32_t q = lane; q < n; q += 32) st.put(q, q);
    __syncwarp();
    const uint32_t steps = n - 1;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t jn = (lane < steps) ? (uint32_t)draws[lane] : 0u;
    for (uint32_t t0 = 0; t0 < steps; t0 += 32) {
        const uint32_t t = t0 + lane;
        const bool active = t < steps;
        const uint32_t i0 = n - 1 - t0;       // lane 0's position
        const uint32_t i = i0 - lane;         // (wraps for inactive lanes; never used then)
        const uint32_t j = active ? jn : 0xFFFFFFFFu - lane;  // inactive lanes: distinct dummies, no peers
        if (t + 32 < steps) jn = (uint32_t)draws[t + 32];     // next batch's targets, in flight during this one
        const unsigned peers = __match_any_sync(kFullMask, j);
        // my target is the position of a LATER lane of this batch (lane i0 - j > mine)
        const uint32_t d = i0 - j;
        const unsigned tgt = (active && j != i && d < 32u) ? (1u << d) : 0u;
        unsigned todo = __ballot_sync(kFullMask, active);
        unsigned hits = __reduce_or_sync(kFullMask, tgt);
        unsigned conf = __ballot_sync(kFullMask, ((peers & lt) != 0u) || ((hits >> lane) & 1u));
        if (conf == 0u) {
            if (active) {
                const uint32_t a = st.get(i), b = st.get(j);
                st.put(i, b);
                st.put(j, a);
            }
            __syncwarp();
            continue;
        }
        // runs: lanes [first remaining, first conflicted remaining lane) commute
        while (todo != 0u) {
            const int first = __ffs(todo) - 1;
            const unsigned conf_rest = conf & todo & ~(1u << first);  // the first remaining lane has nothing before it
            const int stop = conf_rest ? (__ffs(conf_rest) - 1) : 32;
            const unsigned run = todo & ((stop == 32) ? kFullMask : ((1u << stop) - 1u));
            if ((run >> lane) & 1u) {
                const uint32_t a = st.get(i), b = st.get(j);
                st.put(i, b);
                st.put(j, a);
            }
            __syncwarp();
            todo &= ~run;
            if (todo == 0u) break;
            // conflicts among the remaining lanes only
            hits = __reduce_or_sync(kFullMask, ((todo >> lane) & 1u) ? tgt : 0u);
            conf = __ballot_sync(kFullMask, ((todo >> lane) & 1u) && (((peers & lt & todo) != 0u) || ((hits >> lane) & 1u)));
        }
    }
    __syncwarp();
    if (inv_out != nullptr) {   // every cell of the class is written once: its slice, or 0xFFFF beyond the sliced positions
        for (uint32_t q = lane; q < n; q += 32) {
            const uint32_t v = st.get(q);
            const uint32_t cell = map ? (uint32_t)__ldg(map + v) : v;
            inv_out[cell] = q < take ? (uint16_t)min(__umulhi(q, inv.magic_m), inv.last_slice) : (uint16_t)0xFFFFu;
        }
    } else {
        for (uint32_t q = lane; q < take; q += 32) {
            const uint32_t v = st.get(q);
            out[q] = map ? (uint32_t)__ldg(map + v) : v;
        }
    }
    __syncwarp();
}

struct FyArgs {
    const void *draws;     // [count][ld] J
    int64_t ld;            // elements between permutations
    uint32_t n, take;
    int64_t count;
    const int32_t *map;    // [n] or NULL
    uint32_t *out;         // [count][out_ld]
    int64_t out_ld;
    uint32_t *scratch;     // global variant: [grid warps][n]
    uint32_t x_bytes;      // shared variant: bytes per warp
    FyInv inv;
};

template <typename J>
__global__ void __launch_bounds__(512) k_fy_apply_smem(const FyArgs a) {
    extern __shared__ __align__(16) unsigned char fy_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    FyShared<uint16_t> st{reinterpret_cast<uint16_t *>(fy_smem + (size_t)warp * a.x_bytes)};
    for (int64_t p = (int64_t)blockIdx.x * nw + warp; p < a.count; p += (int64_t)gridDim.x * nw)
        fy_one<J>(st, static_cast<const J *>(a.draws) + p * a.ld, a.n, a.take, a.map, a.out ? a.out + p * a.out_ld : nullptr, lane, a.inv,
                  a.inv.out ? a.inv.out + p * a.out_ld : nullptr);
}

template <typename J>
__global__ void __launch_bounds__(256) k_fy_apply_gmem(const FyArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    FyGlobal st{a.scratch + ((size_t)blockIdx.x * nw + warp) * a.n};
    for (int64_t p = (int64_t)blockIdx.x * nw + warp; p < a.count; p += (int64_t)gridDim.x * nw)
        fy_one<J>(st, static_cast<const J *>(a.draws) + p * a.ld, a.n, a.take, a.map, a.out ? a.out + p * a.out_ld : nullptr, lane, a.inv,
                  a.inv.out ? a.inv.out + p * a.out_ld : nullptr);
}

// draws_dev: device [count][ld] of 2- or 4-byte swap targets; out_dev: device uint32 [count][out_ld].
static int launch_fy(phet_ctx *c, const void *draws_dev, int elem_bytes, int64_t ld, int64_t n, int64_t count, int64_t take,
                     const int32_t *map_dev, uint32_t *out_dev, int64_t out_ld, const FyInv &inv, cudaStream_t stream) {
    if (count <= 0 || take <= 0) return PHET_OK;
    PHET_REQUIRE(n >= 1 && n < ((int64_t)1 << 31) && take <= n, "fy_apply: bad permutation length");
    PHET_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "fy_apply: elem_bytes must be 2 or 4");
    FyArgs a;
    a.draws = draws_dev;
    a.ld = ld;
    a.n = (uint32_t)n;
    a.take = (uint32_t)take;
    a.count = count;
    a.map = map_dev;
    a.out = out_dev;
    a.out_ld = out_ld;
    a.scratch = nullptr;
    a.x_bytes = 0;
    a.inv = inv;
    if (n <= 65536) {
        a.x_bytes = (uint32_t)((n * 2 + 15) / 16 * 16);
        int warps = (int)(kMaxSmemOptin / a.x_bytes);
        if (warps > 16) warps = 16;
        int ctas_per_sm = 1;
        if (warps == 16) {  // short permutations: several CTAs per SM
            ctas_per_sm = (int)(kMaxSmemOptin / ((size_t)16 * a.x_bytes + 1024));
            if (ctas_per_sm > 4) ctas_per_sm = 4;
            if (ctas_per_sm < 1) ctas_per_sm = 1;
        }
        const size_t smem = (size_t)warps * a.x_bytes;
        auto kern = elem_bytes == 2 ? k_fy_apply_smem<uint16_t> : k_fy_apply_smem<uint32_t>;
        PHET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t grid = (count + warps - 1) / warps;
        const int64_t cap = (int64_t)c->num_sms * ctas_per_sm;
        if (grid > cap) grid = cap;
        kern<<<(unsigned)grid, warps * 32, smem, stream>>>(a);
    } else {
        const int warps = 8;
        int64_t grid = (count + warps - 1) / warps;
        const int64_t cap = (int64_t)c->num_sms * 4;
        if (grid > cap) grid = cap;
        if (int e = c->fy_scratch.alloc((size_t)grid * warps * (size_t)n)) return e;
        a.scratch = c->fy_scratch.p;
        auto kern = elem_bytes == 2 ? k_fy_apply_gmem<uint16_t> : k_fy_apply_gmem<uint32_t>;
        kern<<<(unsigned)grid, warps * 32, 0, stream>>>(a);
    }
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    return PHET_OK;
}

int launch_fy_apply(phet_ctx *c, const void *draws_dev, int elem_bytes, int64_t ld, int64_t n, int64_t count, int64_t take,
                    const int32_t *map_dev, uint32_t *out_dev, int64_t out_ld, cudaStream_t stream) {
    FyInv inv{nullptr, 1u, 0u, 0u};
    return launch_fy(c, draws_dev, elem_bytes, ld, n, count, take, map_dev, out_dev, out_ld, inv, stream);
}

int launch_fy_apply_inverse(phet_ctx *c, const void *draws_dev, int elem_bytes, int64_t ld, int64_t n, int64_t count, int64_t take,
                            const int32_t *map_dev, uint16_t *slice_out, int64_t out_ld, int64_t m, int64_t last_slice,
                            cudaStream_t stream) {
    PHET_REQUIRE(m >= 1 && last_slice >= 0 && last_slice < 65535 && (uint64_t)take * (uint64_t)m < ((uint64_t)1 << 32),
                 "fy_apply: slice geometry outside the 16-bit inverse table");
    FyInv inv{slice_out, (uint32_t)m, (uint32_t)last_slice, (uint32_t)(((uint64_t)1 << 32) / (uint64_t)m) + 1u};
    return launch_fy(c, draws_dev, elem_bytes, ld, n, count, take, map_dev, nullptr, out_ld, inv, stream);
}

}  // namespace phet
