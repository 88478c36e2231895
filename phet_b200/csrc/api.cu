// C ABI of libphet_b200 (see include/phet_b200.h).  Context management, uploads, and the thin
// launch wrappers; the kernels live in normalize.cu, step1_*.cu, step3_*.cu and finalize.cu.
#include "common.cuh"
#include "pvalue.cuh"
#include "step1_impl.cuh"
#include "step1p_impl.cuh"
#include "step3_impl.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace phet {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %s (%s) at %s:%d in `%s`", cudaGetErrorName(e), cudaGetErrorString(e), file,
             line, what);
    g_last_error = buf;
    return e == cudaErrorMemoryAllocation ? PHET_ERR_NOMEM : PHET_ERR_CUDA;
}

int run_step1_f32(phet_ctx *c, const Step1Args &a);
int run_step1_f64(phet_ctx *c, const Step1Args &a);
int run_step3_f32(phet_ctx *c, const Step3Args &a);
int run_step3_f64(phet_ctx *c, const Step3Args &a);
int launch_ks_big(phet_ctx *c, const Step3Args &a, int fold);

// original row numbers -> class-grouped positions, in place
__global__ void k_remap_idx(int32_t *__restrict__ idx, int64_t n, const int32_t *__restrict__ pos_of_row, int64_t N,
                            int *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t r = idx[i];
    if (r < 0 || r >= N) {
        *bad = 1;
        idx[i] = 0;
    } else {
        idx[i] = pos_of_row[r];
    }
}

// host-mode permutations: class-local index (ascending row order) -> class-relative grouped position
__global__ void k_remap_perm(uint32_t *__restrict__ perm, int64_t n, const int32_t *__restrict__ cls_pos, uint32_t ncls,
                             int *__restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = perm[i];
    if (t >= ncls) {
        *bad = 1;
        perm[i] = 0;
    } else {
        perm[i] = (uint32_t)cls_pos[t];
    }
}

// [S][2][m] grouped positions -> [2][nblk][mp][32] class-local positions, 16 bits each (mp = m rounded
// up to 16, rows >= m are 0): the 32 subsamples of a block are adjacent, so a warp of k_step1p reads 64
// contiguous bytes per element, and whole 16-element chunks can be loaded without bounds checks.
//
// The ORDER of the elements inside a subsample is free (every statistic of Step 1 is a function of the
// set), so it is chosen to spread the warp's shared-memory gathers over the banks.  One WARP lays out the 32
// subsamples of a block together, row by row: every lane proposes, among the residues (position mod 32 = bank
// of a float) it still holds and no other lane has taken in this row, the one it holds most of; equal
// proposals are decided in favour of one lane and the losers propose again; a lane left without a free residue
// doubles one up.  Simulated on random subsamples of 158 of 25,000 cells: 1.40 wavefronts per gather, against 2.5
// for independent per-lane orders (lane l preferring residue l + r) and 3.6 for random order.  m <= 255.
__global__ void __launch_bounds__(128) k_build_idxT(const int32_t *__restrict__ idx, int64_t S, int64_t m, int64_t mp, int64_t n0,
                                                    int64_t nblk, uint16_t *__restrict__ idxT) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // over [2][nblk]
    if (wid >= 2 * nblk) return;
    const int64_t c = wid / nblk, blk = wid % nblk, s = blk * 32 + lane;
    uint16_t *out = idxT + ((c * nblk + blk) * mp) * 32 + lane;  // row r at out[32 * r]
    const bool active = s < S;
    uint16_t sorted[256];   // this lane's positions grouped by residue mod 32
    unsigned char start[33], left[32];
    uint32_t M[9];          // M[k]: residues this lane holds at least k elements of (k = 1..8)
    for (int k = 0; k < 9; ++k) M[k] = 0u;
    if (active) {
        const int32_t *src = idx + (s * 2 + c) * m;
        const int32_t off = c ? (int32_t)n0 : 0;
        int cnt[32];
        for (int b = 0; b < 32; ++b) cnt[b] = 0;
        for (int64_t j = 0; j < m; ++j) cnt[(src[j] - off) & 31]++;
        int acc = 0;
        for (int b = 0; b < 32; ++b) {
            start[b] = (unsigned char)acc;
            left[b] = (unsigned char)cnt[b];
            for (int k = 1; k <= 8; ++k)
                if (cnt[b] >= k) M[k] |= 1u << b;
            acc += cnt[b];
            cnt[b] = 0;
        }
        start[32] = (unsigned char)acc;
        for (int64_t j = 0; j < m; ++j) {
            const int p = src[j] - off, b = p & 31;
            sorted[start[b] + cnt[b]++] = (uint16_t)p;
        }
    }
    // best residue of `set` for this lane: the one it holds most of; ties broken by a rotation that differs per lane and row
    auto best_of = [&](uint32_t set, int rot) -> int {
        uint32_t t = 0u;
        for (int k = 8; k >= 1; --k) {
            t = M[k] & set;
            if (t) break;
        }
        const uint32_t tr = __funnelshift_r(t, t, rot & 31);   // bit b -> bit (b - rot) mod 32
        return ((__ffs((int)tr) - 1) + rot) & 31;
    };
    for (int64_t r = 0; r < m; ++r) {
        const int rr = (int)(r & 31);
        uint32_t taken = 0u;
        bool undecided = active;
        int choice = 0;
        for (int round = 0; round < 32; ++round) {
            const uint32_t cand = M[1] & ~taken;
            const bool bids = undecided && cand != 0u;
            if (!__any_sync(0xffffffffu, bids)) break;
            const int want = bids ? best_of(cand, lane + rr) : 32 + lane;   // non-bidders: distinct dummies, no peers
            const uint32_t peers = __match_any_sync(0xffffffffu, want);
            // among equal proposals the lane first in the order lane rr, rr + 1, ... wins this row
            const uint32_t pr = __funnelshift_r(peers, peers, rr);
            const bool wins = bids && (((__ffs((int)pr) - 1) + rr) & 31) == lane;
            if (wins) {
                choice = want;
                undecided = false;
            }
            taken |= __reduce_or_sync(0xffffffffu, wins ? (1u << want) : 0u);
        }
        if (active) {
            if (undecided) choice = best_of(M[1], lane + rr);   // every residue this lane holds is taken: double one up
            const int old = left[choice];
            left[choice] = (unsigned char)(old - 1);
            if (old <= 8) M[old] &= ~(1u << choice);
            out[32 * r] = sorted[start[choice] + old - 1];
        } else {
            out[32 * r] = 0;
        }
    }
    for (int64_t r = m; r < mp; ++r) out[32 * r] = 0;
}

// BIG variant of the table: [2][nblk][mp][32] class-local positions, 32 bits each, every subsample's positions in
// ASCENDING order (the order is free, see above): the warps of a CTA then walk the class vector front to back
// together and share the lines they pull into L1.  One thread per (class, subsample); m <= 496.
__global__ void k_build_idxT32(const int32_t *__restrict__ idx, int64_t S, int64_t m, int64_t mp, int64_t n0, int64_t nblk,
                               uint32_t *__restrict__ idxT) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over [2][nblk * 32]
    if (t >= 2 * nblk * 32) return;
    const int64_t c = t / (nblk * 32), s = t % (nblk * 32);
    uint32_t *out = idxT + ((c * nblk + (s >> 5)) * mp) * 32 + (s & 31);  // row r at out[32 * r]
    if (s >= S) {
        for (int64_t r = 0; r < mp; ++r) out[32 * r] = 0;
        return;
    }
    const int32_t *src = idx + (s * 2 + c) * m;
    const int32_t off = c ? (int32_t)n0 : 0;
    // insertion sort straight into the (strided) output column
    for (int64_t j = 0; j < m; ++j) {
        const uint32_t p = (uint32_t)(src[j] - off);
        int64_t r = j;
        while (r > 0 && out[32 * (r - 1)] > p) {
            out[32 * r] = out[32 * (r - 1)];
            --r;
        }
        out[32 * r] = p;
    }
    for (int64_t r = m; r < mp; ++r) out[32 * r] = 0;
}

static uint64_t gcd_u64(uint64_t a, uint64_t b) {
    while (b) {
        const uint64_t t = a % b;
        a = b;
        b = t;
    }
    return a;
}

// kCoprimeCount multipliers in [1, n) coprime to n (n == 1: all ones).  Mirrored in phet_b200/perm.py.
static void coprime_list(uint32_t n, uint32_t *out) {
    uint64_t j = 0;
    for (int k = 0; k < kCoprimeCount;) {
        const uint64_t z = splitmix64(0xC0FFEE123456789ull ^ ((uint64_t)n * 0x9E3779B97F4A7C15ull) ^ (j++));
        const uint32_t cand = n > 2 ? (uint32_t)(1 + z % (uint64_t)(n - 1)) : 1u;
        if (gcd_u64(cand, n) == 1) out[k++] = cand;
    }
}

// a^-1 mod n for gcd(a, n) = 1 (n == 1: 0)
static uint32_t inverse_mod(uint32_t a, uint32_t n) {
    int64_t t = 0, nt = 1, r = (int64_t)n, nr = (int64_t)(a % n);
    while (nr != 0) {
        const int64_t qq = r / nr;
        const int64_t t2 = t - qq * nt;
        t = nt;
        nt = t2;
        const int64_t r2 = r - qq * nr;
        r = nr;
        nr = r2;
    }
    if (t < 0) t += (int64_t)n;
    return (uint32_t)(t % (int64_t)n);
}

bool step3_scatter_applies_f32(const phet_ctx *c, const Step3Args &a);   // step3s_f32.cu / step3s_f64.cu
bool step3_scatter_applies_f64(const phet_ctx *c, const Step3Args &a);

// The scatter kernel (step3s_impl.cuh) will serve Step 3 of a fit with this geometry: table modes then provide the
// inverse view of the permutations (slice of every cell) instead of, or next to, the forward one.
bool step3_wants_inverse(phet_ctx *c, int64_t m, int64_t n_last) {
    Step3Args a{};
    a.n0 = c->n0;
    a.n1 = c->n1;
    a.m = m;
    a.n_slices = c->n_slices;
    a.n_last = n_last;
    a.min_c = (c->n_slices - 1) * m + n_last;
    a.cap = m > n_last ? m : n_last;
    a.perm_mode = PHET_PERM_HOST;
    return c->storage == PHET_F32 ? step3_scatter_applies_f32(c, a) : step3_scatter_applies_f64(c, a);
}

// Inverse view of uploaded permutation tables: slice_of[g][perm[g][t]] = min(t / m, last slice).
__global__ void k_slice_of_from_perm(const uint32_t *__restrict__ perm, int64_t n, int64_t min_c, uint16_t *__restrict__ slice_of,
                                     int64_t ld, uint32_t magic_m, uint32_t last_slice) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t g = i / min_c;
    const uint32_t t = (uint32_t)(i - g * min_c);
    slice_of[g * ld + perm[i]] = (uint16_t)min(__umulhi(t, magic_m), last_slice);
}

int launch_step1(phet_ctx *c, int64_t g_begin, int64_t g_end, int64_t s_begin, int64_t s_end,
                 const phet_step1_params *prm, int accumulate) {
    Step1Args a;
    a.xg = c->xg.p;
    a.n_pad = c->n_pad;
    a.G = c->G;
    a.xg_g0 = c->xg_g0;
    a.g_begin = g_begin;
    a.g_end = g_end;
    a.idx = c->idx.p;
    a.m = c->m;
    a.s_begin = s_begin;
    a.s_end = s_end;
    a.acc_f = c->acc.p;
    a.acc_r = c->acc.p + c->G;
    a.dbgP = nullptr;
    a.dbgD = nullptr;
    a.dbg_ld = s_end - s_begin;
    if (prm->capture_debug) {
        const size_t need = (size_t)c->G * (size_t)(s_end - s_begin);
        const bool fresh = c->dbgP.n < need || c->S_local != s_end - s_begin;
        if (int e = c->dbgP.alloc(need)) return e;
        if (int e = c->dbgD.alloc(need)) return e;
        if (fresh) {
            PHET_CUDA(cudaMemsetAsync(c->dbgP.p, 0, need * sizeof(double), c->stream));
            PHET_CUDA(cudaMemsetAsync(c->dbgD.p, 0, need * sizeof(double), c->stream));
        }
        a.dbgP = c->dbgP.p;
        a.dbgD = c->dbgD.p;
        c->S_local = s_end - s_begin;
    }
    a.q = prm->q;
    a.test = prm->test;
    a.df = 2.0 * (double)c->m - 2.0;
    a.ln_beta = prm->ln_beta;
    a.p_flush = prm->p_flush_below;
    a.accumulate = accumulate;
    if (s_end <= s_begin || g_end <= g_begin) {
        if (!accumulate && g_end > g_begin) {
            PHET_CUDA(cudaMemsetAsync(c->acc.p + g_begin, 0, sizeof(double) * (g_end - g_begin), c->stream));
            PHET_CUDA(cudaMemsetAsync(c->acc.p + c->G + g_begin, 0, sizeof(double) * (g_end - g_begin), c->stream));
        }
        return 0;
    }
    return c->storage == PHET_F32 ? run_step1_f32(c, a) : run_step1_f64(c, a);
}

int launch_step3(phet_ctx *c, int64_t g_begin, int64_t g_end, int64_t m, int perm_mode, uint64_t seed,
                 int64_t n_last) {
    Step3Args a;
    a.xg = c->xg.p;
    a.n_pad = c->n_pad;
    a.G = c->G;
    a.xg_g0 = c->xg_g0;
    a.n0 = c->n0;
    a.n1 = c->n1;
    a.m = m;
    a.n_slices = c->n_slices;
    a.n_last = n_last;
    a.min_c = (c->n_slices - 1) * m + n_last;
    a.cap = m > n_last ? m : n_last;
    a.h_ld = c->n_slices;
    a.g_begin = g_begin;
    a.g_end = g_end;
    if (perm_mode == PHET_PERM_RESIDENT) {
        PHET_REQUIRE(g_begin >= c->perm_g0 && g_end <= c->perm_g0 + c->perm_rows, "PHET_PERM_RESIDENT: genes outside the resident tables");
        perm_mode = PHET_PERM_HOST;
    }
    a.perm_mode = perm_mode;
    a.perm0 = c->perm0.p;
    a.perm1 = c->perm1.p;
    a.perm_g0 = c->perm_g0;
    a.coprime0 = c->coprime.p;
    a.coprime1 = c->coprime.p ? c->coprime.p + kCoprimeCount : nullptr;
    a.coprime_inv0 = c->coprime_inv.p;
    a.coprime_inv1 = c->coprime_inv.p ? c->coprime_inv.p + kCoprimeCount : nullptr;
    a.slice_of = (perm_mode == PHET_PERM_HOST && c->slice_of_valid) ? c->slice_of.p : nullptr;
    a.slice_ld = c->n0 + c->n1;
    if (perm_mode == PHET_PERM_HOST && !c->perm_fwd_valid) {
        // only the inverse view of the tables was written (replay mode): the scatter kernel has to be the one that runs
        PHET_REQUIRE(a.slice_of != nullptr && n_last <= 512 && step3_wants_inverse(c, m, n_last),
                     "resident permutation tables hold the inverse view only, but this launch needs the forward view "
                     "(PHET_OPT_KEEP_H / PHET_STEP3_IMPL changed since they were drawn?)");
    }
    a.seed = seed;
    a.tab_full = c->tab_full.p;
    a.tab_last = c->tab_last.p;
    a.o = c->o.p;
    a.H = c->H.p;
    if (n_last > 512) {
        // multi-class fit with unbalanced classes: the slice sorters take the full slices, the block sorter the last one
        if (c->n_slices > 1) {
            Step3Args f = a;
            f.n_slices = c->n_slices - 1;
            f.n_last = m;
            f.cap = m;
            if (int e = c->storage == PHET_F32 ? run_step3_f32(c, f) : run_step3_f64(c, f)) return e;
        }
        return launch_ks_big(c, a, c->n_slices > 1 ? 1 : 0);
    }
    return c->storage == PHET_F32 ? run_step3_f32(c, a) : run_step3_f64(c, a);
}

int phase_mark(phet_ctx *c, int k) {
    if (!c->ev_phase[k]) PHET_CUDA(cudaEventCreate(&c->ev_phase[k]));
    PHET_CUDA(cudaEventRecord(c->ev_phase[k], c->stream));
    return PHET_OK;
}

}  // namespace phet

using namespace phet;

extern "C" {

int phet_abi_version(void) { return PHET_B200_ABI_VERSION; }

const char *phet_last_error(void) { return g_last_error.c_str(); }

int phet_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    return n;
}

int phet_ctx_create(int device, void *stream, phet_ctx **out) {
    PHET_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    int n = 0;
    PHET_CUDA(cudaGetDeviceCount(&n));
    PHET_REQUIRE(device >= 0 && device < n, "device index out of range (no CUDA device? there is no CPU fallback)");
    PHET_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PHET_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("phet_b200 kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major) +
                  std::to_string(prop.minor));
        return PHET_ERR_UNSUPPORTED;
    }
    phet_ctx *c = new (std::nothrow) phet_ctx();
    if (!c) {
        set_error("out of host memory");
        return PHET_ERR_NOMEM;
    }
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    if (const char *impl = getenv("PHET_STEP1_IMPL")) c->step1_impl = (strcmp(impl, "warp") == 0) ? 1 : 0;
    if (const char *impl = getenv("PHET_STEP3_IMPL"))   // sort: exact h of every slice; prune: gather bound-and-prune; scatter: streaming bound-and-prune
        c->step3_impl = strcmp(impl, "sort") == 0 ? 1 : strcmp(impl, "prune") == 0 ? 2 : strcmp(impl, "scatter") == 0 ? 3 : 0;
    if (const char *fb = getenv("PHET_STEP1_FORCE_BIG")) c->pair_force_big = atoi(fb) != 0;
    if (const char *bt = getenv("PHET_STEP1_BIG_THREADS")) {
        const int t = atoi(bt);
        if (t >= 64 && t <= 512) c->pair_big_threads = t / 32 * 32;
    }
    if (const char *nbw = getenv("PHET_STEP1_NBW")) c->pair_nbw = (atoi(nbw) == 16) ? 16 : (atoi(nbw) == 32 ? 32 : 0);
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
        }
        c->own_stream = true;
    }
    if (const char *pf = getenv("PHET_PROFILE")) {
        if (atoi(pf) != 0 && c->prof.alloc(16) == 0) cudaMemset(c->prof.p, 0, 16 * sizeof(unsigned long long));
    }
    if (int e = c->scalars.alloc(16)) {
        delete c;
        return e;
    }
    *out = c;
    return PHET_OK;
}

int phet_ctx_destroy(phet_ctx *c) {
    if (!c) return PHET_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    replay_ring_destroy(c);
    c->cls_rows.release();
    c->fy_scratch.release();
    c->draws_dev.release();
    c->rows_grouped.release();
    c->pos_of_row.release();
    c->xg.release();
    c->mean.release();
    c->sd.release();
    c->colpart.release();
    c->xraw.release();
    c->idx.release();
    c->idxT.release();
    c->idxT32.release();
    c->win.release();
    c->pair_scratch.release();
    c->win3.release();
    c->prof.release();
    c->acc.release();
    c->dbgP.release();
    c->dbgD.release();
    c->o.release();
    c->H.release();
    c->perm0.release();
    c->perm1.release();
    c->tab_full.release();
    c->tab_last.release();
    c->cls_pos.release();
    c->coprime.release();
    c->f.release();
    c->r.release();
    c->fw.release();
    c->scores.release();
    c->bins.release();
    c->scalars.release();
    c->edges.release();
    c->weights.release();
    c->counts.release();
    c->class_stats.release();
    c->mtx_out.release();
    c->mtx_dense.release();
    c->mtx_v.release();
    c->mtx_rowsum.release();
    c->mtx_r.release();
    c->mtx_c.release();
    c->mtx_prog.release();
    c->mtx_rows.release();
    c->mtx_cols.release();
    c->mtx_cnt.release();
    c->raw_keep.release();
    c->anova_idx.release();
    c->anova_idxT.release();
    c->anova_idxU.release();
    c->cs_ovf.release();
    c->rank_keys.release();
    c->rank_order.release();
    c->rank_sorted.release();
    c->rank_in.release();
    c->stage[0].release();
    c->stage[1].release();
    for (int i = 0; i < 4; ++i)
        if (c->ev_phase[i]) cudaEventDestroy(c->ev_phase[i]);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return PHET_OK;
}

int phet_ctx_sync(phet_ctx *c) {
    PHET_REQUIRE(c != nullptr, "ctx is NULL");
    PHET_CUDA(cudaSetDevice(c->device));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    return PHET_OK;
}

int phet_set_classes(phet_ctx *c, const int32_t *ctrl_rows, int64_t n0, const int32_t *case_rows, int64_t n1) {
    PHET_REQUIRE(c != nullptr, "ctx is NULL");
    PHET_REQUIRE(ctrl_rows && case_rows, "row lists are NULL");
    PHET_REQUIRE(n0 >= 1 && n1 >= 1, "both classes need at least one cell");
    PHET_CUDA(cudaSetDevice(c->device));
    const int64_t N = n0 + n1;
    // Layout: the cells of each class are placed in a fixed pseudo-random order (seeded
    // Fisher-Yates), so that cheap per-gene affine walks over positions (Step 3, device mode)
    // visit pseudo-random cells.  order[q] = class-local index held at class-relative position q.
    std::vector<int32_t> grouped((size_t)N), pos((size_t)N, -1), cls_pos((size_t)N), cls_rows((size_t)N);
    for (int cls = 0; cls < 2; ++cls) {
        const int64_t n = cls ? n1 : n0, off = cls ? n0 : 0;
        const int32_t *rows = cls ? case_rows : ctrl_rows;
        std::vector<int32_t> order((size_t)n);
        for (int64_t i = 0; i < n; ++i) order[(size_t)i] = (int32_t)i;
        uint64_t st = 0x5EEDC0DEull + (uint64_t)cls;
        for (int64_t i = n - 1; i > 0; --i) {
            st = splitmix64(st);
            const int64_t j = (int64_t)(st % (uint64_t)(i + 1));
            std::swap(order[(size_t)i], order[(size_t)j]);
        }
        for (int64_t q = 0; q < n; ++q) {
            const int32_t t = order[(size_t)q];
            const int32_t r = rows[t];
            if (r < 0 || r >= N || pos[(size_t)r] != -1) {
                set_error("invalid: class row lists must partition [0, n0+n1)");
                return PHET_ERR_INVALID;
            }
            grouped[(size_t)(off + q)] = r;
            pos[(size_t)r] = (int32_t)(off + q);
            cls_pos[(size_t)(off + t)] = (int32_t)q;
            cls_rows[(size_t)(off + t)] = r;
        }
    }
    if (int e = c->rows_grouped.alloc((size_t)N)) return e;
    if (int e = c->pos_of_row.alloc((size_t)N)) return e;
    if (int e = c->cls_pos.alloc((size_t)N)) return e;
    if (int e = c->cls_rows.alloc((size_t)N)) return e;
    PHET_CUDA(cudaMemcpyAsync(c->cls_rows.p, cls_rows.data(), sizeof(int32_t) * N, cudaMemcpyHostToDevice, c->stream));
    PHET_CUDA(cudaMemcpyAsync(c->rows_grouped.p, grouped.data(), sizeof(int32_t) * N, cudaMemcpyHostToDevice, c->stream));
    PHET_CUDA(cudaMemcpyAsync(c->pos_of_row.p, pos.data(), sizeof(int32_t) * N, cudaMemcpyHostToDevice, c->stream));
    PHET_CUDA(cudaMemcpyAsync(c->cls_pos.p, cls_pos.data(), sizeof(int32_t) * N, cudaMemcpyHostToDevice, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    c->n0 = n0;
    c->n1 = n1;
    c->classes_set = true;
    c->matrix_set = false;
    c->idx_set = false;
    return PHET_OK;
}

extern "C++" {
namespace phet {
// Allocate the per-fit device buffers for genes [g_begin, g_end) of an N x G problem.
int prepare_matrix(phet_ctx *c, int64_t N, int64_t G, int dtype, int normalize, int storage, int64_t g_begin,
                          int64_t g_end) {
    PHET_REQUIRE(c->classes_set, "phet_set_classes must be called before loading the matrix");
    PHET_REQUIRE(N == c->n0 + c->n1, "N does not match the class sizes");
    PHET_REQUIRE(G >= 1, "G must be >= 1");
    PHET_REQUIRE(dtype == PHET_F32 || dtype == PHET_F64, "dtype must be PHET_F32 or PHET_F64");
    PHET_REQUIRE(storage == PHET_F32 || storage == PHET_F64, "storage must be PHET_F32 or PHET_F64");
    PHET_REQUIRE(normalize == PHET_NORM_NONE || normalize == PHET_NORM_ZSCORE, "normalize must be none or zscore");
    PHET_REQUIRE(N < (int64_t)1 << 31, "N too large");
    PHET_REQUIRE(0 <= g_begin && g_begin <= g_end && g_end <= G, "gene range outside [0, G]");
    PHET_CUDA(cudaSetDevice(c->device));
    const size_t e_st = storage == PHET_F32 ? 4 : 8;
    c->N = N;
    c->G = G;
    c->storage = storage;
    c->fast_moments = (storage == PHET_F32 && normalize == PHET_NORM_ZSCORE && getenv("PHET_EXACT_MOMENTS") == nullptr);
    const int64_t per16 = 16 / (int64_t)e_st;
    c->n_pad = (N + per16 - 1) / per16 * per16;
    c->xg_g0 = g_begin;
    c->xg_rows = g_end - g_begin;
    if (int e = c->xg.alloc((size_t)(c->xg_rows > 0 ? c->xg_rows : 1) * (size_t)c->n_pad * e_st)) return e;
    if (int e = c->mean.alloc((size_t)G)) return e;
    if (int e = c->sd.alloc((size_t)G)) return e;
    if (c->acc.n < (size_t)3 * G || c->o.p != c->acc.p + 2 * G) {
        c->o.release();
        if (int e = c->acc.alloc((size_t)3 * G)) return e;
        c->o.alias(c->acc.p + 2 * G, (size_t)G);  // o rides behind the two accumulators
    }
    if (int e = c->f.alloc((size_t)G)) return e;
    if (int e = c->r.alloc((size_t)G)) return e;
    if (int e = c->fw.alloc((size_t)G)) return e;
    if (int e = c->scores.alloc((size_t)G)) return e;
    if (int e = c->bins.alloc((size_t)G)) return e;
    PHET_CUDA(cudaMemsetAsync(c->acc.p, 0, sizeof(double) * 2 * G, c->stream));
    PHET_CUDA(cudaMemsetAsync(c->o.p, 0, sizeof(double) * G, c->stream));
    PHET_CUDA(cudaMemsetAsync(c->mean.p, 0, sizeof(double) * G, c->stream));
    PHET_CUDA(cudaMemsetAsync(c->sd.p, 0, sizeof(double) * G, c->stream));
    return PHET_OK;
}
}  // namespace phet
}  // extern "C++"

int phet_load_matrix(phet_ctx *c, const void *X, int is_device, int64_t N, int64_t G, int dtype, int normalize,
                     int storage) {
    PHET_REQUIRE(c != nullptr, "ctx is NULL");
    PHET_REQUIRE(X != nullptr, "X is NULL");
    if (int e = prepare_matrix(c, N, G, dtype, normalize, storage, 0, G)) return e;
    const size_t e_in = dtype == PHET_F32 ? 4 : 8;
    const void *Xdev = X;
    if (!is_device) {
        const size_t bytes = (size_t)N * (size_t)G * e_in;
        if (int e = c->xraw.alloc(bytes)) return e;
        PHET_CUDA(cudaMemcpyAsync(c->xraw.p, X, bytes, cudaMemcpyHostToDevice, c->stream));
        Xdev = c->xraw.p;
    }
    if (int e = launch_normalize(c, Xdev, G, 0, G, dtype, normalize)) return e;
    if (!is_device) {
        PHET_CUDA(cudaStreamSynchronize(c->stream));
        c->xraw.release();  // the normalised gene-major copy is all later steps need
    }
    c->matrix_set = true;
    return PHET_OK;
}

int phet_load_matrix_genes(phet_ctx *c, const void *X_dev, int64_t ldx, int64_t N, int64_t G, int dtype, int normalize,
                           int storage, int64_t g_begin, int64_t g_end) {
    PHET_REQUIRE(c != nullptr, "ctx is NULL");
    PHET_REQUIRE(X_dev != nullptr, "X is NULL");
    PHET_REQUIRE(0 <= g_begin && g_begin < g_end && g_end <= G, "gene range outside [0, G)");
    PHET_REQUIRE(ldx >= g_end - g_begin, "leading dimension shorter than the gene range");
    if (int e = prepare_matrix(c, N, G, dtype, normalize, storage, g_begin, g_end)) return e;
    if (int e = launch_normalize(c, X_dev, ldx, g_begin, g_end - g_begin, dtype, normalize)) return e;
    c->matrix_set = true;
    return PHET_OK;
}

extern "C++" {
namespace phet {
// c->idx holds int32[S][2][m] ORIGINAL row numbers (uploaded by phet_set_index_sets, or written by the replay
// path): remap to grouped positions, validate, and build the transposed tables of the pair-per-thread kernels.
int finish_index_sets(phet_ctx *c, int64_t S, int64_t m) {
    const int64_t n = S * 2 * m;
    int *bad = reinterpret_cast<int *>(c->scalars.p + 12);
    PHET_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), c->stream));
    k_remap_idx<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->idx.p, n, c->pos_of_row.p, c->n0 + c->n1, bad);
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    int hbad = 0;
    PHET_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    PHET_REQUIRE(hbad == 0, "index sets contain a row number outside [0, N)");
    c->idxT_ok = false;
    // 16-bit table (classes of 65,536 cells and more take the 32-bit table of the global-gather instantiation)
    if (c->n0 <= 65535 && c->n1 <= 65535 && m <= 255) {  // transposed 16-bit copy for the pair-per-thread Step 1
        const int64_t nblk = (S + 31) / 32;
        const int64_t mp = (m + 15) / 16 * 16;
        const int64_t nT = 2 * nblk * mp * 32;
        if (int e = c->idxT.alloc((size_t)nT)) return e;
        k_build_idxT<<<(unsigned)((2 * nblk + 3) / 4), 128, 0, c->stream>>>(c->idx.p, S, m, mp, c->n0, nblk, c->idxT.p);
        c->launches++;
        PHET_CUDA(cudaGetLastError());
        c->idxT_ok = true;
    }
    c->idxT32_ok = false;
    const int64_t ncmax = c->n0 > c->n1 ? c->n0 : c->n1;
    if ((!c->idxT_ok || c->pair_force_big || ncmax * 4 > 90 * 1024) && m <= 496) {  // 32-bit ascending copy for the global-gather variant (big classes or m > 255)
        const int64_t nblk = (S + 31) / 32;
        const int64_t mp = (m + 15) / 16 * 16;
        const int64_t nT = 2 * nblk * mp * 32;
        if (int e = c->idxT32.alloc((size_t)nT)) return e;
        k_build_idxT32<<<(unsigned)((2 * nblk * 32 + 63) / 64), 64, 0, c->stream>>>(c->idx.p, S, m, mp, c->n0, nblk, c->idxT32.p);
        c->launches++;
        PHET_CUDA(cudaGetLastError());
        c->idxT32_ok = true;
    }
    c->S = S;
    c->m = m;
    c->idx_set = true;
    return PHET_OK;
}
}  // namespace phet
}  // extern "C++"

int phet_set_index_sets(phet_ctx *c, const int32_t *idx, int64_t S, int64_t m) {
    PHET_REQUIRE(c != nullptr, "ctx is NULL");
    PHET_REQUIRE(c->classes_set, "phet_set_classes must be called first");
    PHET_REQUIRE(idx != nullptr && S >= 1 && m >= 1, "index sets: need S >= 1, m >= 1");
    PHET_REQUIRE(m <= 512, "subsample size m > 512 is not supported");
    PHET_CUDA(cudaSetDevice(c->device));
    const int64_t n = S * 2 * m;
    if (int e = c->idx.alloc((size_t)n)) return e;
    PHET_CUDA(cudaMemcpyAsync(c->idx.p, idx, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    return finish_index_sets(c, S, m);
}

int phet_step1(phet_ctx *c, int64_t s_begin, int64_t s_end, const phet_step1_params *prm, int accumulate) {
    PHET_REQUIRE(c != nullptr && prm != nullptr, "ctx/params NULL");
    PHET_REQUIRE(c->matrix_set && c->idx_set, "load the matrix and index sets before phet_step1");
    PHET_REQUIRE(0 <= s_begin && s_begin <= s_end && s_end <= c->S, "subsample range outside [0, S]");
    PHET_REQUIRE(prm->test == PHET_TEST_T || prm->test == PHET_TEST_Z, "bad test kind");
    const phet_quantile_pos &q = prm->q;
    PHET_REQUIRE(q.j_lo >= 0 && q.k_lo < c->m && q.j_hi >= 0 && q.k_hi < c->m && q.j_lo <= q.k_lo && q.j_hi <= q.k_hi,
                 "quantile positions outside [0, m)");
    PHET_CUDA(cudaSetDevice(c->device));
    return launch_step1(c, c->xg_g0, c->xg_g0 + c->xg_rows, s_begin, s_end, prm, accumulate);
}

extern "C++" {
namespace phet {
// Allocations, table/permutation uploads and zeroing shared by every Step-3 launch of a fit.
int step3_prepare(phet_ctx *c, int64_t m, int perm_mode, const uint32_t *perm_ctrl, const uint32_t *perm_case,
                         const double *table_full, const double *table_last, int64_t n_last) {
    PHET_REQUIRE(m >= 1 && m <= 512, "slice length m must be in [1, 512]");
    PHET_REQUIRE(perm_mode == PHET_PERM_HOST || perm_mode == PHET_PERM_DEVICE || perm_mode == PHET_PERM_REPLAY ||
                     perm_mode == PHET_PERM_RESIDENT, "bad perm_mode");
    PHET_REQUIRE(table_full && table_last, "KS tables are NULL");
    if (perm_mode != PHET_PERM_RESIDENT) {
        c->perm_g0 = 0;
        c->perm_rows = 0;
        c->slice_of_valid = false;
        c->perm_fwd_valid = false;
    }
    int64_t min_c = c->n0 < c->n1 ? c->n0 : c->n1;
    int64_t n_slices = (min_c + m - 1) / m;
    if (c->step3_min_size > 0) {
        // multi-class fit (phet.py:466-492): slice starts cover [0, min_size) of the smallest of ALL classes, and the
        // last slice runs to the end of the shorter column of the pair, so n_last may exceed m
        PHET_REQUIRE(c->step3_min_size <= min_c, "PHET_OPT_STEP3_MIN_SIZE exceeds min(n0, n1)");
        n_slices = (c->step3_min_size + m - 1) / m;
        PHET_REQUIRE(n_last >= 1 && (n_slices - 1) * m + n_last <= min_c, "n_last runs past the shorter class");
        min_c = (n_slices - 1) * m + n_last;
    } else {
        PHET_REQUIRE(n_last == min_c - (n_slices - 1) * m, "n_last does not match min(n0,n1) and m");
    }
    c->n_slices = n_slices;
    {   // h0: table_full is non-increasing on [h0, m] and table_full[h0] <= every earlier entry (the exact
        // formula is only "essentially" monotone: rounding noise of 1 ulp at h = 1, 2 for some n)
        int64_t h0 = 0;
        for (int64_t h = 0; h < m; ++h)
            if (table_full[h + 1] > table_full[h]) h0 = h + 1;
        double floor0 = INFINITY;
        for (int64_t h = 0; h < h0; ++h) floor0 = fmin(floor0, table_full[h]);
        while (h0 <= m && table_full[h0] > floor0) ++h0;
        c->tab_h0 = (int)h0;  // may be m + 1: every slice is then evaluated exactly
    }
    if (int e = c->H.alloc((size_t)c->G * (size_t)n_slices)) return e;
    if (int e = c->tab_full.alloc((size_t)m + 1)) return e;
    if (int e = c->tab_last.alloc((size_t)n_last + 1)) return e;
    PHET_CUDA(cudaMemsetAsync(c->H.p, 0, sizeof(int32_t) * c->G * n_slices, c->stream));
    PHET_CUDA(cudaMemsetAsync(c->o.p, 0, sizeof(double) * c->G, c->stream));
    // the tables only change with (m, n_last): keep a host copy and upload on change, synchronously, so that
    // no caller-owned host buffer is in flight when the launch calls return
    if (c->tab_full_host.size() != (size_t)m + 1 || memcmp(c->tab_full_host.data(), table_full, sizeof(double) * (m + 1)) != 0) {
        c->tab_full_host.assign(table_full, table_full + m + 1);
        PHET_CUDA(cudaMemcpyAsync(c->tab_full.p, c->tab_full_host.data(), sizeof(double) * (m + 1), cudaMemcpyHostToDevice, c->stream));
        PHET_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (c->tab_last_host.size() != (size_t)n_last + 1 || memcmp(c->tab_last_host.data(), table_last, sizeof(double) * (n_last + 1)) != 0) {
        c->tab_last_host.assign(table_last, table_last + n_last + 1);
        PHET_CUDA(cudaMemcpyAsync(c->tab_last.p, c->tab_last_host.data(), sizeof(double) * (n_last + 1), cudaMemcpyHostToDevice, c->stream));
        PHET_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (perm_mode == PHET_PERM_HOST) {
        PHET_REQUIRE(perm_ctrl && perm_case, "host permutations are NULL");
        const size_t n = (size_t)c->G * (size_t)min_c;
        if (int e = c->perm0.alloc(n)) return e;
        if (int e = c->perm1.alloc(n)) return e;
        PHET_CUDA(cudaMemcpyAsync(c->perm0.p, perm_ctrl, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c->stream));
        PHET_CUDA(cudaMemcpyAsync(c->perm1.p, perm_case, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c->stream));
        int *bad = reinterpret_cast<int *>(c->scalars.p + 12);
        PHET_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), c->stream));
        const unsigned blocks = (unsigned)((n + 255) / 256);
        k_remap_perm<<<blocks, 256, 0, c->stream>>>(c->perm0.p, (int64_t)n, c->cls_pos.p, (uint32_t)c->n0, bad);
        k_remap_perm<<<blocks, 256, 0, c->stream>>>(c->perm1.p, (int64_t)n, c->cls_pos.p + c->n0, (uint32_t)c->n1, bad);
        c->launches += 2;
        PHET_CUDA(cudaGetLastError());
        int hbad = 0;
        PHET_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        PHET_CUDA(cudaStreamSynchronize(c->stream));
        PHET_REQUIRE(hbad == 0, "host permutations contain a position outside its class");
        c->perm_rows = c->G;
        c->perm_cols = min_c;
        c->perm_fwd_valid = true;
        if (n_last <= 512 && step3_wants_inverse(c, m, n_last)) {
            // the scatter kernel reads the INVERSE view: the slice of every cell in storage order
            const int64_t ld = c->n0 + c->n1;
            if (int e = c->slice_of.alloc((size_t)c->G * (size_t)ld)) return e;
            PHET_CUDA(cudaMemsetAsync(c->slice_of.p, 0xFF, sizeof(uint16_t) * (size_t)c->G * (size_t)ld, c->stream));
            const uint32_t magic = (uint32_t)(((uint64_t)1 << 32) / (uint64_t)m) + 1u;
            k_slice_of_from_perm<<<blocks, 256, 0, c->stream>>>(c->perm0.p, (int64_t)n, min_c, c->slice_of.p, ld, magic,
                                                                 (uint32_t)(n_slices - 1));
            k_slice_of_from_perm<<<blocks, 256, 0, c->stream>>>(c->perm1.p, (int64_t)n, min_c, c->slice_of.p + c->n0, ld, magic,
                                                                 (uint32_t)(n_slices - 1));
            c->launches += 2;
            PHET_CUDA(cudaGetLastError());
            c->slice_of_valid = true;
        }
    } else if (perm_mode == PHET_PERM_RESIDENT) {
        PHET_REQUIRE(c->perm_rows > 0 && c->perm_cols == min_c && (c->perm_fwd_valid || c->slice_of_valid),
                     "PHET_PERM_RESIDENT: no permutation tables of this shape are resident");
    } else if (perm_mode == PHET_PERM_REPLAY) {
        // the permutation tables are filled block by block on the device (replay.cu)
    } else {
        if (c->coprime.p == nullptr || c->coprime_n0 != c->n0 || c->coprime_n1 != c->n1) {
            std::vector<uint32_t> cp(2 * kCoprimeCount), cpi(2 * kCoprimeCount);
            coprime_list((uint32_t)c->n0, cp.data());
            coprime_list((uint32_t)c->n1, cp.data() + kCoprimeCount);
            for (int k = 0; k < 2 * kCoprimeCount; ++k)   // cell -> permuted index needs the inverse multipliers (scatter kernel)
                cpi[k] = inverse_mod(cp[k], (uint32_t)(k < kCoprimeCount ? c->n0 : c->n1));
            if (int e = c->coprime.alloc(2 * kCoprimeCount)) return e;
            if (int e = c->coprime_inv.alloc(2 * kCoprimeCount)) return e;
            PHET_CUDA(cudaMemcpyAsync(c->coprime.p, cp.data(), sizeof(uint32_t) * cp.size(), cudaMemcpyHostToDevice,
                                      c->stream));
            PHET_CUDA(cudaMemcpyAsync(c->coprime_inv.p, cpi.data(), sizeof(uint32_t) * cpi.size(), cudaMemcpyHostToDevice,
                                      c->stream));
            PHET_CUDA(cudaStreamSynchronize(c->stream));  // cp / cpi are stack-lifetime host buffers
            c->coprime_n0 = c->n0;
            c->coprime_n1 = c->n1;
        }
    }
    return PHET_OK;
}
}  // namespace phet
}  // extern "C++"

int phet_step3(phet_ctx *c, int64_t g_begin, int64_t g_end, int64_t m, int perm_mode, const uint32_t *perm_ctrl,
               const uint32_t *perm_case, uint64_t seed, const double *table_full, const double *table_last,
               int64_t n_last) {
    PHET_REQUIRE(c != nullptr, "ctx is NULL");
    PHET_REQUIRE(c->matrix_set, "load the matrix before phet_step3");
    PHET_REQUIRE(c->xg_g0 <= g_begin && g_begin <= g_end && g_end <= c->xg_g0 + c->xg_rows,
                 "gene range outside the resident genes");
    PHET_REQUIRE(perm_mode != PHET_PERM_REPLAY, "PHET_PERM_REPLAY is a phet_run_pipeline mode");
    PHET_CUDA(cudaSetDevice(c->device));
    if (int e = step3_prepare(c, m, perm_mode, perm_ctrl, perm_case, table_full, table_last, n_last)) return e;
    int rc = launch_step3(c, g_begin, g_end, m, perm_mode, seed, n_last);
    if (rc) return rc;
    // host permutations are caller-owned: finish their copies before returning (tables were copied in prepare)
    if (perm_mode == PHET_PERM_HOST) PHET_CUDA(cudaStreamSynchronize(c->stream));
    return PHET_OK;
}

// Gene-block pipeline: H2D copy of block b+1 (copy stream, cudaMemcpy2DAsync out of the caller's
// row-major matrix) overlaps normalise + Step 1 + Step 3 of block b (compute stream).
int phet_run_pipeline(phet_ctx *c, const void *X, int is_device, int64_t N, int64_t G, int dtype,
                      const phet_pipeline_params *p) {
    PHET_REQUIRE(c != nullptr && X != nullptr && p != nullptr, "NULL argument");
    if (p->perm_mode == PHET_PERM_REPLAY || p->draw_index_sets) return run_pipeline_replay(c, X, is_device, N, G, dtype, p);
    PHET_REQUIRE(c->idx_set, "phet_set_index_sets must precede phet_run_pipeline");
    PHET_REQUIRE(p->m == c->m, "m does not match the uploaded index sets");
    PHET_REQUIRE(0 <= p->s_begin && p->s_begin <= p->s_end && p->s_end <= c->S, "subsample range outside [0, S]");
    PHET_REQUIRE(p->g_begin <= p->g3_begin && p->g3_begin <= p->g3_end && p->g3_end <= p->g_end,
                 "Step-3 gene range must lie inside the loaded gene range");
    if (int e = prepare_matrix(c, N, G, dtype, p->normalize, p->storage, p->g_begin, p->g_end)) return e;
    c->phases_valid = false;
    const phet_quantile_pos &q = p->step1.q;
    PHET_REQUIRE(q.j_lo >= 0 && q.k_lo < c->m && q.j_hi >= 0 && q.k_hi < c->m && q.j_lo <= q.k_lo && q.j_hi <= q.k_hi,
                 "quantile positions outside [0, m)");
    if (int e = step3_prepare(c, p->m, p->perm_mode, p->perm_ctrl, p->perm_case, p->table_full, p->table_last,
                              p->n_last))
        return e;
    const size_t e_in = dtype == PHET_F32 ? 4 : 8;
    const int64_t gcount = p->g_end - p->g_begin;
    auto compute_block = [&](const void *Xblock, int64_t ldx, int64_t g0, int64_t gc) -> int {
        if (int e = launch_normalize(c, Xblock, ldx, g0, gc, dtype, p->normalize)) return e;
        return 0;
    };
    auto steps_block = [&](int64_t g0, int64_t gc) -> int {
        if (int e = launch_step1(c, g0, g0 + gc, p->s_begin, p->s_end, &p->step1, 0)) return e;
        const int64_t a = g0 > p->g3_begin ? g0 : p->g3_begin;
        const int64_t b = (g0 + gc) < p->g3_end ? (g0 + gc) : p->g3_end;
        if (b > a)
            if (int e = launch_step3(c, a, b, p->m, p->perm_mode, p->seed, p->n_last)) return e;
        return 0;
    };
    if (gcount > 0) {
        if (is_device) {
            const unsigned char *Xb = static_cast<const unsigned char *>(X) + (size_t)p->g_begin * e_in;
            if (int e = phase_mark(c, 0)) return e;
            if (int e = compute_block(Xb, G, p->g_begin, gcount)) return e;
            if (int e = phase_mark(c, 1)) return e;
            if (int e = launch_step1(c, p->g_begin, p->g_end, p->s_begin, p->s_end, &p->step1, 0)) return e;
            if (int e = phase_mark(c, 2)) return e;
            if (p->g3_end > p->g3_begin)
                if (int e = launch_step3(c, p->g3_begin, p->g3_end, p->m, p->perm_mode, p->seed, p->n_last)) return e;
            if (int e = phase_mark(c, 3)) return e;
            c->phases_valid = true;
        } else {
            if (!c->copy_stream) PHET_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                if (!c->ev_copied[i]) PHET_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
                if (!c->ev_consumed[i]) PHET_CUDA(cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
            }
            int64_t gb = p->block_genes;
            if (gb <= 0) {  // ~128 MB per staging buffer, a multiple of the persistent grid (one CTA per SM)
                size_t block_bytes = (size_t)128 << 20;
                if (const char *mb = getenv("PHET_BLOCK_MB")) block_bytes = (size_t)(atoi(mb) > 0 ? atoi(mb) : 128) << 20;
                gb = (int64_t)(block_bytes / ((size_t)N * e_in));
                gb = gb / c->num_sms * c->num_sms;
                if (gb < c->num_sms) gb = c->num_sms;
            }
            if (gb > gcount) gb = gcount;
            const size_t stage_bytes = (size_t)N * (size_t)gb * e_in;
            if (int e = c->stage[0].alloc(stage_bytes)) return e;
            if (gcount > gb)
                if (int e = c->stage[1].alloc(stage_bytes)) return e;
            int b = 0;
            for (int64_t g0 = p->g_begin; g0 < p->g_end; g0 += gb, ++b) {
                const int64_t gc = (g0 + gb <= p->g_end) ? gb : (p->g_end - g0);
                const int buf = b & 1;
                if (b >= 2) PHET_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[buf], 0));
                PHET_CUDA(cudaMemcpy2DAsync(c->stage[buf].p, (size_t)gc * e_in,
                                            static_cast<const unsigned char *>(X) + (size_t)g0 * e_in, (size_t)G * e_in,
                                            (size_t)gc * e_in, (size_t)N, cudaMemcpyHostToDevice, c->copy_stream));
                PHET_CUDA(cudaEventRecord(c->ev_copied[buf], c->copy_stream));
                PHET_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[buf], 0));
                if (int e = compute_block(c->stage[buf].p, gc, g0, gc)) return e;
                PHET_CUDA(cudaEventRecord(c->ev_consumed[buf], c->stream));
                if (int e = steps_block(g0, gc)) return e;
            }
        }
    }
    // host X and host permutations are caller-owned and still being read; a device matrix with device-mode
    // permutations leaves nothing of the caller's in flight, so the call returns with the work queued
    if (!is_device || p->perm_mode == PHET_PERM_HOST) PHET_CUDA(cudaStreamSynchronize(c->stream));
    c->matrix_set = true;
    return PHET_OK;
}

int phet_phase_ms(phet_ctx *c, float out[3]) {
    PHET_REQUIRE(c != nullptr && out != nullptr, "NULL argument");
    out[0] = out[1] = out[2] = -1.0f;
    if (!c->phases_valid) return PHET_OK;
    PHET_CUDA(cudaSetDevice(c->device));
    PHET_CUDA(cudaEventSynchronize(c->ev_phase[3]));
    for (int k = 0; k < 3; ++k) PHET_CUDA(cudaEventElapsedTime(&out[k], c->ev_phase[k], c->ev_phase[k + 1]));
    return PHET_OK;
}

int phet_fisher(phet_ctx *c, int64_t S_total) {
    PHET_REQUIRE(c != nullptr && c->matrix_set, "ctx not ready");
    PHET_REQUIRE(S_total >= 1, "S_total must be >= 1");
    PHET_CUDA(cudaSetDevice(c->device));
    return launch_fisher(c, S_total);
}

int phet_o_minmax(phet_ctx *c, double out_minmax[2]) {
    PHET_REQUIRE(c != nullptr && c->matrix_set && out_minmax, "ctx not ready");
    PHET_CUDA(cudaSetDevice(c->device));
    return launch_o_minmax(c, out_minmax);
}

int phet_bin(phet_ctx *c, const double *edges, int32_t B, int64_t *counts_out) {
    PHET_REQUIRE(c != nullptr && c->matrix_set, "ctx not ready");
    PHET_REQUIRE(edges != nullptr && B >= 1 && B <= 127, "need 1 <= B <= 127 bins");
    PHET_CUDA(cudaSetDevice(c->device));
    if (int e = c->edges.alloc((size_t)B + 1)) return e;
    if (int e = c->counts.alloc((size_t)B)) return e;
    PHET_CUDA(cudaMemcpyAsync(c->edges.p, edges, sizeof(double) * (B + 1), cudaMemcpyHostToDevice, c->stream));
    c->n_bins = B;
    int rc = launch_bin(c, B, counts_out);
    if (rc) return rc;
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    return PHET_OK;
}

int phet_combine(phet_ctx *c, const double *weights, int32_t B, double *scores_out) {
    PHET_REQUIRE(c != nullptr && c->matrix_set, "ctx not ready");
    PHET_REQUIRE(weights != nullptr && B >= 1 && B >= c->n_bins, "weights must cover every bin");
    PHET_CUDA(cudaSetDevice(c->device));
    if (int e = c->weights.alloc((size_t)B)) return e;
    PHET_CUDA(cudaMemcpyAsync(c->weights.p, weights, sizeof(double) * B, cudaMemcpyHostToDevice, c->stream));
    int rc = launch_combine(c, B, scores_out);
    if (rc) return rc;
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    return PHET_OK;
}

int phet_finalize_async(phet_ctx *c, int64_t S_total, const double *weights, int32_t B) {
    PHET_REQUIRE(c != nullptr && c->matrix_set, "ctx not ready");
    PHET_REQUIRE(S_total >= 1, "S_total must be >= 1");
    PHET_REQUIRE(weights != nullptr && B >= 1 && B <= 127, "need 1 <= B <= 127 feature weights");
    PHET_CUDA(cudaSetDevice(c->device));
    if (int e = c->edges.alloc((size_t)B + 1)) return e;
    if (int e = c->counts.alloc((size_t)B)) return e;
    if (int e = c->weights.alloc((size_t)B)) return e;
    // the weights only change with the estimator's hyper-parameters: keep a host copy, upload on change (synchronously,
    // so that no caller-owned host buffer is in flight when the call returns)
    if (c->weights_host.size() != (size_t)B || memcmp(c->weights_host.data(), weights, sizeof(double) * B) != 0) {
        c->weights_host.assign(weights, weights + B);
        PHET_CUDA(cudaMemcpyAsync(c->weights.p, c->weights_host.data(), sizeof(double) * B, cudaMemcpyHostToDevice, c->stream));
        PHET_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->n_bins = B;
    return launch_finalize(c, S_total, B);
}

int phet_fetch_results(phet_ctx *c, double *scores_out, int64_t *counts_out, double *edges_out) {
    PHET_REQUIRE(c != nullptr && c->matrix_set && c->n_bins >= 1, "phet_fetch_results: no finalized fit in this context");
    PHET_CUDA(cudaSetDevice(c->device));
    const int32_t B = c->n_bins;
    if (scores_out) PHET_CUDA(cudaMemcpyAsync(scores_out, c->scores.p, sizeof(double) * c->G, cudaMemcpyDeviceToHost, c->stream));
    if (counts_out)
        PHET_CUDA(cudaMemcpyAsync(counts_out, c->counts.p, sizeof(unsigned long long) * B, cudaMemcpyDeviceToHost, c->stream));
    if (edges_out) PHET_CUDA(cudaMemcpyAsync(edges_out, c->edges.p, sizeof(double) * (B + 1), cudaMemcpyDeviceToHost, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    return PHET_OK;
}

int phet_finalize(phet_ctx *c, int64_t S_total, const double *weights, int32_t B, double *scores_out,
                  int64_t *counts_out, double *edges_out) {
    if (int e = phet_finalize_async(c, S_total, weights, B)) return e;
    return phet_fetch_results(c, scores_out, counts_out, edges_out);
}

int phet_class_stats(phet_ctx *c, const phet_quantile_pos *q3, double *out) {
    PHET_REQUIRE(c != nullptr && c->matrix_set && c->classes_set, "load the classes and the matrix before phet_class_stats");
    PHET_REQUIRE(q3 != nullptr && out != nullptr, "q3/out NULL");
    PHET_REQUIRE(c->n0 >= 1 && c->n1 >= 1, "both classes need at least one cell");
    PHET_REQUIRE(c->n0 + c->n1 < (int64_t(1) << 25), "phet_class_stats: at most 2^25 - 1 cells (16-bit per-thread digit counters)");
    PHET_REQUIRE(c->xg_g0 == 0 && c->xg_rows == c->G, "phet_class_stats needs every gene resident");
    const int64_t nv[3] = {c->n0, c->n1, c->n0 + c->n1};
    for (int i = 0; i < 3; ++i) {
        const phet_quantile_pos &q = q3[i];
        PHET_REQUIRE(q.j_lo >= 0 && q.j_lo <= q.k_lo && q.k_lo <= q.j_lo + 1 && q.k_lo < nv[i] && q.j_hi >= 0 &&
                         q.j_hi <= q.k_hi && q.k_hi <= q.j_hi + 1 && q.k_hi < nv[i],
                     "quantile positions outside the vector");
    }
    PHET_CUDA(cudaSetDevice(c->device));
    if (int e = c->class_stats.alloc((size_t)7 * c->G)) return e;
    if (int e = launch_class_stats(c, 0, c->G, q3, c->class_stats.p)) return e;
    PHET_CUDA(cudaMemcpyAsync(out, c->class_stats.p, sizeof(double) * 7 * c->G, cudaMemcpyDeviceToHost, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    return PHET_OK;
}

// ---- input path: MatrixMarket -> device (mtx.cu)
int phet_mtx_parse(const char *path, int threads, phet_mtx **out, int64_t dims_out[3]) {
    PHET_REQUIRE(path != nullptr && out != nullptr, "NULL argument");
    *out = nullptr;
    phet_mtx *m = new (std::nothrow) phet_mtx();
    if (!m) {
        set_error("out of host memory");
        return PHET_ERR_NOMEM;
    }
    int rc = PHET_OK;
    try {
        rc = mtx_parse(path, threads, *m);
    } catch (const std::bad_alloc &) {
        set_error("out of host memory while reading the MatrixMarket file");
        rc = PHET_ERR_NOMEM;
    }
    if (rc != PHET_OK) {
        delete m;
        return rc;
    }
    if (dims_out) {
        dims_out[0] = m->rows;
        dims_out[1] = m->cols;
        dims_out[2] = (int64_t)m->v.size();
    }
    *out = m;
    return PHET_OK;
}

int phet_mtx_triplets(const phet_mtx *m, const int32_t **rows, const int32_t **cols, const float **vals, int64_t *count,
                      int *symmetric) {
    PHET_REQUIRE(m != nullptr, "NULL argument");
    if (rows) *rows = m->r.data();
    if (cols) *cols = m->c.data();
    if (vals) *vals = m->v.data();
    if (count) *count = (int64_t)m->v.size();
    if (symmetric) *symmetric = m->symmetric ? 1 : 0;
    return PHET_OK;
}

int phet_mtx_free(phet_mtx *m) {
    delete m;
    return PHET_OK;
}

int phet_mtx_stage(phet_ctx *c, const phet_mtx *m, int do_filter, int64_t minimum_samples, int64_t dims_out[2]) {
    PHET_REQUIRE(c != nullptr && m != nullptr && dims_out != nullptr, "NULL argument");
    PHET_REQUIRE(minimum_samples >= 0, "minimum_samples must be >= 0");
    try {
        return mtx_stage(c, *m, do_filter, minimum_samples, dims_out);
    } catch (const std::bad_alloc &) {
        set_error("out of host memory while staging the matrix");
        return PHET_ERR_NOMEM;
    }
}

int phet_mtx_kept(phet_ctx *c, int32_t *rows_out, int32_t *cols_out) {
    PHET_REQUIRE(c != nullptr, "NULL argument");
    PHET_REQUIRE((int64_t)c->mtx_rows_host.size() == c->mtx_N && (int64_t)c->mtx_cols_host.size() == c->mtx_G,
                 "no staged matrix (call phet_mtx_stage first)");
    if (rows_out && c->mtx_N) memcpy(rows_out, c->mtx_rows_host.data(), sizeof(int32_t) * (size_t)c->mtx_N);
    if (cols_out && c->mtx_G) memcpy(cols_out, c->mtx_cols_host.data(), sizeof(int32_t) * (size_t)c->mtx_G);
    return PHET_OK;
}

// Copy a host matrix to HBM once, as given (row-major cells x genes): PHET_BUF_RAW.  For callers that lay the same
// matrix out several times (a multi-class fit regroups the cells once per combination).
int phet_stage_matrix(phet_ctx *c, const void *X, int64_t N, int64_t G, int dtype) {
    PHET_REQUIRE(c != nullptr && X != nullptr, "NULL argument");
    PHET_REQUIRE(N >= 1 && G >= 1, "N, G must be >= 1");
    PHET_REQUIRE(dtype == PHET_F32 || dtype == PHET_F64, "dtype must be PHET_F32 or PHET_F64");
    PHET_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)N * (size_t)G * (dtype == PHET_F32 ? 4 : 8);
    if (int e = c->raw_keep.alloc(bytes)) return e;
    PHET_CUDA(cudaMemcpyAsync(c->raw_keep.p, X, bytes, cudaMemcpyHostToDevice, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    c->raw_N = N;
    c->raw_G = G;
    c->raw_dtype = dtype;
    return PHET_OK;
}

// ---- multi-class Step 1 (anova.cu)
int phet_anova(phet_ctx *c, const int32_t *idx, int64_t S, int64_t K, int64_t m, double ln_beta, double p_flush_below,
               int capture_debug) {
    PHET_REQUIRE(c != nullptr && idx != nullptr, "NULL argument");
    PHET_REQUIRE(c->matrix_set && c->classes_set, "load the classes and the matrix before phet_anova");
    PHET_REQUIRE(S >= 1 && K >= 2 && m >= 2 && m <= 512 && K <= 1024, "need S >= 1, 2 <= K <= 1024, 2 <= m <= 512");
    PHET_CUDA(cudaSetDevice(c->device));
    return launch_anova(c, idx, S, K, m, ln_beta, p_flush_below, capture_debug, 0);
}

int phet_kruskal(phet_ctx *c, const int32_t *idx, int64_t S, int64_t K, int64_t m, double p_flush_below, int capture_debug) {
    PHET_REQUIRE(c != nullptr && idx != nullptr, "NULL argument");
    PHET_REQUIRE(c->matrix_set && c->classes_set, "load the classes and the matrix before phet_kruskal");
    PHET_REQUIRE(S >= 1 && K >= 2 && m >= 1 && m <= 512, "need S >= 1, K >= 2, 1 <= m <= 512");
    PHET_CUDA(cudaSetDevice(c->device));
    return launch_anova(c, idx, S, K, m, 0.0, p_flush_below, capture_debug, 1);
}

// ---- ranking (rank.cu)
int phet_rank(phet_ctx *c, const double *scores, int64_t G, int64_t k, int32_t *order_out, double *sorted_out) {
    PHET_REQUIRE(c != nullptr, "NULL argument");
    PHET_CUDA(cudaSetDevice(c->device));
    const double *src = nullptr;
    if (scores == nullptr) {
        PHET_REQUIRE(c->scores.p != nullptr && c->G >= 1, "no scores on the device yet (phet_combine / phet_finalize first)");
        PHET_REQUIRE(G == 0 || G == c->G, "G does not match the resident scores");
        G = c->G;
        src = c->scores.p;
    } else {
        PHET_REQUIRE(G >= 1, "G must be >= 1");
        if (int e = c->rank_in.alloc((size_t)G)) return e;
        PHET_CUDA(cudaMemcpyAsync(c->rank_in.p, scores, sizeof(double) * (size_t)G, cudaMemcpyHostToDevice, c->stream));
        src = c->rank_in.p;
    }
    PHET_REQUIRE(G <= INT32_MAX, "too many genes");
    PHET_REQUIRE(k >= 0 && k <= G, "need 0 <= k <= G");
    if (int e = launch_rank(c, src, G)) return e;
    if (order_out && k) PHET_CUDA(cudaMemcpyAsync(order_out, c->rank_order.p, sizeof(int32_t) * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
    if (sorted_out && k) PHET_CUDA(cudaMemcpyAsync(sorted_out, c->rank_sorted.p, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    return PHET_OK;
}

static int buffer_info(phet_ctx *c, int which, void **p, int64_t *bytes) {
    switch (which) {
        case PHET_BUF_ACC: *p = c->acc.p; *bytes = (int64_t)sizeof(double) * 2 * c->G; break;
        case PHET_BUF_O: *p = c->o.p; *bytes = (int64_t)sizeof(double) * c->G; break;
        case PHET_BUF_H: *p = c->H.p; *bytes = (int64_t)sizeof(int32_t) * c->G * c->n_slices; break;
        case PHET_BUF_P: *p = c->dbgP.p; *bytes = (int64_t)sizeof(double) * c->G * c->S_local; break;
        case PHET_BUF_DELTA: *p = c->dbgD.p; *bytes = (int64_t)sizeof(double) * c->G * c->S_local; break;
        case PHET_BUF_F: *p = c->f.p; *bytes = (int64_t)sizeof(double) * c->G; break;
        case PHET_BUF_R: *p = c->r.p; *bytes = (int64_t)sizeof(double) * c->G; break;
        case PHET_BUF_BINS: *p = c->bins.p; *bytes = c->G; break;
        case PHET_BUF_SCORES: *p = c->scores.p; *bytes = (int64_t)sizeof(double) * c->G; break;
        case PHET_BUF_XG: *p = c->xg.p; *bytes = c->xg_rows * c->n_pad * (c->storage == PHET_F32 ? 4 : 8); break;  // row 0 = gene xg_g0
        case PHET_BUF_MEAN: *p = c->mean.p; *bytes = (int64_t)sizeof(double) * c->G; break;
        case PHET_BUF_STD: *p = c->sd.p; *bytes = (int64_t)sizeof(double) * c->G; break;
        case PHET_BUF_ROWS: *p = c->rows_grouped.p; *bytes = (int64_t)sizeof(int32_t) * (c->n0 + c->n1); break;
        case PHET_BUF_ACC_O: *p = c->acc.p; *bytes = (int64_t)sizeof(double) * 3 * c->G; break;
        case PHET_BUF_RAW: *p = c->raw_keep.p; *bytes = c->raw_N * c->raw_G * (c->raw_dtype == PHET_F32 ? 4 : 8); break;
        case PHET_BUF_STAGED: *p = c->mtx_out.p; *bytes = (int64_t)sizeof(float) * c->mtx_N * c->mtx_G; break;
        case PHET_BUF_RANK_ORDER: *p = c->rank_order.p; *bytes = (int64_t)sizeof(int32_t) * (int64_t)c->rank_order.n; break;
        case PHET_BUF_PROFILE: *p = c->prof.p; *bytes = 16 * (int64_t)sizeof(unsigned long long); break;
        default:
            set_error("invalid: unknown buffer id");
            return PHET_ERR_INVALID;
    }
    if (*p == nullptr) {
        set_error("invalid: buffer has not been produced yet");
        return PHET_ERR_INVALID;
    }
    return PHET_OK;
}

int phet_device_ptr(phet_ctx *c, int which, void **ptr_out, int64_t *bytes_out) {
    PHET_REQUIRE(c != nullptr && ptr_out && bytes_out, "NULL argument");
    return buffer_info(c, which, ptr_out, bytes_out);
}

int phet_read_buffer(phet_ctx *c, int which, void *host_out, int64_t bytes) {
    PHET_REQUIRE(c != nullptr && host_out, "NULL argument");
    void *p = nullptr;
    int64_t have = 0;
    if (int e = buffer_info(c, which, &p, &have)) return e;
    PHET_REQUIRE(bytes <= have, "requested more bytes than the buffer holds");
    PHET_CUDA(cudaSetDevice(c->device));
    PHET_CUDA(cudaMemcpyAsync(host_out, p, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    return PHET_OK;
}

int phet_write_buffer(phet_ctx *c, int which, const void *host_in, int64_t bytes) {
    PHET_REQUIRE(c != nullptr && host_in, "NULL argument");
    void *p = nullptr;
    int64_t have = 0;
    if (int e = buffer_info(c, which, &p, &have)) return e;
    PHET_REQUIRE(bytes <= have, "more bytes than the buffer holds");
    PHET_CUDA(cudaSetDevice(c->device));
    PHET_CUDA(cudaMemcpyAsync(p, host_in, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    return PHET_OK;
}

int phet_set_option(phet_ctx *c, int option, int64_t value) {
    PHET_REQUIRE(c != nullptr, "ctx is NULL");
    switch (option) {
        case PHET_OPT_KEEP_H: c->keep_H = value != 0; return PHET_OK;
        case PHET_OPT_STEP3_MIN_SIZE:
            PHET_REQUIRE(value >= 0, "PHET_OPT_STEP3_MIN_SIZE must be >= 0");
            c->step3_min_size = value;
            return PHET_OK;
        default: set_error("invalid: unknown option"); return PHET_ERR_INVALID;
    }
}

int64_t phet_launch_count(phet_ctx *c) { return c ? c->launches : 0; }

int phet_step3_path(phet_ctx *c) { return c ? c->step3_path : 0; }

int phet_step1_geometry(phet_ctx *c, int32_t out[5]) {
    PHET_REQUIRE(c != nullptr && out, "NULL argument");
    memcpy(out, c->step1_geom, sizeof(c->step1_geom));
    return PHET_OK;
}

// np.linspace(lo, hi, B + 1) restated (numpy function_base.linspace): step = (hi-lo)/B;
// y[i] = i*step + lo (product and sum rounded separately); y[B] = hi.
static void linspace_edges(double lo, double hi, int B, std::vector<double> &e) {
    e.resize((size_t)B + 1);
    const double delta = hi - lo;
    const double step = delta / (double)B;
    for (int i = 0; i <= B; ++i) {
        volatile double prod = (step != 0.0) ? (double)i * step : ((double)i / (double)B) * delta;
        e[(size_t)i] = prod + lo;
    }
    e[(size_t)B] = hi;
}

int phet_fit_host(phet_ctx *c, const void *X, int dtype, int64_t N, int64_t G, const int32_t *ctrl_rows, int64_t n0,
                  const int32_t *case_rows, int64_t n1, const int32_t *idx, int64_t S, const uint32_t *perm_ctrl,
                  const uint32_t *perm_case, const phet_fit_params *prm, double *scores_out,
                  int64_t *bin_counts_out) {
    PHET_REQUIRE(c != nullptr && prm != nullptr && scores_out != nullptr, "NULL argument");
    PHET_REQUIRE(prm->n_bins >= 2 && prm->weights, "need at least two feature weights");  // phet.py:122-123
    if (int e = phet_set_classes(c, ctrl_rows, n0, case_rows, n1)) return e;
    if (int e = phet_set_index_sets(c, idx, S, prm->m)) return e;
    phet_pipeline_params pp;
    memset(&pp, 0, sizeof(pp));
    pp.normalize = prm->normalize;
    pp.storage = prm->storage;
    pp.g_begin = 0;
    pp.g_end = G;
    pp.s_begin = 0;
    pp.s_end = S;
    pp.g3_begin = 0;
    pp.g3_end = G;
    pp.step1 = prm->step1;
    pp.m = prm->m;
    pp.perm_mode = prm->perm_mode;
    pp.seed = prm->seed;
    pp.perm_ctrl = perm_ctrl;
    pp.perm_case = perm_case;
    pp.table_full = prm->table_full;
    pp.table_last = prm->table_last;
    pp.n_last = prm->n_last;
    pp.block_genes = 0;
    if (int e = phet_run_pipeline(c, X, 0, N, G, dtype, &pp)) return e;
    return phet_finalize(c, S, prm->weights, prm->n_bins, scores_out, bin_counts_out, nullptr);
}

// Host evaluations of the device p-value routines (same source, pvalue.cuh) for CPU unit tests.
double phet_host_p_two_sided_t(double t, double df, double ln_beta) { return p_two_sided_t(t, df, ln_beta); }
double phet_host_p_two_sided_z(double z) { return p_two_sided_z(z); }
uint32_t phet_host_perm_apply(uint64_t seed, uint64_t gene, uint32_t cls, uint32_t l, uint32_t i) {
    static thread_local uint32_t cached_n = 0;
    static thread_local std::vector<uint32_t> cp;
    if (cached_n != l) {
        cp.resize(kCoprimeCount);
        coprime_list(l, cp.data());
        cached_n = l;
    }
    return affine_at(make_affine(seed, gene, cls, l, cp.data()), i);
}

}  // extern "C"
