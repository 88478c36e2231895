// Shared declarations for the phet_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/phet_b200.h"

namespace phet {

constexpr int kWarp = 32;
constexpr int kMaxSmemOptin = 232448;
constexpr int kCoprimeCount = 1024;    // per class; power of two  // 227 KB: per-CTA dynamic shared memory limit on sm_100

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define PHET_CUDA(call)                                                           \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return ::phet::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define PHET_REQUIRE(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            ::phet::set_error(std::string("invalid: ") + (msg)); \
            return PHET_ERR_INVALID;                              \
        }                                                         \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;  // elements
    int alloc(size_t count) {
        if (count <= n && p) return 0;
        release();
        if (count == 0) return 0;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            n = 0;
            return cuda_fail(e, "cudaMalloc", __FILE__, __LINE__);
        }
        n = count;
        return 0;
    }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr;
        n = 0;
        owned = true;
    }
    void alias(T *ptr, size_t count) {  // view into another buffer (not freed here)
        release();
        p = ptr;
        n = count;
        owned = false;
    }
    bool owned = true;
    size_t bytes() const { return n * sizeof(T); }
};

}  // namespace phet

namespace phet {
struct ReplayRing;
}

// A parsed MatrixMarket file (mtx.cu): zero-based triplets in file order.
struct phet_mtx {
    int64_t rows = 0, cols = 0, declared = 0;
    bool symmetric = false, pattern = false;
    std::vector<int32_t> r, c;
    std::vector<float> v;
};

// The context.  Plain struct: the C ABI only ever hands out pointers to it.
struct phet_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;   // H2D staging for the pipelined host path
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    phet::DevBuf<unsigned char> stage[2];
    int64_t launches = 0;

    // problem
    int64_t N = 0, G = 0, n_pad = 0;   // cells, genes, padded row length of the gene-major copy
    int64_t xg_g0 = 0, xg_rows = 0;    // genes [xg_g0, xg_g0 + xg_rows) are resident in xg
    int64_t n0 = 0, n1 = 0;            // class sizes
    int storage = PHET_F32;
    bool classes_set = false, matrix_set = false, idx_set = false;
    int64_t S = 0, m = 0;              // uploaded index sets
    int64_t S_local = 0;               // subsamples covered by the debug buffers
    int64_t n_slices = 0;
    int64_t coprime_n0 = -1, coprime_n1 = -1;
    int32_t n_bins = 0;

    phet::DevBuf<int32_t> rows_grouped;  // [N]: source row of grouped position q (controls, then cases)
    phet::DevBuf<int32_t> pos_of_row;    // [N]: grouped position of original row
    phet::DevBuf<int32_t> cls_pos;       // [N]: class-relative grouped position of class-local index t (ctrl, then case)
    phet::DevBuf<uint32_t> coprime;      // [2][kCoprimeCount]: affine multipliers coprime to n0 / n1
    phet::DevBuf<unsigned char> xg;      // [G][n_pad] storage dtype
    phet::DevBuf<double> mean, sd;       // [G]
    phet::DevBuf<double> colpart;        // scratch for column statistics
    phet::DevBuf<unsigned char> xraw;    // staging copy of a host matrix
    phet::DevBuf<int32_t> idx;           // [S][2][m] grouped positions
    phet::DevBuf<uint16_t> idxT;         // [2][m][S] class-local positions (pair-per-thread Step 1)
    bool idxT_ok = false;
    phet::DevBuf<uint32_t> idxT32;       // same table with 32-bit positions in ascending order (k_step1p, BIG variant)
    bool idxT32_ok = false;
    bool pair_force_big = false;         // env PHET_STEP1_FORCE_BIG=1: BIG variant even where staging applies (tests)
    int pair_big_threads = 256;          // threads per CTA of the BIG variant (env PHET_STEP1_BIG_THREADS): what is
                                         // left of the 256 KB L1/SMEM array is the L1 that serves the gathers --
                                         // config 5: 512 threads (205 KB SMEM) 175 ms, 384: 73 ms, 256-352: 64-65 ms
    int step1_impl = 0;                  // 0 auto, 1 warp sorters only (env PHET_STEP1_IMPL=warp)
    bool fast_moments = false;           // float storage of z-scored data: fp32 chunk sums in k_step1p (set by prepare_matrix)
    int pair_nbw = 0;                    // 0 auto, 16 / 32: forced histogram words (env PHET_STEP1_NBW, tests)
    phet::DevBuf<float2> win;            // [G][2] bucket map (scale, offset) per (gene, class)
    phet::DevBuf<double> pair_scratch;   // [grid][3][slots] control-class results awaiting the case sample
    phet::DevBuf<unsigned long long> prof;  // [16] optional per-phase cycle counters (env PHET_PROFILE=1)
    phet::DevBuf<float2> win3;           // [G] common bucket map of a gene (bound-and-prune Step 3)
    bool keep_H = false;                 // phet_set_option(PHET_OPT_KEEP_H): per-slice KS statistics wanted
    int step3_impl = 0;                  // 0 auto, 1 sorting kernels only (env PHET_STEP3_IMPL=sort)
    int step3_path = 0;                  // last Step-3 launch: 1 full-warp sort, 2 half-warp sort, 3 bound-and-prune
    int tab_h0 = -1;                     // table_full is non-increasing from here on (host-checked); -1: unknown
    phet::DevBuf<double> acc;            // [3][G]: sum -2 ln p, sum delta, and o (aliased below): one all-reduce covers all three
    std::vector<double> tab_full_host, tab_last_host;  // last uploaded KS tables (re-uploaded only when they change)
    std::vector<double> weights_host;                  // last uploaded feature weights (phet_finalize_async)
    phet::DevBuf<double> dbgP, dbgD;     // [G][S_local]
    phet::DevBuf<double> o;              // [G]
    phet::DevBuf<int32_t> H;             // [G][n_slices]
    phet::DevBuf<uint32_t> perm0, perm1; // [genes from perm_g0][min_c] (host mode: all G genes; replay mode: one block)
    int64_t perm_g0 = 0;
    phet::DevBuf<uint16_t> slice_of;       // inverse view of the same tables: [perm_rows][n0 + n1] slice of every cell (scatter Step 3)
    bool slice_of_valid = false, perm_fwd_valid = false;   // which views of the tables are filled
    phet::DevBuf<uint32_t> coprime_inv;    // [2][kCoprimeCount]: inverses of the affine multipliers mod n0 / n1
    int64_t perm_rows = 0, perm_cols = 0;  // resident tables: genes [perm_g0, perm_g0 + perm_rows) x perm_cols positions (0: none)
    cudaEvent_t ev_phase[4] = {nullptr, nullptr, nullptr, nullptr};
    bool phases_valid = false;
    phet::DevBuf<double> tab_full, tab_last;
    phet::DevBuf<double> f, r, fw, scores;   // [G]
    phet::DevBuf<signed char> bins;      // [G]
    phet::DevBuf<double> scalars;        // small scratch: sums, min/max
    phet::DevBuf<double> edges, weights;
    phet::DevBuf<unsigned long long> counts;
    phet::DevBuf<double> class_stats;    // [7][G] (phet_class_stats)

    // MatrixMarket ingestion (mtx.cu): triplets, the dense matrix as read, what the filter keeps
    phet::DevBuf<float> mtx_dense, mtx_out, mtx_v, mtx_rowsum;
    phet::DevBuf<int32_t> mtx_r, mtx_c, mtx_prog, mtx_rows, mtx_cols, mtx_cnt;
    std::vector<int32_t> mtx_rows_host, mtx_cols_host;   // kept cells / genes, ascending (train.py:92,111)
    int64_t mtx_N = 0, mtx_G = 0;                        // shape of mtx_out
    phet::DevBuf<unsigned char> raw_keep;   // phet_stage_matrix: the caller's matrix as given, kept for repeated layouts
    int64_t raw_N = 0, raw_G = 0;
    int raw_dtype = PHET_F32;
    // multi-class Step 1 (anova.cu) and the Step-3 slicing range of a multi-class fit
    phet::DevBuf<int32_t> anova_idx;
    phet::DevBuf<uint32_t> anova_idxT, anova_idxU;
    phet::DevBuf<int> cs_ovf;            // [1 + G]: count | genes the bucket-selection launch of phet_class_stats hands to the general one
    int64_t step3_min_size = 0;          // phet_set_option(PHET_OPT_STEP3_MIN_SIZE): slices cover [0, this) instead of [0, min(n0, n1))
    // ranking (rank.cu)
    phet::DevBuf<unsigned long long> rank_keys;
    phet::DevBuf<int32_t> rank_order;
    phet::DevBuf<double> rank_sorted, rank_in;

    int32_t step1_geom[5] = {0, 0, 0, 0, 0};

    // replay of the reference's np.random draws (mt_replay.cpp -> fy_apply.cu, replay.cu)
    phet::DevBuf<int32_t> cls_rows;        // [N]: original row of class-local index t (controls, then cases)
    phet::DevBuf<uint32_t> fy_scratch;     // global-memory Fisher-Yates rows (classes longer than 65,536 cells)
    phet::DevBuf<unsigned char> draws_dev; // one block of swap targets as uploaded
    phet::ReplayRing *ring = nullptr;     // pinned host slots + producer thread state (replay.cu)
    double replay_host_s = 0.0;            // host seconds the last replay spent drawing (producer thread)
    double replay_wait_s = 0.0;            // seconds the issuing thread waited for draws
};

namespace phet {

// kernels' host launchers (defined in the .cu files)
int launch_normalize(phet_ctx *c, const void *Xblock, int64_t ldx, int64_t g0, int64_t gcount, int dtype,
                     int normalize);
int launch_step1(phet_ctx *c, int64_t g_begin, int64_t g_end, int64_t s_begin, int64_t s_end,
                 const phet_step1_params *prm, int accumulate);
int launch_step3(phet_ctx *c, int64_t g_begin, int64_t g_end, int64_t m, int perm_mode, uint64_t seed,
                 int64_t n_last);
int launch_fisher(phet_ctx *c, int64_t S_total);
int launch_class_stats(phet_ctx *c, int64_t g_begin, int64_t g_end, const phet_quantile_pos *q3, double *out_dev);
int launch_o_minmax(phet_ctx *c, double out[2]);
int launch_bin(phet_ctx *c, int32_t B, int64_t *counts_out);
int launch_combine(phet_ctx *c, int32_t B, double *scores_out);
int launch_finalize(phet_ctx *c, int64_t S_total, int32_t B);
int launch_rank(phet_ctx *c, const double *scores_dev, int64_t G);
int launch_fy_apply(phet_ctx *c, const void *draws_dev, int elem_bytes, int64_t ld, int64_t n, int64_t count, int64_t take,
                    const int32_t *map_dev, uint32_t *out_dev, int64_t out_ld, cudaStream_t stream);
// the same shuffles written as the INVERSE view: slice_out[p * out_ld + map[x[q]]] = q < take ? min(q / m, last_slice) : 0xFFFF
int launch_fy_apply_inverse(phet_ctx *c, const void *draws_dev, int elem_bytes, int64_t ld, int64_t n, int64_t count, int64_t take,
                            const int32_t *map_dev, uint16_t *slice_out, int64_t out_ld, int64_t m, int64_t last_slice, cudaStream_t stream);
bool step3_wants_inverse(phet_ctx *c, int64_t m, int64_t n_last);   // the scatter kernel will serve Step 3 of this fit
// shared by api.cu and replay.cu
int prepare_matrix(phet_ctx *c, int64_t N, int64_t G, int dtype, int normalize, int storage, int64_t g_begin, int64_t g_end);
int step3_prepare(phet_ctx *c, int64_t m, int perm_mode, const uint32_t *perm_ctrl, const uint32_t *perm_case,
                  const double *table_full, const double *table_last, int64_t n_last);
int finish_index_sets(phet_ctx *c, int64_t S, int64_t m);   // c->idx holds ORIGINAL rows: remap, build the transposed tables
int run_pipeline_replay(phet_ctx *c, const void *X, int is_device, int64_t N, int64_t G, int dtype,
                        const phet_pipeline_params *p);
void replay_ring_destroy(phet_ctx *c);
int phase_mark(phet_ctx *c, int k);   // record phase event k on the stream (creates the events on first use)
int launch_anova(phet_ctx *c, const int32_t *idx_host, int64_t S, int64_t K, int64_t m, double ln_beta, double p_flush,
                 int capture_debug, int kruskal);
int mtx_parse(const char *path, int threads, phet_mtx &m);
int mtx_stage(phet_ctx *c, const phet_mtx &m, int do_filter, int64_t minimum_samples, int64_t dims_out[2]);

}  // namespace phet
