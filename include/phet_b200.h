/*
 * phet_b200.h -- C ABI of the B200-native PHet fit_predict hot path.
 *
 * The reference (kleelab-bch/phet) is pure Python and has no FFI layer; its de-facto operator
 * interface for this path is the model-class convention `PHet(...).fit_predict(X, y)`
 * (src/model/phet.py:33-40, 163-518) called from src/train.py:474-477.  The entry points
 * below are the device boundary a binding for that method needs: each one replaces the span
 * of phet.py cited beside it.  The Python mirror of the class lives in phet_b200/phet.py and
 * binds these with ctypes (phet_b200/_capi.py); INTEGRATION.md shows the stub a maintainer
 * of the reference would add.
 *
 * Conventions: every function returns 0 on success or a negative phet_status; a human-readable
 * message for the calling thread's last failure is available from phet_last_error().  No C++
 * exceptions cross the boundary.  Host buffers are caller-allocated and caller-owned.  A
 * context is bound to one CUDA device and one stream and is not thread-safe.  All kernels are
 * sm_100a; there is no CPU fallback -- phet_ctx_create fails if no such device is present.
 */
#ifndef PHET_B200_H
#define PHET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHET_B200_ABI_VERSION 1

typedef struct phet_ctx phet_ctx;

typedef enum {
    PHET_OK = 0,
    PHET_ERR_INVALID = -1,   /* bad argument / call order                          */
    PHET_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError name) */
    PHET_ERR_UNSUPPORTED = -3, /* size outside what the kernels are built for      */
    PHET_ERR_NOMEM = -4
} phet_status;

typedef enum { PHET_F32 = 0, PHET_F64 = 1 } phet_dtype;
typedef enum { PHET_NORM_NONE = 0, PHET_NORM_ZSCORE = 1 } phet_normalize;
typedef enum { PHET_TEST_T = 0, PHET_TEST_Z = 1 } phet_test;          /* phet.py:425-428 */
/* PHET_PERM_RESIDENT: reuse the permutation tables a previous call left in HBM (a PHET_PERM_HOST upload, or a
 * PHET_PERM_REPLAY pipeline run with keep_perm_tables) -- the "inputs already resident" case of a benchmark. */
typedef enum { PHET_PERM_HOST = 0, PHET_PERM_DEVICE = 1, PHET_PERM_REPLAY = 2, PHET_PERM_RESIDENT = 3 } phet_perm_mode;
typedef struct phet_mt phet_mt;  /* NumPy's legacy global stream, held natively (see below) */

/* Device buffers a caller may want to wrap (e.g. for an NCCL all-reduce).  phet_device_ptr. */
typedef enum {
    PHET_BUF_ACC = 0,      /* double[2][G]: sum_s -2 ln p  then  sum_s delta-IQR  (phet.py:417,432,449-451) */
    PHET_BUF_O = 1,        /* double[G]: min-over-slices KS p-value (phet.py:496-498)                        */
    PHET_BUF_H = 2,        /* int32[G][n_slices]: integer KS statistics (debug / parity)                     */
    PHET_BUF_P = 3,        /* double[G][S_local]: p-values, only when debug capture is on                    */
    PHET_BUF_DELTA = 4,    /* double[G][S_local]: delta-IQR, only when debug capture is on                   */
    PHET_BUF_F = 5,        /* double[G]: standardised Fisher statistic (phet.py:458-462)                     */
    PHET_BUF_R = 6,        /* double[G]: mean delta-IQR (phet.py:435-444)                                    */
    PHET_BUF_BINS = 7,     /* int8[G]: ordinal bin of o (phet.py:500-504)                                    */
    PHET_BUF_SCORES = 8,   /* double[G]: final scores (phet.py:510-516)                                      */
    PHET_BUF_XG = 9,       /* normalised gene-major matrix, storage dtype, [G][n_pad], class-grouped cells   */
    PHET_BUF_MEAN = 10,    /* double[G] column mean  */
    PHET_BUF_STD = 11,     /* double[G] column population sd */
    PHET_BUF_ROWS = 12,    /* int32[N]: original row held at each position of a gene-major row (the layout) */
    PHET_BUF_ACC_O = 14,   /* double[3][G]: PHET_BUF_ACC followed by PHET_BUF_O, contiguous (one all-reduce for both)    */
    PHET_BUF_STAGED = 15,  /* float32 [cells kept][genes kept], row-major: the matrix phet_mtx_stage left in HBM */
    PHET_BUF_RANK_ORDER = 16, /* int32[G]: gene index at each rank, descending score (phet_rank)                        */
    PHET_BUF_RAW = 17,     /* the caller's matrix as given to phet_stage_matrix (row-major cells x genes)                 */
    PHET_BUF_PROFILE = 13  /* uint64[16]: per-phase cycle counters of the Step-3 kernel, only with env PHET_PROFILE=1 */
} phet_buffer;

/* Order-statistic positions and lerp weights of the two percentiles, computed on the host
 * with the index arithmetic of scipy.stats.iqr -> stats.quantile(method="linear")
 * (called at phet.py:413-414): jg = p*m + (1-p); j = floor(jg)-1; g = frac(jg). */
typedef struct {
    int32_t j_lo, k_lo;   /* sorted positions y[j], y[j+1] (clipped to [0, m-1]) for the low percentile */
    int32_t j_hi, k_hi;
    double g_lo, g_hi;    /* q = (1-g)*y[j] + g*y[k] */
} phet_quantile_pos;

typedef struct {
    int32_t test;             /* phet_test: Student pooled t (m < 30) or pooled z (phet.py:425-428)     */
    double ln_beta;           /* PHET_TEST_T: ln B(df/2, 1/2), df = 2m-2 (host lgamma)                    */
    double p_flush_below;     /* p-values <= this become 0 (float32 stdtr underflow emulation; 0 = off)  */
    phet_quantile_pos q;
    int32_t capture_debug;    /* 1: also write PHET_BUF_P / PHET_BUF_DELTA                               */
} phet_step1_params;

int phet_abi_version(void);
const char *phet_last_error(void);

/* Number of CUDA devices visible, or a negative status. */
int phet_device_count(void);

/* Create a context on `device`.  `stream` is a cudaStream_t owned by the caller (e.g. torch's
 * current stream) or NULL to let the context create and own one. */
int phet_ctx_create(int device, void *stream, phet_ctx **out);
int phet_ctx_destroy(phet_ctx *ctx);
int phet_ctx_sync(phet_ctx *ctx);

/* Class membership (phet.py:388-392): row numbers of control (y == 0) and case (y == 1) cells,
 * each ascending.  Must precede phet_load_matrix: the device layout groups cells by class. */
int phet_set_classes(phet_ctx *ctx, const int32_t *ctrl_rows, int64_t n0,
                     const int32_t *case_rows, int64_t n1);

/* Replaces phet.py:291,299 (zscore + nan_to_num) and builds the HBM layout.
 * X: row-major cells x genes, N x G, element type `dtype`, in host memory (is_device = 0,
 * pageable or pinned) or device memory on the context's device (is_device = 1; not modified).
 * storage: element type of the normalised gene-major copy kept in HBM and staged to SMEM. */
int phet_load_matrix(phet_ctx *ctx, const void *X, int is_device, int64_t N, int64_t G,
                     int dtype, int normalize, int storage);

/* The same for a gene range only (a rank of a gene-sharded fit): X_dev is a DEVICE pointer to element (0, g_begin)
 * of a row-major matrix with leading dimension ldx (>= g_end - g_begin; a rank may hold just its columns).
 * Genes [g_begin, g_end) become resident; G stays the gene count of the whole problem, so accumulators, gene
 * numbering and the keyed Step-3 permutations are those of the unsharded fit. */
int phet_load_matrix_genes(phet_ctx *ctx, const void *X_dev, int64_t ldx, int64_t N, int64_t G, int dtype,
                           int normalize, int storage, int64_t g_begin, int64_t g_end);

/* Copy a host matrix (row-major cells x genes, `dtype`) to HBM once, unchanged: PHET_BUF_RAW.  For callers that
 * lay the same matrix out several times with phet_load_matrix(is_device = 1) -- the multi-class fit regroups the
 * cells once per combination (phet.py:387-392) and should cross PCIe once. */
int phet_stage_matrix(phet_ctx *ctx, const void *X, int64_t N, int64_t G, int dtype);

/* Subsample index sets (phet.py:401-402), host memory, int32[S][2][m]: for each subsample the
 * control draw then the case draw, ORIGINAL row numbers (duplicates allowed: replace=True). */
int phet_set_index_sets(phet_ctx *ctx, const int32_t *idx, int64_t S, int64_t m);

/* Replaces the loop phet.py:397-432 for subsamples [s_begin, s_end) over all genes, and
 * accumulates PHET_BUF_ACC (zeroed first unless accumulate != 0). */
int phet_step1(phet_ctx *ctx, int64_t s_begin, int64_t s_end, const phet_step1_params *prm,
               int accumulate);

/* Replaces phet.py:473-498 for genes [g_begin, g_end): per gene, permute each class column,
 * cut slices of length m over [0, min(n0,n1)), two-sample KS per slice, o = min p.
 * perm_mode HOST: perm_ctrl / perm_case are host uint32 [G][min_c] class-local positions (the
 *   first min_c entries of np.random.permutation(n_class), phet.py:486-487) for ALL genes.
 * perm_mode DEVICE: the cells of each class sit in a fixed pseudo-random order in HBM (see
 *   PHET_BUF_ROWS) and gene g walks them with its own affine bijection t -> (a_g t + b_g) mod n
 *   keyed by `seed`; the pointers are ignored.  Pseudo-random slice membership per gene, not
 *   the reference's MT19937 stream.
 * table_full[0..m] / table_last[0..n_last]: exact two-sided p-value of ks_2samp for equal
 * sizes as a function of the integer statistic h (host-built, scipy semantics). */
int phet_step3(phet_ctx *ctx, int64_t g_begin, int64_t g_end, int64_t m, int perm_mode,
               const uint32_t *perm_ctrl, const uint32_t *perm_case, uint64_t seed,
               const double *table_full, const double *table_last, int64_t n_last);

/* Stage 0 + Step 1 + Step 3 fused over gene blocks (the end-to-end path).  For a host matrix
 * (is_device = 0; pinned memory recommended) block b+1 is copied H2D on a second stream while
 * block b is normalised and processed, so the fit is bounded by max(PCIe, compute) instead of
 * their sum.  Requires phet_set_classes and phet_set_index_sets.  Genes [g_begin, g_end) are
 * loaded (only they become resident) and run through Step 1 for subsamples [s_begin, s_end);
 * Step 3 runs for genes [g3_begin, g3_end), a sub-range of the loaded genes.  Single GPU:
 * all ranges full.  Subsample sharding (north star): all genes, a slice of s, a slice of g3.
 * Gene sharding (capacity layout): a slice of genes, all s, g3 = the same slice. */
typedef struct {
    int32_t normalize, storage;
    int64_t g_begin, g_end;
    int64_t s_begin, s_end;
    int64_t g3_begin, g3_end;
    phet_step1_params step1;
    int64_t m;
    int32_t perm_mode;
    uint64_t seed;
    const uint32_t *perm_ctrl, *perm_case; /* PHET_PERM_HOST: uint32 [G][min_c] for all genes */
    const double *table_full, *table_last;
    int64_t n_last;
    int64_t block_genes;                   /* genes per H2D block; 0 = auto */
    /* PHET_PERM_REPLAY: the Step-3 permutations are the ones the reference would draw from `mt` (phet.py:486-487),
     * for ALL G genes in gene order (a rank that owns a gene range consumes the other genes' draws without
     * storing them, so every rank leaves `mt` in the reference's final state).  The draws are made on a host
     * thread while the device works on earlier blocks; the swaps run on the device.
     * draw_index_sets != 0: the S x 2 index sets are drawn from `mt` first (phet.py:397-402, replace = False;
     * requires m <= min(n0, n1)) instead of coming from phet_set_index_sets; S_draw / m give their shape and
     * idx_out (host int32[S][2][m], may be NULL) receives them as ORIGINAL row numbers. */
    phet_mt *mt;
    int32_t draw_index_sets;
    int64_t S_draw;
    int32_t *idx_out;
    int32_t keep_perm_tables;              /* PHET_PERM_REPLAY: keep the tables of genes [g3_begin, g3_end) for PHET_PERM_RESIDENT */
    /* PHET_PERM_REPLAY with several ranks on one host: != 0 stops the replay after the draws of gene g3_end - 1 instead
     * of running through the remaining genes' draws.  `mt` is then left mid-stream; the rank that owns the LAST genes
     * (g3_end == G) ends in the reference's final state and hands it to the others (phet_b200/phet.py broadcasts it).
     * The scan of the stream is sequential, so rank r pays (r + 1) / N of it instead of all of it, and the host's cores
     * free up as the early ranks finish. */
    int32_t stop_after_own_genes;
} phet_pipeline_params;

int phet_run_pipeline(phet_ctx *ctx, const void *X, int is_device, int64_t N, int64_t G, int dtype,
                      const phet_pipeline_params *prm);

/* Replaces phet.py:435-444 and 449-462: r = acc_r / S_total, f = standardised Fisher. */
int phet_fisher(phet_ctx *ctx, int64_t S_total);

/* min and max of o over all genes (inputs of np.linspace at sklearn _discretization.py:339). */
int phet_o_minmax(phet_ctx *ctx, double out_minmax[2]);

/* Replaces phet.py:500-504.  edges: host double[B+1] (np.linspace(min o, max o, B+1), or
 * {-inf, +inf} with B = 1 when o is constant).  counts_out: host int64[B] bin counts. */
int phet_bin(phet_ctx *ctx, const double *edges, int32_t B, int64_t *counts_out);

/* Replaces phet.py:510-516.  weights: host double[B], already normalised (phet.py:124).
 * scores_out: host double[G] or NULL (leave on device, PHET_BUF_SCORES). */
int phet_combine(phet_ctx *ctx, const double *weights, int32_t B, double *scores_out);

/* phet_fisher + phet_o_minmax + np.linspace edges + phet_bin + phet_combine in one call with no host round
 * trip in between (the edges are formed on the device with numpy's operation order, so bins are the same as
 * with host-supplied edges).  Replaces phet.py:435-462 and 500-516.  weights: host double[B], normalised.
 * scores_out: host double[G]; counts_out: host int64[B]; edges_out: host double[B+1]; each may be NULL. */
int phet_finalize(phet_ctx *ctx, int64_t S_total, const double *weights, int32_t B, double *scores_out,
                  int64_t *counts_out, double *edges_out);

/* The same work queued on the context's stream WITHOUT a host synchronisation: scores stay in PHET_BUF_SCORES,
 * counts and edges on the device (the weights are cached on the device and re-uploaded only when they change).
 * A caller that runs fit after fit -- or follows the fit with a collective on the same stream -- keeps the GPU
 * busy back to back; phet_fetch_results copies the last results out and synchronises.  Replaces the same spans
 * as phet_finalize (phet.py:435-462, 500-516).  Each *_out may be NULL. */
int phet_finalize_async(phet_ctx *ctx, int64_t S_total, const double *weights, int32_t B);
int phet_fetch_results(phet_ctx *ctx, double *scores_out, int64_t *counts_out, double *edges_out);

/* Context options.  PHET_OPT_KEEP_H (default 0): 1 = phet_step3 / phet_run_pipeline also fill PHET_BUF_H
 * with the integer KS statistic of EVERY slice (parity / debugging; the sorting kernels run); 0 = only
 * o[g] = min p is produced, which lets Step 3 bound h per slice from bucket histograms and evaluate
 * exactly only the slices that can attain the gene's minimum p (phet.py:496 takes the min). */
/* PHET_OPT_STEP3_MIN_SIZE (default 0 = min(n0, n1)): the Step-3 slices cover positions [0, value) of each permuted
 * class column.  A multi-class fit slices every (class i, other cells) pair over the smallest of ALL classes
 * (phet.py:466-471, min_size), which is smaller than min(n0, n1) of the pair; host permutations are then
 * uint32 [G][value]. */
typedef enum { PHET_OPT_KEEP_H = 1, PHET_OPT_STEP3_MIN_SIZE = 2 } phet_option;
int phet_set_option(phet_ctx *ctx, int option, int64_t value);

/* Raw device pointer / element count of a context buffer (for wrapping, no ownership change). */
int phet_device_ptr(phet_ctx *ctx, int which, void **ptr_out, int64_t *bytes_out);
/* Copy a context buffer to host memory (synchronises the stream). */
int phet_read_buffer(phet_ctx *ctx, int which, void *host_out, int64_t bytes);
/* Overwrite a context buffer from host memory (e.g. externally reduced accumulators). */
int phet_write_buffer(phet_ctx *ctx, int which, const void *host_in, int64_t bytes);

/* Introspection for benchmarks: number of kernel launches issued by this context so far,
 * and the launch geometry chosen for the step-1 kernel (grid, block, dynamic smem bytes,
 * elements per lane E, staged = gene vector held in shared memory). */
int64_t phet_launch_count(phet_ctx *ctx);
int phet_step1_geometry(phet_ctx *ctx, int32_t out[5]);
/* Kernel family of the last Step-3 launch: 1 full-warp sort, 2 half-warp sort, 3 bound-and-prune. */
int phet_step3_path(phet_ctx *ctx);

/* Whole-vector statistics of every resident gene, for the reference's sibling IQR scorers
 * (replaces the per-gene Python loops of src/model/hvf.py:86-93 [HIQR] and
 * src/model/deltaiqrmean.py:53-85 [DeltaIQRMean]: scipy.stats.iqr, np.mean and the moments that
 * scipy.stats.ttest_ind needs, over ALL cells of a class / of the data set).
 * q3[0], q3[1], q3[2]: sorted positions of the two percentiles for n0, n1 and n0+n1 elements
 * (scipy's 'linear' rule, as phet_step1_params.q).  out: host double[7][G] =
 * iqr(class 0), iqr(class 1), iqr(all cells), mean(class 0), mean(class 1),
 * sum of squared deviations (class 0), (class 1).  Order statistics are exact (radix select). */
int phet_class_stats(phet_ctx *ctx, const phet_quantile_pos *q3, double *out);

/* Multi-class Step 1: replaces phet.py:348-382 with group_test="anova" -- for each of S
 * subsamples, scipy.stats.f_oneway over K groups of m cells per gene, nan_to_num(p), and Fisher's sum
 * sum_s -2 ln p (p == 0 dropped, phet.py:449-451), written to PHET_BUF_ACC[0 .. G) (and p to PHET_BUF_P when
 * capture_debug != 0).  idx: host int32[S][K][m], ORIGINAL row numbers (any grouping of the cells into the two
 * phet_set_classes lists works: only the layout depends on it).  ln_beta = ln B((K m - K)/2, (K - 1)/2) (host
 * lgamma).  p_flush_below: p-values <= this become 0 (float32 fdtrc underflow for float32 input; 0 = off). */
int phet_anova(phet_ctx *ctx, const int32_t *idx, int64_t S, int64_t K, int64_t m, double ln_beta,
               double p_flush_below, int capture_debug);

/* The same with group_test="kruskal" (phet.py:373-375, taken when m > 5): scipy.stats.kruskal over the K groups --
 * average ranks of the pooled K m values (exact, by counting), H with the tie correction, p = chi2.sf(H, K - 1).
 * K <= 64, K m <= 3072 (float storage) / 1536 (double storage). */
int phet_kruskal(phet_ctx *ctx, const int32_t *idx, int64_t S, int64_t K, int64_t m, double p_flush_below,
                 int capture_debug);

/* Input path.  phet_mtx_parse replaces src/train.py:82-83 (sc.read_mtx: scipy.io.mmread,
 * float32): a multi-threaded host parse of a MatrixMarket `coordinate` file (real | integer | pattern, general |
 * symmetric) into zero-based triplets; values are rounded text -> double -> float like mmread(...).astype("float32").
 * threads <= 0: one per hardware thread.  dims_out: rows (cells), columns (genes), entries read.  No GPU needed. */
typedef struct phet_mtx phet_mtx;
int phet_mtx_parse(const char *path, int threads, phet_mtx **out, int64_t dims_out[3]);
/* Borrowed views of the parsed triplets (valid until phet_mtx_free); any out pointer may be NULL. */
int phet_mtx_triplets(const phet_mtx *m, const int32_t **rows, const int32_t **cols, const float **vals,
                      int64_t *count, int *symmetric);
int phet_mtx_free(phet_mtx *m);
/* Uploads the triplets (12 bytes per entry instead of the dense matrix), scatters them into a dense float32
 * cells x genes matrix in HBM (duplicates add up, as coo -> csr does), zeroes NaN / +-inf cells (np.nan_to_num,
 * train.py:85) and, with do_filter != 0, applies train.py:88-118 on the device: cells with sum|x| >= 5 are kept
 * (row sums in NumPy's pairwise fp32 order, so the decision is the reference's), then genes whose counts per
 * million exceed 1 in at least `minimum_samples` kept cells (train.py's local default is 5; lowered to
 * cells/2 under the reference's rule, train.py:108-109).  dims_out: cells kept, genes kept.  The result is
 * PHET_BUF_STAGED, to be handed to phet_load_matrix / phet_run_pipeline with is_device = 1. */
int phet_mtx_stage(phet_ctx *ctx, const phet_mtx *m, int do_filter, int64_t minimum_samples, int64_t dims_out[2]);
/* Ascending row / column numbers the filter kept (examples_ids, feature_ids of train.py:92,111): the caller
 * filters y and the feature names with them.  rows_out: int32[cells kept], cols_out: int32[genes kept]. */
int phet_mtx_kept(phet_ctx *ctx, int32_t *rows_out, int32_t *cols_out);

/* Ranking: replaces the ordering of src/utility/utils.py:107-136 (sort_features:
 * df.sort_values(by=["score"], ascending=False)).  scores: host double[G], or NULL to rank the resident
 * PHET_BUF_SCORES (G may then be 0).  The first k entries are returned: order_out[r] = gene index at rank r,
 * sorted_out[r] = its score (either may be NULL); the whole order stays in PHET_BUF_RANK_ORDER.  Descending;
 * equal scores keep ascending gene order (pandas kind="stable"; the reference's default quicksort leaves the
 * order inside a tie group unspecified); NaN last. */
int phet_rank(phet_ctx *ctx, const double *scores, int64_t G, int64_t k, int32_t *order_out, double *sorted_out);

/* One-call convenience: everything above on one device with host buffers (used by C callers
 * and by the end-to-end benchmark leg).  perm_mode/seed as in phet_step3; with PHET_PERM_HOST
 * perm_ctrl/perm_case must be given.  scores_out: host double[G]. */
typedef struct {
    int32_t normalize, storage;
    phet_step1_params step1;
    int64_t m;
    int32_t perm_mode;
    uint64_t seed;
    int32_t n_bins;
    const double *weights;      /* double[n_bins] normalised */
    const double *table_full;   /* double[m+1] */
    const double *table_last;   /* double[n_last+1] */
    int64_t n_last;
} phet_fit_params;

int phet_fit_host(phet_ctx *ctx, const void *X, int dtype, int64_t N, int64_t G,
                  const int32_t *ctrl_rows, int64_t n0, const int32_t *case_rows, int64_t n1,
                  const int32_t *idx, int64_t S, const uint32_t *perm_ctrl,
                  const uint32_t *perm_case, const phet_fit_params *prm,
                  double *scores_out, int64_t *bin_counts_out);

/* ---- Replay of NumPy's legacy global RNG (csrc/mt_replay.cpp; host code, no GPU needed) -------------------------
 * The reference draws from the process-global np.random stream (MT19937, seeded once at src/train.py:26):
 * np.random.choice(rows, m, replace) twice per subsample (phet.py:401-402), then np.random.permutation of both
 * class columns per gene (phet.py:486-487).  A phet_mt holds that stream natively: created from
 * np.random.get_state()[1:3] (key[624], pos), advanced by the calls below exactly as NumPy would advance it, and
 * read back with phet_mt_state for np.random.set_state.
 *
 * np.random.permutation(n) is `x = arange(n); for i = n-1 .. 1: j = rk_interval(i); swap(x[i], x[j])` where
 * rk_interval(i) is the first stream value, masked to the bits of i, that is <= i.  phet_mt_shuffle_rounds emits the
 * swap targets only (the sequential part, vectorised); applying them is independent per permutation and is done on
 * the device (phet_draw_index_sets, PHET_PERM_REPLAY) or by phet_host_fisher_yates. */
int phet_mt_create(const uint32_t *key624, int32_t pos, phet_mt **out);
int phet_mt_state(const phet_mt *mt, uint32_t *key624_out, int32_t *pos_out);
int phet_mt_destroy(phet_mt *mt);
/* 512 when the AVX-512 scan is in use, 0 for the scalar one (mt may be NULL: what a new generator would use). */
int phet_mt_simd(const phet_mt *mt);
/* One job = one np.random.permutation(n) per round: its n - 1 swap targets, in draw order (t = 0 .. n-2 is the
 * target of position i = n - 1 - t), are written to (char *)out + round * stride_bytes as 2-byte (n <= 65536) or
 * 4-byte integers; out == NULL consumes the stream without storing (genes another rank owns).  Rounds are the
 * outer loop, jobs the inner one -- {n0, n1} x S rounds is phet.py:397-402, {n0, n1} x G rounds is phet.py:484-487. */
typedef struct {
    int64_t n;
    void *out;
    int64_t stride_bytes;
    int32_t elem_bytes;
} phet_mt_job;
int phet_mt_shuffle_rounds(phet_mt *mt, const phet_mt_job *jobs, int32_t n_jobs, int64_t rounds);
/* The same draws with the work spread over threads.  Only WHERE each permutation starts in the stream is sequential
 * by nature, so: generator threads (env PHET_REPLAY_GENERATORS, default 4) each walk the whole stream with the twist
 * alone and temper + store every G-th region into a cache-resident ring (16 bits per output when max_n <= 65536); ONE
 * scanner thread walks the permutations with a count-only scan of the ring and records where each starts; `workers`
 * threads (<= 0: env PHET_REPLAY_WORKERS or 4) redo the scan from there, storing the targets.  With too few cores, or
 * max_n < 2048, the scanner generates for itself and the workers regenerate from recorded generator states.
 * Submit blocks of rounds in stream order (returns at once; outputs must stay valid), wait on the ticket of a block
 * before reading its outputs, end with phet_mt_par_end, which writes the stream's final state back into `mt` (do not
 * touch `mt` in between).  max_n: longest permutation that will be submitted. */
typedef struct phet_mt_par phet_mt_par;
int phet_mt_par_begin(phet_mt *mt, int32_t workers, int64_t max_n, phet_mt_par **out);
int phet_mt_par_threads(const phet_mt_par *par, int32_t out[3]);   /* generator, scanner, worker threads in use */
int phet_mt_par_submit(phet_mt_par *par, const phet_mt_job *jobs, int32_t n_jobs, int64_t rounds, int64_t *ticket_out);
int phet_mt_par_wait(phet_mt_par *par, int64_t ticket);
int phet_mt_par_end(phet_mt_par *par);
/* count draws of rk_interval(max): what np.random.randint(0, max + 1, count) consumes (choice(..., replace=True)). */
int phet_mt_bounded(phet_mt *mt, uint32_t max, int64_t count, uint32_t *out);
/* The first `take` entries of the permutation of arange(n) that the draws produce (host; tests, small problems). */
int phet_host_fisher_yates(const void *draws, int32_t elem_bytes, int64_t n, int64_t take, uint32_t *out);

/* Permutations from swap targets on the device (csrc/fy_apply.cu; what PHET_PERM_REPLAY runs per block): draws is
 * host memory, [count][ld] targets of elem_bytes (2 or 4) as phet_mt_shuffle_rounds writes them; out: host
 * uint32[count][take], the first `take` entries of each permutation of arange(n). */
int phet_fy_apply(phet_ctx *ctx, const void *draws, int32_t elem_bytes, int64_t ld, int64_t n, int64_t count,
                  int64_t take, uint32_t *out);
/* The index sets of phet.py:397-402 (replace = False, m <= min(n0, n1)) drawn from `mt`: for each of S subsamples the
 * first m entries of np.random.permutation over the control rows, then over the case rows -- what
 * np.random.choice(rows, m, replace=False) returns -- installed like phet_set_index_sets.  idx_out: host
 * int32[S][2][m] (ORIGINAL row numbers) or NULL.  Requires phet_set_classes. */
int phet_draw_index_sets(phet_ctx *ctx, phet_mt *mt, int64_t S, int64_t m, int32_t *idx_out);
/* Device milliseconds of the last phet_run_pipeline call on a DEVICE matrix (one gene block): out[0] normalise +
 * transpose, out[1] Step 1, out[2] Step 3 (incl. the device Fisher-Yates in replay mode), from CUDA events recorded
 * on the context's stream between the phases (synchronises).  Host matrices interleave the phases per block: -1. */
int phet_phase_ms(phet_ctx *ctx, float out[3]);
/* out[0]: host seconds the draw thread spent in the last replay; out[1]: seconds the issuing thread waited for it. */
int phet_replay_stats(phet_ctx *ctx, double out[2]);

/* Testing hooks: host evaluations of the exact device routines (same source file) so the
 * CPU test-suite can check them without a GPU. */
double phet_host_p_two_sided_t(double t, double df, double ln_beta);
double phet_host_p_two_sided_z(double z);
uint32_t phet_host_perm_apply(uint64_t seed, uint64_t gene, uint32_t cls, uint32_t l, uint32_t i);

#ifdef __cplusplus
}
#endif
#endif /* PHET_B200_H */
