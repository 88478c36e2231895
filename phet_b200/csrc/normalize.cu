// Stage 0: per-gene z-score (population sd) + NaN/inf -> 0, fused with the transpose from the
// caller's cells x genes row-major matrix to the gene-major, class-grouped HBM layout.
// Replaces /root/reference/src/model/phet.py:291,299 (scipy.stats.zscore(X, axis=0); nan_to_num).
//
// HBM layout produced: xg[g][n_pad] (storage dtype), position q < n0 holds control cell
// ctrl_rows[q], position n0 + q' holds case cell case_rows[q'], positions >= N are 0.
// Algorithmic traffic: read X twice (statistics, then normalise) + write once: 3*N*G*e bytes.
#include "common.cuh"

namespace phet {

constexpr int kTile = 32;
constexpr int kRowsPerBlock = 8;

// Shifted sums per column over a chunk of rows: s = sum(x - K), ss = sum((x - K)^2), K = X[0][g].
// fp64 accumulation; the shift removes the cancellation of the one-pass variance formula.
template <typename Tin>
__global__ void __launch_bounds__(kTile *kRowsPerBlock)
    k_colstats_partial(const Tin *__restrict__ X, int64_t ldx, int64_t N, int64_t G, int64_t rows_per_chunk,
                       double *__restrict__ part) {
    __shared__ double sm_s[kRowsPerBlock][kTile + 1];
    __shared__ double sm_q[kRowsPerBlock][kTile + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t g = (int64_t)blockIdx.x * kTile + tx;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = min(r0 + rows_per_chunk, N);
    double s = 0.0, q = 0.0;
    if (g < G) {
        const double K = (double)X[g];
        int64_t r = r0 + ty;
        for (; r + 3 * kRowsPerBlock < r1; r += 4 * kRowsPerBlock) {
            const double d0 = (double)X[r * ldx + g] - K;
            const double d1 = (double)X[(r + kRowsPerBlock) * ldx + g] - K;
            const double d2 = (double)X[(r + 2 * kRowsPerBlock) * ldx + g] - K;
            const double d3 = (double)X[(r + 3 * kRowsPerBlock) * ldx + g] - K;
            s += d0; q += d0 * d0;
            s += d1; q += d1 * d1;
            s += d2; q += d2 * d2;
            s += d3; q += d3 * d3;
        }
        for (; r < r1; r += kRowsPerBlock) {
            const double d = (double)X[r * ldx + g] - K;
            s += d;
            q += d * d;
        }
    }
    sm_s[ty][tx] = s;
    sm_q[ty][tx] = q;
    __syncthreads();
    if (ty == 0 && g < G) {
        double ts = 0.0, tq = 0.0;
#pragma unroll
        for (int i = 0; i < kRowsPerBlock; ++i) {
            ts += sm_s[i][tx];
            tq += sm_q[i][tx];
        }
        part[((int64_t)blockIdx.y * 2 + 0) * G + g] = ts;
        part[((int64_t)blockIdx.y * 2 + 1) * G + g] = tq;
    }
}

template <typename Tin>
__global__ void k_colstats_final(const Tin *__restrict__ X, int64_t N, int64_t G, int chunks,
                                 const double *__restrict__ part, double *__restrict__ mean,
                                 double *__restrict__ sd) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double s = 0.0, q = 0.0;
    for (int c = 0; c < chunks; ++c) {  // fixed order: reproducible
        s += part[((int64_t)c * 2 + 0) * G + g];
        q += part[((int64_t)c * 2 + 1) * G + g];
    }
    const double n = (double)N;
    const double K = (double)X[g];
    double var = (q - s * s / n) / n;
    if (var < 0.0) var = 0.0;
    mean[g] = K + s / n;
    sd[g] = sqrt(var);
}

// Tile of kTQ grouped positions x kTG genes.  A warp reads one source row at a time: 4 x 128
// contiguous bytes (the rows of a class-grouped, shuffled layout are scattered over X, so wide
// row segments are what keeps DRAM pages busy); a thread owns genes lane, lane+32, ..., so the
// transposed SMEM writes hit 32 distinct banks (tile pitch kTQ+1 is odd).  Each gene row of the
// tile leaves as kTQ contiguous elements (2 per lane, 8- or 16-byte stores).
// z = (x - mu) * (1 / sd) in fp64: the reciprocal is taken once per gene (a per-element fp64 division
// costs ~15 instructions); the map stays monotone in x, so ties and order are those of x.
constexpr int kTQ = 64;
constexpr int kTG = 128;
constexpr int kNTThreads = 256;
constexpr int kGPT = kTG / 32;  // genes per thread

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(kNTThreads, 3)
    k_normalize_transpose(const Tin *__restrict__ X, int64_t ldx, int64_t N, int64_t G, int64_t n_pad,
                          const int32_t *__restrict__ rows_grouped, const double *__restrict__ mean,
                          const double *__restrict__ sd, int normalize, Tout *__restrict__ xg) {
    extern __shared__ __align__(16) unsigned char tile_raw[];
    Tout(*tile)[kTQ + 1] = reinterpret_cast<Tout(*)[kTQ + 1]>(tile_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g0 = (int64_t)blockIdx.x * kTG;
    const int64_t q0 = (int64_t)blockIdx.y * kTQ;
    double mu[kGPT], rs[kGPT];
#pragma unroll
    for (int i = 0; i < kGPT; ++i) {
        const int64_t g = g0 + lane + 32 * i;
        mu[i] = 0.0;
        rs[i] = 1.0;
        if (normalize && g < G) {
            mu[i] = mean[g];
            rs[i] = 1.0 / sd[g];  // sd == 0 -> inf -> 0 * inf = NaN -> 0 below, like x / 0
        }
    }
    constexpr int kRowsPerWarp = kTQ / (kNTThreads / 32);
    int64_t src[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
        const int64_t q = q0 + warp * kRowsPerWarp + r;
        src[r] = (q < N) ? (int64_t)rows_grouped[q] : (int64_t)-1;
    }
    Tin x[kRowsPerWarp][kGPT];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
#pragma unroll
        for (int i = 0; i < kGPT; ++i) {
            const int64_t g = g0 + lane + 32 * i;
            x[r][i] = (src[r] >= 0 && g < G) ? X[src[r] * ldx + g] : Tin(0);
        }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
#pragma unroll
        for (int i = 0; i < kGPT; ++i) {
            double z = normalize ? ((double)x[r][i] - mu[i]) * rs[i] : (double)x[r][i];
            if (!isfinite(z) || src[r] < 0) z = 0.0;  // nan_to_num(nan=0, posinf=0, neginf=0), phet.py:299
            tile[lane + 32 * i][warp * kRowsPerWarp + r] = (Tout)z;
        }
    }
    __syncthreads();
    constexpr int kGenesPerWarp = kTG / (kNTThreads / 32);
    const int64_t q = q0 + 2 * lane;  // n_pad is even for both storage types: q < n_pad implies q + 1 < n_pad
    if (q < n_pad) {
#pragma unroll 4
        for (int k = 0; k < kGenesPerWarp; ++k) {
            const int gi = warp * kGenesPerWarp + k;
            const int64_t g = g0 + gi;
            if (g < G) {
                struct alignas(2 * sizeof(Tout)) Pair { Tout a, b; };
                Pair pr;
                pr.a = tile[gi][2 * lane];
                pr.b = tile[gi][2 * lane + 1];
                *reinterpret_cast<Pair *>(xg + g * n_pad + q) = pr;
            }
        }
    }
}

// X points at the first gene of the block; ldx = row pitch of X in elements; the block covers
// genes [g0, g0 + gcount) and is written to rows (g0 - xg_g0 ...) of the gene-major buffer.
template <typename Tin>
static int run_normalize(phet_ctx *c, const Tin *X, int64_t ldx, int64_t g0, int64_t gcount, int normalize) {
    const int64_t N = c->N, G = gcount;
    double *mean = c->mean.p + g0, *sd = c->sd.p + g0;
    int chunks = 1;
    if (normalize) {
        const int64_t gx = (G + kTile - 1) / kTile;
        // Row chunks are a function of N ALONE: the merge order of the column sums -- and with it every bit of the
        // mean and sd -- must not depend on how many genes are normalised together (gene blocks of the pipelined
        // path, gene shards of a multi-GPU fit), so that any partition of the genes gives the same z-scores.
        int64_t rows_per_chunk = 2048;
        if ((N + rows_per_chunk - 1) / rows_per_chunk > 64) rows_per_chunk = (N + 63) / 64;
        chunks = (int)((N + rows_per_chunk - 1) / rows_per_chunk);
        if (int e = c->colpart.alloc((size_t)chunks * 2 * G)) return e;
        dim3 blk(kTile, kRowsPerBlock);
        dim3 grd((unsigned)gx, (unsigned)chunks);
        k_colstats_partial<Tin><<<grd, blk, 0, c->stream>>>(X, ldx, N, G, rows_per_chunk, c->colpart.p);
        c->launches++;
        k_colstats_final<Tin><<<(unsigned)((G + 255) / 256), 256, 0, c->stream>>>(X, N, G, chunks, c->colpart.p, mean,
                                                                                  sd);
        c->launches++;
    }
    dim3 grd((unsigned)((G + kTG - 1) / kTG), (unsigned)((c->n_pad + kTQ - 1) / kTQ));
    const size_t e_st = c->storage == PHET_F32 ? 4 : 8;
    unsigned char *out = c->xg.p + (size_t)(g0 - c->xg_g0) * (size_t)c->n_pad * e_st;
    const size_t tile_bytes = (size_t)kTG * (kTQ + 1) * e_st;
    if (c->storage == PHET_F32) {
        auto kern = k_normalize_transpose<Tin, float>;
        PHET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));
        kern<<<grd, kNTThreads, tile_bytes, c->stream>>>(X, ldx, N, G, c->n_pad, c->rows_grouped.p, mean, sd, normalize,
                                                         (float *)out);
    } else {
        auto kern = k_normalize_transpose<Tin, double>;
        PHET_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));
        kern<<<grd, kNTThreads, tile_bytes, c->stream>>>(X, ldx, N, G, c->n_pad, c->rows_grouped.p, mean, sd, normalize,
                                                         (double *)out);
    }
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    return 0;
}

int launch_normalize(phet_ctx *c, const void *Xblock, int64_t ldx, int64_t g0, int64_t gcount, int dtype,
                     int normalize) {
    if (dtype == PHET_F32) return run_normalize<float>(c, (const float *)Xblock, ldx, g0, gcount, normalize);
    return run_normalize<double>(c, (const double *)Xblock, ldx, g0, gcount, normalize);
}

}  // namespace phet
