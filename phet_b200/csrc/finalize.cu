// Steps 2, 3b and 4 on per-gene vectors: Fisher standardisation, uniform binning of o with
// host-supplied edges (shared-memory atomics for the bin counts), and the final combine.
// Replaces /root/reference/src/model/phet.py:435-444, 449-462, 500-504, 510-516.
#include "common.cuh"

#include <math_constants.h>

namespace phet {

constexpr int kRedThreads = 1024;
constexpr double kSeedValue = 0.001;  // phet.py:29

// f = (sum - 2S)/sqrt(4S); f < 0 -> 0.001; |f|; nan/inf -> 0.   r = sum_delta / S; nan/inf -> 0.
__global__ void k_fisher(const double *__restrict__ acc_f, const double *__restrict__ acc_r, int64_t G,
                         double S, double *__restrict__ f, double *__restrict__ r) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double v = (acc_f[g] - 2.0 * S) / sqrt(4.0 * S);
    if (v < 0.0) v = kSeedValue;
    v = fabs(v);
    if (!isfinite(v)) v = 0.0;
    f[g] = v;
    double d = acc_r[g] / S;
    if (!isfinite(d)) d = 0.0;
    r[g] = d;
}

// Single-block deterministic reductions: out[0] = min(x), out[1] = max(x), out[2] = sum(x).
__global__ void __launch_bounds__(kRedThreads) k_reduce3(const double *__restrict__ x, int64_t G,
                                                        double *__restrict__ out) {
    __shared__ double smin[kRedThreads], smax[kRedThreads], ssum[kRedThreads];
    const int t = threadIdx.x;
    double mn = CUDART_INF, mx = -CUDART_INF, sm = 0.0;
    for (int64_t g = t; g < G; g += kRedThreads) {
        const double v = x[g];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
        sm += v;
    }
    smin[t] = mn;
    smax[t] = mx;
    ssum[t] = sm;
    __syncthreads();
    for (int s = kRedThreads / 2; s > 0; s >>= 1) {
        if (t < s) {
            smin[t] = fmin(smin[t], smin[t + s]);
            smax[t] = fmax(smax[t], smax[t + s]);
            ssum[t] += ssum[t + s];
        }
        __syncthreads();
    }
    if (t == 0) {
        out[0] = smin[0];
        out[1] = smax[0];
        out[2] = ssum[0];
    }
}

// bin = searchsorted(edges[1:-1], o, side="right") = #{inner edges <= o}; counts via SMEM atomics.
__global__ void k_bin(const double *__restrict__ o, int64_t G, const double *__restrict__ edges, int B,
                      signed char *__restrict__ bins, unsigned long long *__restrict__ counts) {
    extern __shared__ unsigned int hist[];
    for (int i = threadIdx.x; i < B; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) {
        const double v = o[g];
        int b = 0;
        for (int i = 1; i < B; ++i) b += (edges[i] <= v) ? 1 : 0;
        bins[g] = (signed char)b;
        atomicAdd(&hist[b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B; i += blockDim.x)
        if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
}

// fw = f * weight[bin]  (phet.py:511-512)
__global__ void k_weight(const double *__restrict__ f, const signed char *__restrict__ bins,
                         const double *__restrict__ w, int64_t G, double *__restrict__ fw) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) fw[g] = f[g] * w[bins[g]];
}

// score = r/sum(r) + fw/sum(fw); nan/inf -> 0  (phet.py:510,513-515)
__global__ void k_combine(const double *__restrict__ r, const double *__restrict__ fw,
                          const double *__restrict__ sums_r, const double *__restrict__ sums_fw, int64_t G,
                          double *__restrict__ scores) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double v = r[g] / sums_r[2] + fw[g] / sums_fw[2];
    if (!isfinite(v)) v = 0.0;
    scores[g] = v;
}

// np.linspace(min o, max o, B + 1) on the device, operation for operation (numpy function_base.linspace:
// step = (hi - lo) / B; y[i] = i * step + lo with the product and the sum rounded separately; y[B] = hi), so
// the edges -- and with them every bin -- are the ones sklearn's KBinsDiscretizer(strategy="uniform") computes
// (_discretization.py:339).  min == max: one bin holding everything (edges -inf, +inf, ...).
__global__ void k_edges(const double *__restrict__ mm, int B, double *__restrict__ edges) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double lo = mm[0], hi = mm[1];
    if (lo == hi) {
        edges[0] = -CUDART_INF;
        for (int i = 1; i <= B; ++i) edges[i] = CUDART_INF;
        return;
    }
    const double delta = __dsub_rn(hi, lo);
    const double step = __ddiv_rn(delta, (double)B);
    for (int i = 0; i <= B; ++i) {
        const double prod = (step != 0.0) ? __dmul_rn((double)i, step) : __dmul_rn(__ddiv_rn((double)i, (double)B), delta);
        edges[i] = __dadd_rn(prod, lo);
    }
    edges[B] = hi;
}

static inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

int launch_fisher(phet_ctx *c, int64_t S_total) {
    k_fisher<<<blocks_for(c->G, 256), 256, 0, c->stream>>>(c->acc.p, c->acc.p + c->G, c->G, (double)S_total, c->f.p,
                                                          c->r.p);
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    return 0;
}

int launch_o_minmax(phet_ctx *c, double out[2]) {
    k_reduce3<<<1, kRedThreads, 0, c->stream>>>(c->o.p, c->G, c->scalars.p);
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    double h[3];
    PHET_CUDA(cudaMemcpyAsync(h, c->scalars.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    out[0] = h[0];
    out[1] = h[1];
    return 0;
}

int launch_bin(phet_ctx *c, int32_t B, int64_t *counts_out) {
    PHET_CUDA(cudaMemsetAsync(c->counts.p, 0, sizeof(unsigned long long) * B, c->stream));
    k_bin<<<blocks_for(c->G, 256), 256, sizeof(unsigned int) * B, c->stream>>>(c->o.p, c->G, c->edges.p, B, c->bins.p,
                                                                              c->counts.p);
    c->launches++;
    PHET_CUDA(cudaGetLastError());
    if (counts_out) {
        PHET_CUDA(cudaMemcpyAsync(counts_out, c->counts.p, sizeof(unsigned long long) * B, cudaMemcpyDeviceToHost,
                                  c->stream));
        PHET_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int launch_combine(phet_ctx *c, int32_t B, double *scores_out) {
    (void)B;
    double *fw = c->fw.p;
    double *sums_r = c->scalars.p + 4;
    double *sums_fw = c->scalars.p + 8;
    k_weight<<<blocks_for(c->G, 256), 256, 0, c->stream>>>(c->f.p, c->bins.p, c->weights.p, c->G, fw);
    k_reduce3<<<1, kRedThreads, 0, c->stream>>>(c->r.p, c->G, sums_r);
    k_reduce3<<<1, kRedThreads, 0, c->stream>>>(fw, c->G, sums_fw);
    k_combine<<<blocks_for(c->G, 256), 256, 0, c->stream>>>(c->r.p, fw, sums_r, sums_fw, c->G, c->scores.p);
    c->launches += 4;
    PHET_CUDA(cudaGetLastError());
    if (scores_out) {
        PHET_CUDA(cudaMemcpyAsync(scores_out, c->scores.p, sizeof(double) * c->G, cudaMemcpyDeviceToHost, c->stream));
        PHET_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// Fisher standardisation, edges, binning and combine back to back on the stream; one synchronisation at the end.
int launch_finalize(phet_ctx *c, int64_t S_total, int32_t B) {   // queued on the stream, no synchronisation
    if (int e = launch_fisher(c, S_total)) return e;
    k_reduce3<<<1, kRedThreads, 0, c->stream>>>(c->o.p, c->G, c->scalars.p);
    k_edges<<<1, 32, 0, c->stream>>>(c->scalars.p, B, c->edges.p);
    c->launches += 2;
    if (int e = launch_bin(c, B, nullptr)) return e;
    if (int e = launch_combine(c, B, nullptr)) return e;
    PHET_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace phet
