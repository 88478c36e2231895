// Input path: MatrixMarket file -> dense float32 cells x genes matrix resident in HBM,
// optionally filtered, ready for phet_load_matrix / phet_run_pipeline with is_device = 1.
// Replaces /root/reference/src/train.py:82-85 (sc.read_mtx -> AnnData -> to_df().to_numpy(), nan_to_num) and
// train.py:88-118 (drop cells with sum|x| < 5; keep genes whose CPM exceeds 1 in >= minimum_samples cells).
//
// Host: the file is mapped, cut at line boundaries into one span per thread, and each span is parsed straight
// into its slice of the triplet arrays (values go text -> double -> float, the rounding scipy.io.mmread followed
// by .astype(float32) performs).  Device: the 12-byte triplets are what crosses PCIe (not the dense matrix);
// they are scattered into the dense matrix, and the filter runs where the matrix lives.  The two decisions of the
// filter are integer outputs and are reproduced exactly: the row sums follow NumPy's pairwise summation order
// (blocks of <= 128 elements, eight strided partial sums per block, halves split at multiples of eight) in fp32,
// and the CPM test is the same two correctly rounded fp32 operations (x * 1e6f, then / row sum).
#include "common.cuh"

#include <charconv>
#include <cmath>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

namespace phet {

namespace {

inline bool is_space(char ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\v' || ch == '\f'; }

inline const char *skip_space(const char *p, const char *end) {
    while (p < end && is_space(*p)) ++p;
    return p;
}

inline const char *line_end(const char *p, const char *end) {
    const void *q = memchr(p, '\n', (size_t)(end - p));
    return q ? static_cast<const char *>(q) : end;
}

inline bool parse_i64(const char *&p, const char *end, int64_t &out) {
    p = skip_space(p, end);
    bool neg = false;
    if (p < end && (*p == '+' || *p == '-')) neg = (*p++ == '-');
    if (p >= end || *p < '0' || *p > '9') return false;
    int64_t x = 0;
    while (p < end && *p >= '0' && *p <= '9') x = x * 10 + (*p++ - '0');
    out = neg ? -x : x;
    return true;
}

inline bool parse_f64(const char *&p, const char *end, double &out) {
    p = skip_space(p, end);
    if (p < end && *p == '+') ++p;
    const std::from_chars_result res = std::from_chars(p, end, out);
    if (res.ec == std::errc::result_out_of_range) {   // strtod semantics: +-inf on overflow, 0 on underflow
        const char *q = p;
        const bool neg = (*q == '-');
        bool big = false;   // decide by the exponent sign
        for (const char *t = p; t < res.ptr; ++t)
            if (*t == 'e' || *t == 'E') big = !(t + 1 < res.ptr && t[1] == '-');
        out = big ? (neg ? -HUGE_VAL : HUGE_VAL) : (neg ? -0.0 : 0.0);
        p = res.ptr;
        return true;
    }
    if (res.ec != std::errc()) return false;
    p = res.ptr;
    return true;
}

std::string lower_token(const char *&p, const char *end) {
    p = skip_space(p, end);
    std::string t;
    while (p < end && !is_space(*p) && *p != '\n') t.push_back((char)tolower((unsigned char)*p++));
    return t;
}

struct Span {
    const char *b, *e;
    int64_t lines = 0, first = 0;
    int bad = 0;
    int64_t bad_line = 0;
};

int64_t count_lines(const char *p, const char *end) {
    int64_t n = 0;
    while (p < end) {
        const char *le = line_end(p, end);
        if (skip_space(p, le) < le) ++n;
        p = le + 1;
    }
    return n;
}

void parse_span(Span &s, const phet_mtx &hdr, int32_t *r, int32_t *c, float *v) {
    const char *p = s.b;
    int64_t k = s.first;
    while (p < s.e) {
        const char *le = line_end(p, s.e);
        const char *q = skip_space(p, le);
        if (q < le) {
            int64_t i = 0, j = 0;
            double x = 1.0;
            bool ok = parse_i64(q, le, i) && parse_i64(q, le, j);
            if (ok && !hdr.pattern) ok = parse_f64(q, le, x);
            if (ok) ok = skip_space(q, le) == le && i >= 1 && i <= hdr.rows && j >= 1 && j <= hdr.cols;
            if (!ok) {
                if (!s.bad) s.bad_line = k - s.first;
                s.bad = 1;
                i = j = 1;
                x = 0.0;
            }
            r[k] = (int32_t)(i - 1);
            c[k] = (int32_t)(j - 1);
            v[k] = (float)x;
            ++k;
        }
        p = le + 1;
    }
}

}  // namespace

int mtx_parse(const char *path, int threads, phet_mtx &m) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        set_error(std::string("invalid: cannot open ") + path);
        return PHET_ERR_INVALID;
    }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) {
        close(fd);
        set_error(std::string("invalid: empty or unreadable file ") + path);
        return PHET_ERR_INVALID;
    }
    const size_t size = (size_t)st.st_size;
    void *map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) {
        set_error(std::string("invalid: mmap failed for ") + path);
        return PHET_ERR_INVALID;
    }
    madvise(map, size, MADV_SEQUENTIAL);
    const char *p = static_cast<const char *>(map), *end = p + size;
    int rc = PHET_OK;
    do {
        // banner: %%MatrixMarket matrix coordinate <real|double|integer|pattern> <general|symmetric>
        const char *le = line_end(p, end);
        const char *q = p;
        const std::string banner = lower_token(q, le), object = lower_token(q, le), format = lower_token(q, le),
                          field = lower_token(q, le), symmetry = lower_token(q, le);
        if (banner != "%%matrixmarket" || object != "matrix") {
            set_error("invalid: not a MatrixMarket matrix file (banner missing)");
            rc = PHET_ERR_INVALID;
            break;
        }
        if (format != "coordinate" || !(field == "real" || field == "double" || field == "integer" || field == "pattern") ||
            !(symmetry == "general" || symmetry == "symmetric")) {
            set_error("unsupported: MatrixMarket '" + format + " " + field + " " + symmetry +
                      "' (coordinate real|integer|pattern general|symmetric are read)");
            rc = PHET_ERR_UNSUPPORTED;
            break;
        }
        m.pattern = field == "pattern";
        m.symmetric = symmetry == "symmetric";
        p = le + 1;
        // comments and blank lines, then the size line
        bool have_size = false;
        while (p < end) {
            le = line_end(p, end);
            q = skip_space(p, le);
            if (q < le && *q != '%') {
                have_size = parse_i64(q, le, m.rows) && parse_i64(q, le, m.cols) && parse_i64(q, le, m.declared);
                p = le + 1;
                break;
            }
            p = le + 1;
        }
        if (!have_size || m.rows < 0 || m.cols < 0 || m.declared < 0 || m.rows > INT32_MAX || m.cols > INT32_MAX) {
            set_error("invalid: MatrixMarket size line missing or out of range");
            rc = PHET_ERR_INVALID;
            break;
        }
        if (p > end) p = end;
        // spans at line boundaries
        int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
        if (T < 1) T = 1;
        if (T > 64) T = 64;
        const size_t body = (size_t)(end - p);
        if (body < (size_t)T * 65536) T = (int)(body / 65536) + 1;
        std::vector<Span> spans((size_t)T);
        const char *cur = p;
        for (int t = 0; t < T; ++t) {
            const char *stop = (t == T - 1) ? end : p + body * (size_t)(t + 1) / (size_t)T;
            if (stop < cur) stop = cur;
            if (t != T - 1 && stop < end) {
                stop = line_end(stop, end);
                if (stop < end) ++stop;
            }
            spans[(size_t)t].b = cur;
            spans[(size_t)t].e = stop;
            cur = stop;
        }
        {
            std::vector<std::thread> pool;
            for (int t = 1; t < T; ++t) pool.emplace_back([&spans, t] { spans[(size_t)t].lines = count_lines(spans[(size_t)t].b, spans[(size_t)t].e); });
            spans[0].lines = count_lines(spans[0].b, spans[0].e);
            for (auto &th : pool) th.join();
        }
        int64_t total = 0;
        for (auto &s : spans) {
            s.first = total;
            total += s.lines;
        }
        if (total != m.declared) {
            set_error("invalid: MatrixMarket file declares " + std::to_string(m.declared) + " entries but holds " +
                      std::to_string(total));
            rc = PHET_ERR_INVALID;
            break;
        }
        m.r.resize((size_t)total);
        m.c.resize((size_t)total);
        m.v.resize((size_t)total);
        {
            std::vector<std::thread> pool;
            for (int t = 1; t < T; ++t)
                pool.emplace_back([&spans, &m, t] { parse_span(spans[(size_t)t], m, m.r.data(), m.c.data(), m.v.data()); });
            parse_span(spans[0], m, m.r.data(), m.c.data(), m.v.data());
            for (auto &th : pool) th.join();
        }
        for (auto &s : spans)
            if (s.bad) {
                set_error("invalid: malformed MatrixMarket entry " + std::to_string(s.first + s.bad_line + 1) +
                          " (expected 'row col value' with 1 <= row <= rows, 1 <= col <= cols)");
                rc = PHET_ERR_INVALID;
                break;
            }
    } while (false);
    munmap(map, size);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// device side

constexpr int kMtxThreads = 256;

__global__ void k_mtx_scatter(const int32_t *__restrict__ r, const int32_t *__restrict__ c, const float *__restrict__ v,
                              int64_t nnz, float *__restrict__ X, int64_t ld, int symmetric) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = r[k], j = c[k];
        const float x = v[k];
        atomicAdd(X + i * ld + j, x);   // duplicates add up (coo -> csr sums them)
        if (symmetric && i != j) atomicAdd(X + j * ld + i, x);
    }
}

// Row sums of |x| in NumPy's pairwise order; non-finite cells are zeroed in place first (np.nan_to_num,
// train.py:85).  prog[0] = number of leaves L; prog[1 .. 2L]: (offset, length) of every leaf; then the
// post-order program of the summation tree: 0 = push the next leaf sum, 1 = pop two, push their sum.
__global__ void __launch_bounds__(kMtxThreads) k_row_abs_sums(float *__restrict__ X, int64_t N, int64_t G,
                                                             const int32_t *__restrict__ prog, int prog_len,
                                                             float *__restrict__ rowsum) {
    extern __shared__ float leaf[];
    const int L = prog[0];
    for (int64_t row = blockIdx.x; row < N; row += gridDim.x) {
        float *x = X + row * G;
        for (int l = threadIdx.x; l < L; l += kMtxThreads) {
            float *a = x + prog[1 + 2 * l];
            const int n = prog[2 + 2 * l];
            auto cell = [&](int i) {
                float t = a[i];
                if (!isfinite(t)) {
                    t = 0.0f;
                    a[i] = 0.0f;
                }
                return fabsf(t);
            };
            float res;
            if (n < 8) {
                res = 0.0f;
                for (int i = 0; i < n; ++i) res = __fadd_rn(res, cell(i));
            } else {
                float p[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) p[k] = cell(k);
                int i = 8;
                for (; i < n - (n % 8); i += 8) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) p[k] = __fadd_rn(p[k], cell(i + k));
                }
                res = __fadd_rn(__fadd_rn(__fadd_rn(p[0], p[1]), __fadd_rn(p[2], p[3])),
                                __fadd_rn(__fadd_rn(p[4], p[5]), __fadd_rn(p[6], p[7])));
                for (; i < n; ++i) res = __fadd_rn(res, cell(i));
            }
            leaf[l] = res;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float stk[40];
            int sp = 0, next = 0;
            for (int k = 1 + 2 * L; k < prog_len; ++k) {
                if (prog[k] == 0) stk[sp++] = leaf[next++];
                else {
                    --sp;
                    stk[sp - 1] = __fadd_rn(stk[sp - 1], stk[sp]);
                }
            }
            rowsum[row] = L > 0 ? stk[0] : 0.0f;
        }
        __syncthreads();
    }
}

// cnt[j] += #{ kept rows i in this slab : fl(fl(|x_ij| * 1e6f) / rowsum_i) > 1 }   (train.py:103-107)
__global__ void __launch_bounds__(kMtxThreads) k_cpm_counts(const float *__restrict__ X, int64_t G,
                                                           const int32_t *__restrict__ rows, int64_t n_rows,
                                                           const float *__restrict__ rowsum, int32_t *__restrict__ cnt) {
    const int64_t j = (int64_t)blockIdx.x * kMtxThreads + threadIdx.x;
    if (j >= G) return;
    const int64_t per = (n_rows + gridDim.y - 1) / gridDim.y;
    const int64_t b = per * blockIdx.y, e = (b + per < n_rows) ? b + per : n_rows;
    int32_t n = 0;
    for (int64_t t = b; t < e; ++t) {
        const int64_t i = rows[t];
        const float q = __fdiv_rn(__fmul_rn(fabsf(X[i * G + j]), 1000000.0f), rowsum[i]);
        n += (q > 1.0f) ? 1 : 0;
    }
    if (n) atomicAdd(cnt + j, n);
}

__global__ void k_mtx_compact(const float *__restrict__ X, int64_t G, const int32_t *__restrict__ rows, int64_t n_rows,
                              const int32_t *__restrict__ cols, int64_t n_cols, float *__restrict__ out) {
    const int64_t total = n_rows * n_cols;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = k / n_cols, b = k - a * n_cols;
        out[k] = X[(int64_t)rows[a] * G + cols[b]];
    }
}

// leaves and post-order program of NumPy's pairwise_sum recursion for n elements (numpy/_core/src/umath/
// loops_utils.h.src, @TYPE@_pairwise_sum: n <= 128 is a leaf, else n2 = n / 2; n2 -= n2 % 8)
static void pairwise_program(int64_t off, int64_t n, std::vector<int32_t> &leaves, std::vector<int32_t> &ops) {
    if (n <= 128) {
        leaves.push_back((int32_t)off);
        leaves.push_back((int32_t)n);
        ops.push_back(0);
        return;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    pairwise_program(off, n2, leaves, ops);
    pairwise_program(off + n2, n - n2, leaves, ops);
    ops.push_back(1);
}

int mtx_stage(phet_ctx *c, const phet_mtx &m, int do_filter, int64_t minimum_samples, int64_t dims_out[2]) {
    const int64_t N = m.rows, G = m.cols, nnz = (int64_t)m.v.size();
    PHET_REQUIRE(N >= 1 && G >= 1, "empty matrix");
    PHET_REQUIRE(G <= INT32_MAX / 2, "too many columns");
    PHET_CUDA(cudaSetDevice(c->device));
    if (int e = c->mtx_dense.alloc((size_t)N * (size_t)G)) return e;
    if (int e = c->mtx_r.alloc((size_t)nnz + 1)) return e;
    if (int e = c->mtx_c.alloc((size_t)nnz + 1)) return e;
    if (int e = c->mtx_v.alloc((size_t)nnz + 1)) return e;
    if (int e = c->mtx_rowsum.alloc((size_t)N)) return e;
    PHET_CUDA(cudaMemsetAsync(c->mtx_dense.p, 0, sizeof(float) * (size_t)N * (size_t)G, c->stream));
    if (nnz > 0) {
        PHET_CUDA(cudaMemcpyAsync(c->mtx_r.p, m.r.data(), sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, c->stream));
        PHET_CUDA(cudaMemcpyAsync(c->mtx_c.p, m.c.data(), sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, c->stream));
        PHET_CUDA(cudaMemcpyAsync(c->mtx_v.p, m.v.data(), sizeof(float) * nnz, cudaMemcpyHostToDevice, c->stream));
        int64_t grid = (nnz + kMtxThreads - 1) / kMtxThreads;
        if (grid > (int64_t)c->num_sms * 16) grid = (int64_t)c->num_sms * 16;
        k_mtx_scatter<<<(unsigned)grid, kMtxThreads, 0, c->stream>>>(c->mtx_r.p, c->mtx_c.p, c->mtx_v.p, nnz, c->mtx_dense.p, G,
                                                                     m.symmetric ? 1 : 0);
        c->launches++;
    }
    // row sums (and nan_to_num) in NumPy's order
    std::vector<int32_t> leaves, ops, prog;
    pairwise_program(0, G, leaves, ops);
    const int L = (int)(leaves.size() / 2);
    prog.push_back(L);
    prog.insert(prog.end(), leaves.begin(), leaves.end());
    prog.insert(prog.end(), ops.begin(), ops.end());
    if (int e = c->mtx_prog.alloc(prog.size())) return e;
    PHET_CUDA(cudaMemcpyAsync(c->mtx_prog.p, prog.data(), sizeof(int32_t) * prog.size(), cudaMemcpyHostToDevice, c->stream));
    {
        int64_t grid = N < (int64_t)c->num_sms * 8 ? N : (int64_t)c->num_sms * 8;
        const size_t smem = sizeof(float) * (size_t)L;
        PHET_REQUIRE(smem <= (size_t)kMaxSmemOptin, "too many columns for the row-sum kernel");
        if (smem > 48 * 1024) PHET_CUDA(cudaFuncSetAttribute(k_row_abs_sums, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_row_abs_sums<<<(unsigned)grid, kMtxThreads, smem, c->stream>>>(c->mtx_dense.p, N, G, c->mtx_prog.p,
                                                                                             (int)prog.size(), c->mtx_rowsum.p);
        c->launches++;
    }
    PHET_CUDA(cudaGetLastError());
    c->mtx_rows_host.clear();
    c->mtx_cols_host.clear();
    if (!do_filter) {
        PHET_CUDA(cudaStreamSynchronize(c->stream));   // the triplet vectors may be freed by the caller after this
        for (int64_t i = 0; i < N; ++i) c->mtx_rows_host.push_back((int32_t)i);
        for (int64_t j = 0; j < G; ++j) c->mtx_cols_host.push_back((int32_t)j);
        c->mtx_out.alias(c->mtx_dense.p, (size_t)N * (size_t)G);
        c->mtx_N = N;
        c->mtx_G = G;
        dims_out[0] = N;
        dims_out[1] = G;
        return PHET_OK;
    }
    // cells with sum |x| >= 5 (train.py:90-92)
    std::vector<float> rs((size_t)N);
    PHET_CUDA(cudaMemcpyAsync(rs.data(), c->mtx_rowsum.p, sizeof(float) * (size_t)N, cudaMemcpyDeviceToHost, c->stream));
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < N; ++i)
        if (rs[(size_t)i] >= 5.0f) c->mtx_rows_host.push_back((int32_t)i);
    const int64_t Nk = (int64_t)c->mtx_rows_host.size();
    std::vector<int32_t> cnt((size_t)G, 0);
    if (Nk > 0) {
        if (int e = c->mtx_rows.alloc((size_t)Nk)) return e;
        if (int e = c->mtx_cnt.alloc((size_t)G)) return e;
        PHET_CUDA(cudaMemcpyAsync(c->mtx_rows.p, c->mtx_rows_host.data(), sizeof(int32_t) * (size_t)Nk, cudaMemcpyHostToDevice,
                                  c->stream));
        PHET_CUDA(cudaMemsetAsync(c->mtx_cnt.p, 0, sizeof(int32_t) * (size_t)G, c->stream));
        const int64_t gx = (G + kMtxThreads - 1) / kMtxThreads;
        int64_t gy = ((int64_t)c->num_sms * 8 + gx - 1) / gx;
        if (gy > Nk) gy = Nk;
        if (gy > 65535) gy = 65535;
        if (gy < 1) gy = 1;
        k_cpm_counts<<<dim3((unsigned)gx, (unsigned)gy), kMtxThreads, 0, c->stream>>>(c->mtx_dense.p, G, c->mtx_rows.p, Nk,
                                                                                    c->mtx_rowsum.p, c->mtx_cnt.p);
        c->launches++;
        PHET_CUDA(cudaMemcpyAsync(cnt.data(), c->mtx_cnt.p, sizeof(int32_t) * (size_t)G, cudaMemcpyDeviceToHost, c->stream));
        PHET_CUDA(cudaStreamSynchronize(c->stream));
    }
    // train.py:108-110
    if (Nk <= minimum_samples || minimum_samples > Nk / 2) minimum_samples = Nk / 2;
    for (int64_t j = 0; j < G; ++j)
        if ((int64_t)cnt[(size_t)j] >= minimum_samples) c->mtx_cols_host.push_back((int32_t)j);
    const int64_t Gk = (int64_t)c->mtx_cols_host.size();
    c->mtx_N = Nk;
    c->mtx_G = Gk;
    dims_out[0] = Nk;
    dims_out[1] = Gk;
    c->mtx_out.release();
    if (Nk == 0 || Gk == 0) return PHET_OK;
    if (int e = c->mtx_cols.alloc((size_t)Gk)) return e;
    if (int e = c->mtx_out.alloc((size_t)Nk * (size_t)Gk)) return e;
    PHET_CUDA(cudaMemcpyAsync(c->mtx_cols.p, c->mtx_cols_host.data(), sizeof(int32_t) * (size_t)Gk, cudaMemcpyHostToDevice,
                              c->stream));
    {
        int64_t grid = (Nk * Gk + kMtxThreads - 1) / kMtxThreads;
        if (grid > (int64_t)c->num_sms * 32) grid = (int64_t)c->num_sms * 32;
        k_mtx_compact<<<(unsigned)grid, kMtxThreads, 0, c->stream>>>(c->mtx_dense.p, G, c->mtx_rows.p, Nk, c->mtx_cols.p, Gk,
                                                                     c->mtx_out.p);
        c->launches++;
    }
    PHET_CUDA(cudaGetLastError());
    PHET_CUDA(cudaStreamSynchronize(c->stream));
    c->mtx_dense.release();   // only the filtered copy stays resident
    c->mtx_r.release();
    c->mtx_c.release();
    c->mtx_v.release();
    return PHET_OK;
}

}  // namespace phet
