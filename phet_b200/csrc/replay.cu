// Reference-identical random draws inside the fit (PHET_PERM_REPLAY / draw_index_sets of phet_run_pipeline).
//
// The reference consumes the global NumPy stream in this order (phet.py:397-402, 484-487):
//     for s in range(S):  choice(control rows, m)  choice(case rows, m)          -> two whole-class shuffles each
//     for g in range(G):  permutation(control column)  permutation(case column)  -> two whole-class shuffles each
// A host thread replays that stream (csrc/mt_replay.cpp: sequential scan, swap targets out) block by block into a
// ring of pinned slots; the issuing thread uploads each block, turns it into permutations on the device
// (csrc/fy_apply.cu) and launches the consumers -- the index-set tables of Step 1, then Step 3 gene block by gene
// block -- so the device works on block b while the host draws block b+1, and the H2D copy of the matrix runs
// beside both.  A fit is then bounded by the sequential scan of the stream, not by Python loops or PCIe.
#include "common.cuh"

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>

namespace phet {

struct ReplayBlock {
    int kind;            // 0: index-set rounds (subsamples), 1: Step-3 rounds (genes)
    int64_t first, count;
    bool store;          // false: consume the stream only (genes of another rank)
};

struct ReplayRing {
    static constexpr int kSlots = 4;
    unsigned char *slot[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    size_t slot_bytes = 0;
    cudaEvent_t copied[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    std::mutex mu;
    std::condition_variable cv;
    int64_t produced = 0, released = 0;  // blocks drawn / slots handed back (counted over storing blocks only)
    int64_t blocks_done = 0;             // all blocks, storing or not
    bool abort = false;
    int error = 0;
    std::thread th;
    double host_seconds = 0.0;
};

void replay_ring_destroy(phet_ctx *c) {
    ReplayRing *r = c->ring;
    if (!r) return;
    if (r->th.joinable()) {
        {
            std::lock_guard<std::mutex> lk(r->mu);
            r->abort = true;
        }
        r->cv.notify_all();
        r->th.join();
    }
    for (int i = 0; i < ReplayRing::kSlots; ++i) {
        if (r->slot[i]) cudaFreeHost(r->slot[i]);
        if (r->copied[i]) cudaEventDestroy(r->copied[i]);
    }
    delete r;
    c->ring = nullptr;
}

static int ring_prepare(phet_ctx *c, size_t slot_bytes) {
    if (!c->ring) c->ring = new (std::nothrow) ReplayRing();
    ReplayRing *r = c->ring;
    if (!r) {
        set_error("out of host memory");
        return PHET_ERR_NOMEM;
    }
    if (r->slot_bytes < slot_bytes) {
        for (int i = 0; i < ReplayRing::kSlots; ++i) {
            if (r->slot[i]) cudaFreeHost(r->slot[i]);
            r->slot[i] = nullptr;
        }
        for (int i = 0; i < ReplayRing::kSlots; ++i) PHET_CUDA(cudaHostAlloc((void **)&r->slot[i], slot_bytes, cudaHostAllocDefault));
        r->slot_bytes = slot_bytes;
    }
    for (int i = 0; i < ReplayRing::kSlots; ++i)
        if (!r->copied[i]) PHET_CUDA(cudaEventCreateWithFlags(&r->copied[i], cudaEventDisableTiming));
    r->produced = r->released = r->blocks_done = 0;
    r->abort = false;
    r->error = 0;
    r->host_seconds = 0.0;
    return PHET_OK;
}

// Layout of one block in a slot / in draws_dev: [count][ld0] targets of the control shuffles, then [count][ld1] of the
// case shuffles; ld = class size rounded up to 8 elements.
struct DrawLayout {
    int eb0, eb1;          // bytes per target (2 when the class has <= 65536 cells)
    int64_t ld0, ld1;      // elements per row
    size_t off1(int64_t count) const { return ((size_t)count * (size_t)ld0 * (size_t)eb0 + 255) & ~(size_t)255; }
    size_t bytes(int64_t count) const { return off1(count) + (size_t)count * (size_t)ld1 * (size_t)eb1; }
};

static DrawLayout draw_layout(int64_t n0, int64_t n1) {
    DrawLayout L;
    L.eb0 = n0 <= 65536 ? 2 : 4;
    L.eb1 = n1 <= 65536 ? 2 : 4;
    L.ld0 = (n0 + 7) / 8 * 8;
    L.ld1 = (n1 + 7) / 8 * 8;
    return L;
}

// Draw thread: feeds the blocks, in stream order, to the parallel replay (csrc/mt_replay.cpp: generator / scanner /
// store workers) as ring slots become free, and publishes each storing block once its targets are complete.
static void producer_main(ReplayRing *r, phet_mt *mt, std::vector<ReplayBlock> blocks, DrawLayout L, int64_t n0, int64_t n1) {
    const auto t_begin = std::chrono::steady_clock::now();
    phet_mt_par *par = nullptr;
    int rc = phet_mt_par_begin(mt, 0, n0 > n1 ? n0 : n1, &par);
    std::vector<int64_t> tickets(blocks.size(), -1);
    size_t ns = 0, nw = 0;       // blocks submitted / completed
    int64_t stored_sub = 0, stored_done = 0;
    bool aborted = false;
    while (rc == 0 && nw < blocks.size() && !aborted) {
        // submit whatever has a slot to land in
        while (ns < blocks.size()) {
            const ReplayBlock &blk = blocks[ns];
            unsigned char *dst = nullptr;
            if (blk.store) {
                std::lock_guard<std::mutex> lk(r->mu);
                if (r->abort) {
                    aborted = true;
                    break;
                }
                if (stored_sub - r->released >= ReplayRing::kSlots) break;
                dst = r->slot[stored_sub % ReplayRing::kSlots];
                ++stored_sub;
            }
            phet_mt_job jobs[2];
            jobs[0].n = n0;
            jobs[0].out = dst;
            jobs[0].stride_bytes = L.ld0 * L.eb0;
            jobs[0].elem_bytes = L.eb0;
            jobs[1].n = n1;
            jobs[1].out = dst ? dst + L.off1(blk.count) : nullptr;
            jobs[1].stride_bytes = L.ld1 * L.eb1;
            jobs[1].elem_bytes = L.eb1;
            rc = phet_mt_par_submit(par, jobs, 2, blk.count, &tickets[ns]);
            if (rc != 0) break;
            ++ns;
        }
        if (rc != 0 || aborted) break;
        if (nw < ns) {
            rc = phet_mt_par_wait(par, tickets[nw]);
            if (rc != 0) break;
            {
                std::lock_guard<std::mutex> lk(r->mu);
                if (blocks[nw].store) r->produced = ++stored_done;
                r->blocks_done = (int64_t)nw + 1;
            }
            r->cv.notify_all();
            ++nw;
        } else {  // nothing in flight and no free slot: wait for the consumer to hand one back
            std::unique_lock<std::mutex> lk(r->mu);
            r->cv.wait(lk, [&] { return r->abort || stored_sub - r->released < ReplayRing::kSlots; });
            if (r->abort) aborted = true;
        }
    }
    if (par) phet_mt_par_end(par);  // joins its threads; the stream's final state goes back into mt
    {
        std::lock_guard<std::mutex> lk(r->mu);
        r->host_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
        if (rc != 0) r->error = rc;
    }
    r->cv.notify_all();
}

// Wait for the next storing block, upload it, and return the device pointers of its two halves.
struct ReplayConsumer {
    phet_ctx *c;
    ReplayRing *r;
    DrawLayout L;
    int64_t consumed = 0;  // storing blocks taken so far
    double wait_s = 0.0;

    int take(int64_t count, const unsigned char **d0, const unsigned char **d1) {
        const auto t0 = std::chrono::steady_clock::now();
        {
            std::unique_lock<std::mutex> lk(r->mu);
            r->cv.wait(lk, [&] { return r->error != 0 || r->produced > consumed; });
            if (r->error) {
                set_error("invalid: the draw thread failed");
                return r->error;
            }
        }
        wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const int s = (int)(consumed % ReplayRing::kSlots);
        const size_t bytes = L.bytes(count);
        if (int e = c->draws_dev.alloc(bytes)) return e;
        PHET_CUDA(cudaMemcpyAsync(c->draws_dev.p, r->slot[s], bytes, cudaMemcpyHostToDevice, c->stream));
        PHET_CUDA(cudaEventRecord(r->copied[s], c->stream));
        *d0 = c->draws_dev.p;
        *d1 = c->draws_dev.p + L.off1(count);
        ++consumed;
        // hand back the slot of the block before this one: its copy was queued a whole block of device work ago
        if (consumed >= 2) {
            const int sp = (int)((consumed - 2) % ReplayRing::kSlots);
            PHET_CUDA(cudaEventSynchronize(r->copied[sp]));
            {
                std::lock_guard<std::mutex> lk(r->mu);
                r->released = consumed - 1;
            }
            r->cv.notify_all();
        }
        return PHET_OK;
    }
};

int run_pipeline_replay(phet_ctx *c, const void *X, int is_device, int64_t N, int64_t G, int dtype,
                        const phet_pipeline_params *p) {
    const bool replay3 = p->perm_mode == PHET_PERM_REPLAY;
    PHET_REQUIRE(p->mt != nullptr, "replay needs a phet_mt (np.random state)");
    PHET_REQUIRE(c->classes_set, "phet_set_classes must be called first");
    PHET_REQUIRE(p->draw_index_sets || c->idx_set, "phet_set_index_sets must precede phet_run_pipeline (or set draw_index_sets)");
    PHET_REQUIRE(p->g_begin <= p->g3_begin && p->g3_begin <= p->g3_end && p->g3_end <= p->g_end,
                 "Step-3 gene range must lie inside the loaded gene range");
    const int64_t n0 = c->n0, n1 = c->n1, min_c = n0 < n1 ? n0 : n1;
    const int64_t m = p->m;
    int64_t S = c->S;
    if (p->draw_index_sets) {
        S = p->S_draw;
        PHET_REQUIRE(S >= 1 && m >= 1 && m <= 512, "index sets: need S >= 1, 1 <= m <= 512");
        PHET_REQUIRE(m <= min_c, "draw_index_sets covers replace=False only (m <= min(n0, n1)); draw on the host otherwise");
    } else {
        PHET_REQUIRE(m == c->m, "m does not match the uploaded index sets");
    }
    PHET_REQUIRE(0 <= p->s_begin && p->s_begin <= p->s_end && p->s_end <= S, "subsample range outside [0, S]");
    PHET_REQUIRE(c->step3_min_size == 0, "replay mode serves the two-class fit (PHET_OPT_STEP3_MIN_SIZE must be 0)");
    if (int e = prepare_matrix(c, N, G, dtype, p->normalize, p->storage, p->g_begin, p->g_end)) return e;
    c->phases_valid = false;
    const phet_quantile_pos &q = p->step1.q;
    PHET_REQUIRE(q.j_lo >= 0 && q.k_lo < m && q.j_hi >= 0 && q.k_hi < m && q.j_lo <= q.k_lo && q.j_hi <= q.k_hi,
                 "quantile positions outside [0, m)");
    if (int e = step3_prepare(c, m, p->perm_mode, p->perm_ctrl, p->perm_case, p->table_full, p->table_last, p->n_last)) return e;

    // ---- block plan
    const DrawLayout L = draw_layout(n0, n1);
    const size_t per_round = (size_t)L.ld0 * L.eb0 + (size_t)L.ld1 * L.eb1;
    size_t slot_target = (size_t)16 << 20;
    if (const char *mb = getenv("PHET_REPLAY_SLOT_MB")) slot_target = (size_t)(atoi(mb) > 0 ? atoi(mb) : 16) << 20;
    int64_t rounds_per_block = (int64_t)(slot_target / per_round);
    if (rounds_per_block < 1) rounds_per_block = 1;
    if (rounds_per_block >= c->num_sms) rounds_per_block = rounds_per_block / c->num_sms * c->num_sms;  // whole waves of the persistent Step-3 grid
    const size_t e_in = dtype == PHET_F32 ? 4 : 8;
    const int64_t gcount = p->g_end - p->g_begin;
    int64_t gb = gcount;  // genes per matrix block
    if (!is_device && gcount > 0) {
        gb = p->block_genes;
        if (gb <= 0) {
            size_t block_bytes = (size_t)128 << 20;
            if (const char *mb = getenv("PHET_BLOCK_MB")) block_bytes = (size_t)(atoi(mb) > 0 ? atoi(mb) : 128) << 20;
            gb = (int64_t)(block_bytes / ((size_t)N * e_in));
            gb = gb / c->num_sms * c->num_sms;
            if (gb < c->num_sms) gb = c->num_sms;
        }
        if (gb > gcount) gb = gcount;
    }
    std::vector<ReplayBlock> blocks;
    if (p->draw_index_sets)
        for (int64_t s0 = 0; s0 < S; s0 += rounds_per_block)
            blocks.push_back({0, s0, (s0 + rounds_per_block <= S) ? rounds_per_block : S - s0, true});
    if (replay3) {
        // genes outside the Step-3 range of this rank: consumed, not stored; inside: cut at matrix-block boundaries
        if (p->g3_begin > 0) blocks.push_back({1, 0, p->g3_begin, false});
        for (int64_t x0 = p->g_begin; x0 < p->g_end; x0 += gb) {
            const int64_t x1 = (x0 + gb < p->g_end) ? x0 + gb : p->g_end;
            const int64_t a = x0 > p->g3_begin ? x0 : p->g3_begin, b = x1 < p->g3_end ? x1 : p->g3_end;
            for (int64_t g0 = a; g0 < b; g0 += rounds_per_block)
                blocks.push_back({1, g0, (g0 + rounds_per_block <= b) ? rounds_per_block : b - g0, true});
        }
        if (p->g3_end < G && !p->stop_after_own_genes) blocks.push_back({1, p->g3_end, G - p->g3_end, false});
    }
    int64_t max_count = 1;
    for (const ReplayBlock &b : blocks)
        if (b.store && b.count > max_count) max_count = b.count;
    if (int e = ring_prepare(c, L.bytes(max_count))) return e;
    ReplayRing *r = c->ring;
    ReplayConsumer cons{c, r, L};
    r->th = std::thread(producer_main, r, p->mt, blocks, L, n0, n1);
    struct Joiner {  // whatever path leaves this function, the draw thread is stopped and joined first
        ReplayRing *r;
        ~Joiner() {
            if (!r->th.joinable()) return;
            {
                std::lock_guard<std::mutex> lk(r->mu);
                r->abort = true;
            }
            r->cv.notify_all();
            r->th.join();
        }
    } joiner{r};
    auto fail = [&](int e) {
        {
            std::lock_guard<std::mutex> lk(r->mu);
            r->abort = true;
        }
        r->cv.notify_all();
        return e;
    };

    // ---- the first matrix block starts crossing PCIe while the index sets are drawn
    if (!is_device && gcount > 0) {
        if (!c->copy_stream) PHET_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            if (!c->ev_copied[i]) PHET_CUDA(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
            if (!c->ev_consumed[i]) PHET_CUDA(cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
        }
        const size_t stage_bytes = (size_t)N * (size_t)gb * e_in;
        if (int e = c->stage[0].alloc(stage_bytes)) return fail(e);
        if (gcount > gb)
            if (int e = c->stage[1].alloc(stage_bytes)) return fail(e);
    }
    auto copy_block = [&](int64_t g0, int64_t gc, int b) -> int {
        const int buf = b & 1;
        if (b >= 2) PHET_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[buf], 0));
        PHET_CUDA(cudaMemcpy2DAsync(c->stage[buf].p, (size_t)gc * e_in, static_cast<const unsigned char *>(X) + (size_t)g0 * e_in,
                                    (size_t)G * e_in, (size_t)gc * e_in, (size_t)N, cudaMemcpyHostToDevice, c->copy_stream));
        PHET_CUDA(cudaEventRecord(c->ev_copied[buf], c->copy_stream));
        return 0;
    };
    if (!is_device && gcount > 0)
        if (int e = copy_block(p->g_begin, gb < gcount ? gb : gcount, 0)) return fail(e);

    // ---- index sets (phet.py:397-402): first m entries of each class shuffle, as ORIGINAL row numbers
    size_t bi = 0;
    if (p->draw_index_sets) {
        const int64_t n_idx = S * 2 * m;
        if (int e = c->idx.alloc((size_t)n_idx)) return fail(e);
        for (; bi < blocks.size() && blocks[bi].kind == 0; ++bi) {
            const ReplayBlock &blk = blocks[bi];
            const unsigned char *d0, *d1;
            if (int e = cons.take(blk.count, &d0, &d1)) return fail(e);
            uint32_t *out = reinterpret_cast<uint32_t *>(c->idx.p) + blk.first * 2 * m;
            if (int e = launch_fy_apply(c, d0, L.eb0, L.ld0, n0, blk.count, m, c->cls_rows.p, out, 2 * m, c->stream)) return fail(e);
            if (int e = launch_fy_apply(c, d1, L.eb1, L.ld1, n1, blk.count, m, c->cls_rows.p + n0, out + m, 2 * m, c->stream)) return fail(e);
        }
        if (p->idx_out) PHET_CUDA(cudaMemcpyAsync(p->idx_out, c->idx.p, sizeof(int32_t) * n_idx, cudaMemcpyDeviceToHost, c->stream));
        if (int e = finish_index_sets(c, S, m)) return fail(e);  // synchronises: idx_out is complete
    }

    // ---- matrix blocks: normalise + Step 1, then the Step-3 blocks of those genes as their draws arrive
    const bool keep = replay3 && p->keep_perm_tables != 0;
    // which view of the permutations Step 3 reads: the scatter kernel wants the slice of every cell (16 bits per cell
    // of both classes), the gather kernels the cell at every permuted position (32 bits per sliced position)
    const bool inverse = replay3 && step3_wants_inverse(c, m, p->n_last);
    const int64_t n_slices3 = (min_c + m - 1) / m;
    if (replay3) {
        const int64_t rows = keep ? (p->g3_end - p->g3_begin) : (rounds_per_block < max_count ? rounds_per_block : max_count);
        if (inverse) {
            if (int e = c->slice_of.alloc((size_t)(rows > 0 ? rows : 1) * (size_t)(n0 + n1))) return fail(e);
        } else {
            if (int e = c->perm0.alloc((size_t)(rows > 0 ? rows : 1) * (size_t)min_c)) return fail(e);
            if (int e = c->perm1.alloc((size_t)(rows > 0 ? rows : 1) * (size_t)min_c)) return fail(e);
        }
        c->slice_of_valid = inverse;
        c->perm_fwd_valid = !inverse;
    }
    auto step3_blocks = [&](int64_t x0, int64_t x1) -> int {
        if (!replay3) {
            const int64_t a = x0 > p->g3_begin ? x0 : p->g3_begin, b = x1 < p->g3_end ? x1 : p->g3_end;
            if (b > a) return launch_step3(c, a, b, m, p->perm_mode, p->seed, p->n_last);
            return 0;
        }
        for (; bi < blocks.size(); ++bi) {
            const ReplayBlock &blk = blocks[bi];
            if (!blk.store) continue;          // another rank's genes: the draw thread just runs through them
            if (blk.first >= x1) break;        // belongs to a later matrix block
            const unsigned char *d0, *d1;
            if (int e = cons.take(blk.count, &d0, &d1)) return e;
            const size_t grow = keep ? (size_t)(blk.first - p->g3_begin) : 0;  // kept tables: one row per gene of the range
            if (inverse) {
                uint16_t *so = c->slice_of.p + grow * (size_t)(n0 + n1);
                if (int e = launch_fy_apply_inverse(c, d0, L.eb0, L.ld0, n0, blk.count, min_c, c->cls_pos.p, so, n0 + n1, m, n_slices3 - 1, c->stream)) return e;
                if (int e = launch_fy_apply_inverse(c, d1, L.eb1, L.ld1, n1, blk.count, min_c, c->cls_pos.p + n0, so + n0, n0 + n1, m, n_slices3 - 1, c->stream)) return e;
            } else {
                const size_t row0 = grow * (size_t)min_c;
                if (int e = launch_fy_apply(c, d0, L.eb0, L.ld0, n0, blk.count, min_c, c->cls_pos.p, c->perm0.p + row0, min_c, c->stream)) return e;
                if (int e = launch_fy_apply(c, d1, L.eb1, L.ld1, n1, blk.count, min_c, c->cls_pos.p + n0, c->perm1.p + row0, min_c, c->stream)) return e;
            }
            c->perm_g0 = keep ? p->g3_begin : blk.first;
            if (int e = launch_step3(c, blk.first, blk.first + blk.count, m, PHET_PERM_HOST, 0, p->n_last)) return e;
        }
        return 0;
    };
    if (gcount > 0) {
        if (is_device) {
            const unsigned char *Xb = static_cast<const unsigned char *>(X) + (size_t)p->g_begin * e_in;
            if (int e = phase_mark(c, 0)) return fail(e);
            if (int e = launch_normalize(c, Xb, G, p->g_begin, gcount, dtype, p->normalize)) return fail(e);
            if (int e = phase_mark(c, 1)) return fail(e);
            if (int e = launch_step1(c, p->g_begin, p->g_end, p->s_begin, p->s_end, &p->step1, 0)) return fail(e);
            if (int e = phase_mark(c, 2)) return fail(e);
            if (int e = step3_blocks(p->g_begin, p->g_end)) return fail(e);
            if (int e = phase_mark(c, 3)) return fail(e);
            c->phases_valid = true;
        } else {
            int b = 0;
            for (int64_t g0 = p->g_begin; g0 < p->g_end; g0 += gb, ++b) {
                const int64_t gc = (g0 + gb <= p->g_end) ? gb : (p->g_end - g0);
                const int buf = b & 1;
                PHET_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copied[buf], 0));
                if (int e = launch_normalize(c, c->stage[buf].p, gc, g0, gc, dtype, p->normalize)) return fail(e);
                PHET_CUDA(cudaEventRecord(c->ev_consumed[buf], c->stream));
                if (g0 + gb < p->g_end) {  // the next block crosses PCIe while this one is computed and its draws are awaited
                    const int64_t g1 = g0 + gb, gc1 = (g1 + gb <= p->g_end) ? gb : (p->g_end - g1);
                    if (int e = copy_block(g1, gc1, b + 1)) return fail(e);
                }
                if (int e = launch_step1(c, g0, g0 + gc, p->s_begin, p->s_end, &p->step1, 0)) return fail(e);
                if (int e = step3_blocks(g0, g0 + gc)) return fail(e);
            }
        }
    }
    // trailing non-storing blocks (genes after this rank's range): let the stream reach the reference's final state
    r->th.join();
    if (r->error) {
        set_error("invalid: the draw thread failed");
        return r->error;
    }
    if (keep) {
        c->perm_g0 = p->g3_begin;
        c->perm_rows = p->g3_end - p->g3_begin;
        c->perm_cols = min_c;
    } else {
        c->perm_g0 = 0;
        c->perm_rows = 0;
    }
    c->replay_host_s = r->host_seconds;
    c->replay_wait_s = cons.wait_s;
    PHET_CUDA(cudaStreamSynchronize(c->stream));  // pinned slots and the caller's matrix are no longer read
    c->matrix_set = true;
    return PHET_OK;
}


// The index sets alone (phet.py:397-402, replace = False): S x {control shuffle, case shuffle} from `mt`, first m
// entries of each as ORIGINAL row numbers into c->idx, then the tables of the Step-1 kernels.
int replay_index_sets(phet_ctx *c, phet_mt *mt, int64_t S, int64_t m, int32_t *idx_out) {
    const int64_t n0 = c->n0, n1 = c->n1;
    const DrawLayout L = draw_layout(n0, n1);
    const size_t per_round = (size_t)L.ld0 * L.eb0 + (size_t)L.ld1 * L.eb1;
    int64_t rounds_per_block = (int64_t)(((size_t)16 << 20) / per_round);
    if (rounds_per_block < 1) rounds_per_block = 1;
    std::vector<ReplayBlock> blocks;
    for (int64_t s0 = 0; s0 < S; s0 += rounds_per_block)
        blocks.push_back({0, s0, (s0 + rounds_per_block <= S) ? rounds_per_block : S - s0, true});
    if (int e = ring_prepare(c, L.bytes(rounds_per_block < S ? rounds_per_block : S))) return e;
    ReplayRing *r = c->ring;
    ReplayConsumer cons{c, r, L};
    r->th = std::thread(producer_main, r, mt, blocks, L, n0, n1);
    struct Joiner {
        ReplayRing *r;
        ~Joiner() {
            if (!r->th.joinable()) return;
            {
                std::lock_guard<std::mutex> lk(r->mu);
                r->abort = true;
            }
            r->cv.notify_all();
            r->th.join();
        }
    } joiner{r};
    const int64_t n_idx = S * 2 * m;
    if (int e = c->idx.alloc((size_t)n_idx)) return e;
    for (const ReplayBlock &blk : blocks) {
        const unsigned char *d0, *d1;
        if (int e = cons.take(blk.count, &d0, &d1)) return e;
        uint32_t *out = reinterpret_cast<uint32_t *>(c->idx.p) + blk.first * 2 * m;
        if (int e = launch_fy_apply(c, d0, L.eb0, L.ld0, n0, blk.count, m, c->cls_rows.p, out, 2 * m, c->stream)) return e;
        if (int e = launch_fy_apply(c, d1, L.eb1, L.ld1, n1, blk.count, m, c->cls_rows.p + n0, out + m, 2 * m, c->stream)) return e;
    }
    r->th.join();
    if (r->error) {
        set_error("invalid: the draw thread failed");
        return r->error;
    }
    if (idx_out) PHET_CUDA(cudaMemcpyAsync(idx_out, c->idx.p, sizeof(int32_t) * n_idx, cudaMemcpyDeviceToHost, c->stream));
    c->replay_host_s = r->host_seconds;
    c->replay_wait_s = cons.wait_s;
    return finish_index_sets(c, S, m);  // synchronises
}

}  // namespace phet

using namespace phet;

extern "C" {

// Test hook / building block: permutations from swap targets on the device.  draws: host [count][ld] of 2- or
// 4-byte targets (as phet_mt_shuffle_rounds writes them); out: host uint32 [count][take].
int phet_fy_apply(phet_ctx *c, const void *draws, int32_t elem_bytes, int64_t ld, int64_t n, int64_t count, int64_t take,
                  uint32_t *out) {
    PHET_REQUIRE(c != nullptr && draws != nullptr && out != nullptr, "NULL argument");
    PHET_REQUIRE(n >= 1 && count >= 1 && take >= 1 && take <= n && ld >= n - 1, "bad shape");
    PHET_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "elem_bytes must be 2 or 4");
    PHET_REQUIRE(elem_bytes == 4 || n <= 65536, "16-bit targets need n <= 65536");
    PHET_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)count * (size_t)ld * (size_t)elem_bytes;
    if (int e = c->draws_dev.alloc(bytes + 16)) return e;
    DevBuf<uint32_t> tmp;
    if (int e = tmp.alloc((size_t)count * (size_t)take)) return e;
    PHET_CUDA(cudaMemcpyAsync(c->draws_dev.p, draws, bytes, cudaMemcpyHostToDevice, c->stream));
    int rc = launch_fy_apply(c, c->draws_dev.p, elem_bytes, ld, n, count, take, nullptr, tmp.p, take, c->stream);
    if (rc == 0) {
        cudaError_t e = cudaMemcpyAsync(out, tmp.p, sizeof(uint32_t) * (size_t)count * (size_t)take, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "phet_fy_apply copy back", __FILE__, __LINE__);
    }
    tmp.release();
    return rc;
}

int phet_draw_index_sets(phet_ctx *c, phet_mt *mt, int64_t S, int64_t m, int32_t *idx_out) {
    PHET_REQUIRE(c != nullptr && mt != nullptr, "NULL argument");
    PHET_REQUIRE(c->classes_set, "phet_set_classes must be called first");
    PHET_REQUIRE(S >= 1 && m >= 1 && m <= 512, "index sets: need S >= 1, 1 <= m <= 512");
    PHET_REQUIRE(m <= c->n0 && m <= c->n1, "phet_draw_index_sets covers replace=False only (m <= min(n0, n1))");
    PHET_CUDA(cudaSetDevice(c->device));
    return replay_index_sets(c, mt, S, m, idx_out);
}

// Host seconds the draw thread spent in the last replay, and seconds the issuing thread waited for it.
int phet_replay_stats(phet_ctx *c, double out[2]) {
    PHET_REQUIRE(c != nullptr && out != nullptr, "NULL argument");
    out[0] = c->replay_host_s;
    out[1] = c->replay_wait_s;
    return PHET_OK;
}

}  // extern "C"
