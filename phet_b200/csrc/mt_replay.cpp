// Native replay of NumPy's legacy global RNG (RandomState: MT19937 + masked rejection) for the draws PHet makes:
//
//   phet.py:401-402   np.random.choice(rows, m, replace=False)  ==  rows[np.random.permutation(len(rows))[:m]]
//                     np.random.choice(rows, m, replace=True)   ==  rows[np.random.randint(0, len(rows), m)]
//   phet.py:486-487   np.random.permutation(column)             ==  column[np.random.permutation(len(column))]
//
// (equivalences and identical state advance: SURVEY.md section 8c, tests/test_mt_replay_cpu.py).  The reference
// replays 2 S + 2 G whole-class shuffles per fit -- 1.5e9 accepted draws at BASELINE config 4 -- and a Python
// loop over np.random.permutation needs 25.7 s for them.  Here the work is split in two:
//
//   (1) the DRAWS: np.random.permutation(n) is `for i = n-1 .. 1: j = rk_interval(i); swap(x[i], x[j])` with
//       rk_interval(i) = first value of the stream (masked to the bits of i) that is <= i.  Which raw outputs are
//       accepted depends on the running i, so this part is sequential -- but it is pure streaming: generate
//       tempered MT19937 outputs, mask, compare, compact the accepted ones.  It is vectorised 64 outputs at a
//       time (AVX-512): a value <= i - k at lane k is accepted whatever the earlier lanes did, a value > i is
//       rejected whatever they did, and the few in between are settled in lane order with the exact count.
//       Output: the swap targets j, 16 bits each when n <= 65536.  (The generator's twist can be issued in small
//       steps between the scan's trips -- env PHET_MT_PUMP -- to fill the slots the scan's dependency chain leaves
//       idle; measured on the GPU box's Xeon it changes nothing, both halves being bound by the two 512-bit ports.)
//   (2) the SWAPS: random access, but independent across permutations -- applied on the GPU (csrc/fy_apply.cu)
//       from the uploaded j streams, or by phet_host_fisher_yates for tests and small problems.
//
// Only WHERE each permutation starts in the stream is inherently sequential.  phet_mt_par_* therefore runs one
// scanner thread with a count-only scan (no compaction, no stores) that records the generator state (2.5 KB) at
// the start of every permutation, and worker threads that redo the scan from those states, storing the targets.
//
// The state goes in and out in np.random.get_state() / set_state() form (key[624], pos), so the caller's global
// stream continues exactly where the reference's would.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#define PHET_X86 1
#else
#define PHET_X86 0
#endif

#include "../../include/phet_b200.h"

namespace phet {
void set_error(const std::string &msg);
}

#define MT_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,bmi,bmi2,popcnt")))

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;
constexpr int kCarry = 128;     // outputs of the previous block kept in front of the current one (>= the widest trip of a scan)
// bytes the count-only scan prefetches ahead of itself (0: off); measured on the GPU box's Xeon: 5.6 -> 5.0 us per permutation
static const int kScanPrefetch = getenv("PHET_SCAN_PREFETCH") ? atoi(getenv("PHET_SCAN_PREFETCH")) : 4096;
constexpr int kRingMirror = 320;  // outputs of the ring's start mirrored behind its end (>= the count-only scan's look-ahead of 256)
constexpr int kGenSteps = 41;   // copy + 40 vector steps make the next block (see phet_mt::gen_step)
// generator steps per 64-output trip of the scan (5 * 15.6 outputs >= 64); env PHET_MT_PUMP overrides (0: whole blocks)
static const int kPump = getenv("PHET_MT_PUMP") ? atoi(getenv("PHET_MT_PUMP")) : 0;

inline void cpu_relax() {
#if PHET_X86
    __builtin_ia32_pause();
#endif
}

inline uint32_t temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

inline uint32_t twist1(uint32_t a, uint32_t b) {
    const uint32_t y = (a & kUpper) | (b & kLower);
    return (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
}

void twist_scalar(uint32_t *key) {
    int i = 0;
    for (; i < kN - kM; ++i) key[i] = key[i + kM] ^ twist1(key[i], key[i + 1]);
    for (; i < kN - 1; ++i) key[i] = key[i + (kM - kN)] ^ twist1(key[i], key[i + 1]);
    key[kN - 1] = key[kM - 1] ^ twist1(key[kN - 1], key[0]);
}

void temper_scalar(const uint32_t *key, uint32_t *out) {
    for (int i = 0; i < kN; ++i) out[i] = temper(key[i]);
}

inline uint32_t bitmask(uint32_t v) {  // smallest 2^k - 1 >= v  (numpy: random_interval / gen_mask)
    v |= v >> 1;
    v |= v >> 2;
    v |= v >> 4;
    v |= v >> 8;
    v |= v >> 16;
    return v;
}

#if PHET_X86
MT_AVX512 inline __m512i twist16(__m512i a, __m512i b) {
    const __m512i y = _mm512_or_si512(_mm512_and_si512(a, _mm512_set1_epi32((int)kUpper)),
                                      _mm512_and_si512(b, _mm512_set1_epi32((int)kLower)));
    const __mmask16 odd = _mm512_test_epi32_mask(y, _mm512_set1_epi32(1));
    const __m512i sh = _mm512_srli_epi32(y, 1);
    return _mm512_mask_xor_epi32(sh, odd, sh, _mm512_set1_epi32((int)kMatrixA));
}

MT_AVX512 inline __m512i temper16(__m512i y) {
    y = _mm512_xor_si512(y, _mm512_srli_epi32(y, 11));
    y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 7), _mm512_set1_epi32((int)0x9d2c5680u)));
    y = _mm512_xor_si512(y, _mm512_and_si512(_mm512_slli_epi32(y, 15), _mm512_set1_epi32((int)0xefc60000u)));
    return _mm512_xor_si512(y, _mm512_srli_epi32(y, 18));
}

MT_AVX512 void temper_avx512(const uint32_t *key, uint32_t *out) {
    for (int k = 0; k < kN; k += 16)  // 624 = 39 * 16
        _mm512_storeu_si512(out + k, temper16(_mm512_loadu_si512(key + k)));
}

MT_AVX512 void temper16_avx512(const uint32_t *key, uint16_t *out) {  // low 16 bits of the tempered outputs
    for (int k = 0; k < kN; k += 16)
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(out + k), _mm512_cvtepi32_epi16(temper16(_mm512_loadu_si512(key + k))));
}

// One vector of the in-place twist of key[0..624) (the array has >= 32 words of padding behind it), tempered into
// out.  Steps 1..40 in order make one block:
//   1..14   k = 0, 16, .., 208     new[k] = old[k + 397] ^ tw(old[k], old[k + 1])        (reads up to old[620])
//   15      k = 224 (3 lanes: the others would clobber old words the second half still needs); then old[624] := new[0]
//   16..39  k = 227, 243, .., 595  new[k] = new[k - 227] ^ tw(old[k], old[k + 1])
//   40      k = 611 (13 lanes; lane 12 pairs old[623] with new[0] through key[624])
template <bool TEMPER>
MT_AVX512 inline void twist_step_avx512(uint32_t *key, uint32_t *out, int step) {
    int k, off;
    __mmask16 m = 0xFFFF;
    if (step <= 15) {
        k = 16 * (step - 1);
        off = kM;
        if (step == 15) m = 0x7;
    } else {
        k = (kN - kM) + 16 * (step - 16);
        off = -(kN - kM);
        if (step == 40) m = (__mmask16)((1u << (kN - k)) - 1u);
    }
    const __m512i a = _mm512_loadu_si512(key + k), b = _mm512_loadu_si512(key + k + 1);
    const __m512i c = _mm512_loadu_si512(key + k + off);
    const __m512i v = _mm512_xor_si512(c, twist16(a, b));
    if (m == 0xFFFF) {  // (plain stores: a masked store does not forward to the loads of the next step)
        _mm512_storeu_si512(key + k, v);
        if (TEMPER) _mm512_storeu_si512(out + k, temper16(v));
    } else {
        _mm512_mask_storeu_epi32(key + k, m, v);
        if (TEMPER) _mm512_mask_storeu_epi32(out + k, m, temper16(v));
        if (step == 15) key[kN] = key[0];
    }
}

MT_AVX512 void twist_block_avx512(uint32_t *key) {  // state -> next state, no outputs (a generator that only runs ahead)
    for (int step = 1; step <= 40; ++step) twist_step_avx512<false>(key, nullptr, step);
}
#endif

bool cpu_has_avx512() {
#if PHET_X86
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                           __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq") &&
                           __builtin_cpu_supports("bmi2") && getenv("PHET_MT_SCALAR") == nullptr;
    return ok;
#else
    return false;
#endif
}

}  // namespace

// The generator: NumPy's (key, pos) plus a window of tempered outputs.  buf = [kCarry carried outputs of the
// previous block | 624 outputs of the current block]; [pos, end) are the unconsumed ones.  The NEXT block is made
// in small steps (gen_step) into nbuf / work while the current one is consumed.
struct phet_mt {
    alignas(64) uint32_t key_a[kN + 32];
    alignas(64) uint32_t key_b[kN + 32];
    alignas(64) uint32_t buf_a[kCarry + kN + 32];
    alignas(64) uint32_t buf_b[kCarry + kN + 32];
    uint32_t *cur = key_a;    // state of the block being consumed
    uint32_t *work = key_b;   // gstep == 0: state of the block BEFORE it; otherwise the next block in the making
    uint32_t *buf = buf_a, *nbuf = buf_b;
    int pos = kCarry + kN, end = kCarry + kN;
    int gstep = 0;            // 0 .. kGenSteps: progress of the next block
    bool avx512 = false;
    int64_t block_no = 0;     // blocks entered since load(): stream position = 624 * block_no + (pos - kCarry)

    void load(const uint32_t *k, int p) {
        cur = key_a;
        work = key_b;
        buf = buf_a;
        nbuf = buf_b;
        memcpy(cur, k, sizeof(uint32_t) * kN);
        memset(cur + kN, 0, sizeof(uint32_t) * 32);
        memcpy(work, cur, sizeof(uint32_t) * (kN + 32));
#if PHET_X86
        if (avx512) temper_avx512(cur, buf + kCarry);
        else
#endif
            temper_scalar(cur, buf + kCarry);
        pos = kCarry + p;
        end = kCarry + kN;
        gstep = 0;
        block_no = 0;
    }
    inline int64_t consumed() const { return block_no * kN + (pos - kCarry); }  // outputs consumed since load(key, 0)
    // one step of the next block (step 0 copies the state: `work` stops being the previous block's state, so it
    // must not run while outputs of the previous block are still unread, pos < kCarry -- see state())
    inline void gen_step() {
        if (gstep == 0) {
            memcpy(work, cur, sizeof(uint32_t) * kN);
            gstep = 1;
#if PHET_X86
            if (avx512) return;
#endif
            twist_scalar(work);
            temper_scalar(work, nbuf + kCarry);
            gstep = kGenSteps;
            return;
        }
#if PHET_X86
        twist_step_avx512<true>(work, nbuf + kCarry, gstep);
#endif
        ++gstep;
    }
    // between two trips of a scan: keep the next block ahead of the reader
    inline void pump() {
        if (pos < kCarry) return;
        for (int s = 0; s < kPump && gstep < kGenSteps; ++s) gen_step();
    }
    // fewer than kCarry outputs left: finish the next block and put the leftovers in front of it
    void refill() {
        const int rem = end - pos;
        while (gstep < kGenSteps) gen_step();
        if (rem > 0) memmove(nbuf + kCarry - rem, buf + pos, sizeof(uint32_t) * (size_t)rem);
        std::swap(buf, nbuf);
        std::swap(cur, work);  // work now holds the state of the block just left: the "previous" state until step 0
        pos = kCarry - rem;
        end = kCarry + kN;
        gstep = 0;
        ++block_no;
    }
    inline void ensure(int k) {  // k <= kCarry
        if (end - pos < k) refill();
    }
    inline uint32_t next32() {
        if (pos == end) refill();
        return buf[pos++];
    }
    void state(uint32_t *k, int *p) const {
        if (pos >= kCarry) {
            memcpy(k, cur, sizeof(uint32_t) * kN);
            *p = pos - kCarry;  // 624: regenerate on the next draw (NumPy's convention)
        } else {  // still inside the carried tail of the previous block (gstep == 0: `work` is that block's state)
            memcpy(k, work, sizeof(uint32_t) * kN);
            *p = kN - (kCarry - pos);
        }
    }
};

namespace {

// ---- where a scan takes the tempered outputs from -----------------------------------------------------------------
// (a) the generator's own window: one thread generates and scans
struct SeqSrc {
    phet_mt &e;
    const uint32_t *p = nullptr;
    inline void need(int k) {  // >= k contiguous outputs at the current position (k <= kCarry)
        e.ensure(k);
        p = e.buf + e.pos;
    }
#if PHET_X86
    MT_AVX512 inline __m512i vec(int q) const { return _mm512_loadu_si512(p + 16 * q); }
#endif
    inline uint32_t at(int k) const { return p[k]; }
    inline void advance(int used) {
        e.pos += used;
        e.pump();
    }
    inline uint32_t next() { return e.next32(); }
};

// (b) a ring that generator threads fill ahead of the readers (parallel replay below).  E = uint16_t when every
// permutation is of <= 65536 cells (only the low 16 bits of an output can matter): half the bytes between cores.
struct OutRing {
    void *w = nullptr;            // [cap + kRingMirror] elements: the first kRingMirror are mirrored behind the end
    int ebytes = 4;
    size_t cap = 0;               // elements; a multiple of region_elems
    size_t region_elems = 0;      // 624 * blocks per region
    size_t nreg = 0;
    std::unique_ptr<std::atomic<uint64_t>[]> ready;   // per ring slot: 1 + index of the region it holds, once complete
    alignas(64) std::atomic<uint64_t> low{0};         // outputs below this stream position are no longer read
    std::atomic<bool> stop{false};
    std::atomic<uint64_t> reader_spins{0}, writer_spins{0};   // diagnostics: who waits for whom
};

template <typename E>
struct RingSrc {
    OutRing *r;
    uint64_t abs;            // stream position (outputs since the start of block 0)
    size_t idx;              // abs mod cap
    uint64_t known;          // outputs in [region start of abs, known) are known to be in the ring
    const E *p = nullptr;
    RingSrc(OutRing *ring, uint64_t a)
        : r(ring), abs(a), idx((size_t)(a % ring->cap)), known(a / ring->region_elems * ring->region_elems) {}  // regions before a's own may be gone
    inline void need(int k) {
        if (__builtin_expect(abs + (uint64_t)k > known, 0)) {
            unsigned spins = 0;
            for (;;) {
                const uint64_t reg = known / r->region_elems;   // next region not yet seen complete
                if (r->ready[reg % r->nreg].load(std::memory_order_acquire) == reg + 1) {
                    known = (reg + 1) * r->region_elems;
                    if (abs + (uint64_t)k <= known) break;
                    continue;
                }
                if (r->stop.load(std::memory_order_relaxed)) break;
                cpu_relax();
                if ((++spins & 4095u) == 0) {
                    r->reader_spins.fetch_add(4096, std::memory_order_relaxed);
                    std::this_thread::yield();
                }
            }
        }
        p = static_cast<const E *>(r->w) + idx;
    }
#if PHET_X86
    MT_AVX512 inline __m512i vec(int q) const {
        if (sizeof(E) == 2) return _mm512_cvtepu16_epi32(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(p + 16 * q)));
        return _mm512_loadu_si512(p + 16 * q);
    }
#endif
    inline uint32_t at(int k) const { return (uint32_t)p[k]; }
    inline void advance(int used) {
        abs += (uint64_t)used;
        idx += (size_t)used;
        if (idx >= r->cap) idx -= r->cap;
    }
    inline uint32_t next() {
        need(1);
        const uint32_t v = (uint32_t)p[0];
        advance(1);
        return v;
    }
};

// ---- the draws of one np.random.permutation(n): dst[t] = rk_interval(n - 1 - t), t = 0 .. n - 2
template <typename J, bool STORE, typename Src>
void shuffle_draws_scalar(Src &in, uint32_t n, J *dst) {
    size_t t = 0;
    for (uint32_t i = n - 1; i >= 1; --i) {
        const uint32_t mask = bitmask(i);
        uint32_t v;
        while ((v = (in.next() & mask)) > i) {
        }
        if (STORE) dst[t] = (J)v;
        ++t;
    }
}

#if PHET_X86
template <typename J>
MT_AVX512 inline void store_compressed(J *dst, __m512i v, __mmask16 keep, int cnt);

template <>
MT_AVX512 inline void store_compressed<uint32_t>(uint32_t *dst, __m512i v, __mmask16 keep, int cnt) {
    const __m512i c = _mm512_maskz_compress_epi32(keep, v);
    _mm512_mask_storeu_epi32(dst, (__mmask16)((1u << cnt) - 1u), c);
}

template <>
MT_AVX512 inline void store_compressed<uint16_t>(uint16_t *dst, __m512i v, __mmask16 keep, int cnt) {
    const __m512i c = _mm512_maskz_compress_epi32(keep, v);
    _mm256_mask_storeu_epi16(dst, (__mmask16)((1u << cnt) - 1u), _mm512_cvtepi32_epi16(c));
}

// 64 outputs per trip.  With i the running position at the start of the trip, lane k (k-th output of the trip)
// sees i_k = i - #{accepted among lanes < k} in [i - k, i]:  v > i is rejected and v <= i - k is accepted whatever
// the earlier lanes did.  The few lanes in between (probability ~ k / 2^bits(i) each) are settled one by one, in
// order, with the exact i_k -- the lanes before them are all decided by then.  A trip stops where i would leave
// the octave [lo, mask] that fixes the bit mask.
constexpr uint32_t kScalarBelow = 96;  // the last draws of a permutation go one output at a time

template <typename J, bool STORE, typename Src>
MT_AVX512 void shuffle_draws_avx512(Src &in, uint32_t n, J *dst) {
    const __m512i lane = _mm512_set_epi32(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m512i lanek[4] = {lane, _mm512_add_epi32(lane, _mm512_set1_epi32(16)), _mm512_add_epi32(lane, _mm512_set1_epi32(32)),
                              _mm512_add_epi32(lane, _mm512_set1_epi32(48))};
    size_t t = 0;
    uint32_t i = n - 1;
    while (i >= kScalarBelow) {
        const uint32_t mask = bitmask(i), lo = (mask >> 1) + 1;  // every i in [lo, mask] shares this mask
        const uint32_t floor_i = lo > kScalarBelow ? lo : kScalarBelow;
        const __m512i vm = _mm512_set1_epi32((int)mask);
        // count-only scans of the wide octaves take 128 outputs per trip: the trip is bound by the dependency chain
        // i -> compare -> popcount -> i, not by the number of lanes, and with 2^15+ values per octave the lanes left
        // undecided by the wider uncertainty of i stay rare (~0.25 per trip at mask 32767)
        if (!STORE && kCarry >= 128 && mask >= 8191u) {
            while (i >= floor_i) {
                in.need(128);
                const __m512i vi = _mm512_set1_epi32((int)i);
                uint64_t sure[2] = {0, 0}, amb[2] = {0, 0};
#pragma GCC unroll 8
                for (int q = 0; q < 8; ++q) {
                    const __m512i vq = _mm512_and_si512(in.vec(q), vm);
                    const __mmask16 nrej = _mm512_cmple_epu32_mask(vq, vi);
                    const __mmask16 sq = _mm512_cmple_epu32_mask(_mm512_add_epi32(vq, _mm512_add_epi32(lanek[q & 3], _mm512_set1_epi32(64 * (q >> 2)))), vi);
                    sure[q >> 2] |= (uint64_t)sq << (16 * (q & 3));
                    amb[q >> 2] |= (uint64_t)(__mmask16)(nrej & ~sq) << (16 * (q & 3));
                }
                while (__builtin_expect(amb[0] != 0, 0)) {
                    const int k = __builtin_ctzll(amb[0]);
                    const uint32_t ik = i - (uint32_t)__builtin_popcountll(sure[0] & ((1ull << k) - 1ull));
                    if ((in.at(k) & mask) <= ik) sure[0] |= 1ull << k;
                    amb[0] &= amb[0] - 1;
                }
                const uint32_t t0 = (uint32_t)__builtin_popcountll(sure[0]);
                while (__builtin_expect(amb[1] != 0, 0)) {
                    const int k = __builtin_ctzll(amb[1]);
                    const uint32_t ik = i - t0 - (uint32_t)__builtin_popcountll(sure[1] & ((1ull << k) - 1ull));
                    if ((in.at(64 + k) & mask) <= ik) sure[1] |= 1ull << k;
                    amb[1] &= amb[1] - 1;
                }
                uint32_t total = t0 + (uint32_t)__builtin_popcountll(sure[1]);
                const uint32_t room = i - lo + 1;
                int used = 128;
                if (__builtin_expect(total >= room, 0)) {
                    if (room <= t0) used = __builtin_ctzll(_pdep_u64(1ull << (room - 1), sure[0])) + 1;
                    else used = 64 + __builtin_ctzll(_pdep_u64(1ull << (room - t0 - 1), sure[1])) + 1;
                    total = room;
                }
                t += total;
                i -= total;
                in.advance(used);
            }
            continue;
        }
        while (i >= floor_i) {
            in.need(64);
            const __m512i vi = _mm512_set1_epi32((int)i);
            __m512i v[4];
            uint64_t sure = 0, amb = 0;
#pragma GCC unroll 4
            for (int q = 0; q < 4; ++q) {
                v[q] = _mm512_and_si512(in.vec(q), vm);
                // v > i: rejected; v + k <= i: accepted (both against the same broadcast of i; v + k cannot wrap: v < 2^31)
                const __mmask16 nrej = _mm512_cmple_epu32_mask(v[q], vi);
                const __mmask16 sq = _mm512_cmple_epu32_mask(_mm512_add_epi32(v[q], lanek[q]), vi);
                sure |= (uint64_t)sq << (16 * q);
                amb |= (uint64_t)(__mmask16)(nrej & ~sq) << (16 * q);
            }
            while (__builtin_expect(amb != 0, 0)) {  // in lane order: everything before lane k is decided
                const int k = __builtin_ctzll(amb);
                const uint32_t ik = i - (uint32_t)__builtin_popcountll(sure & ((1ull << k) - 1ull));
                if ((in.at(k) & mask) <= ik) sure |= 1ull << k;
                amb &= amb - 1;
            }
            uint32_t total = (uint32_t)__builtin_popcountll(sure);
            const uint32_t room = i - lo + 1;  // accepts left in this octave
            int used = 64;
            if (__builtin_expect(total >= room, 0)) {  // stop right after the room-th accepted lane: the outputs
                                                       // behind it meet a narrower mask (even the rejected ones)
                used = __builtin_ctzll(_pdep_u64(1ull << (room - 1), sure)) + 1;
                sure &= (used == 64) ? ~0ull : ((1ull << used) - 1ull);
                total = room;
            }
            if (STORE) {
                size_t tt = t;
#pragma GCC unroll 4
                for (int q = 0; q < 4; ++q) {
                    const __mmask16 sq = (__mmask16)(sure >> (16 * q));
                    const int cnt = __builtin_popcount((unsigned)sq);
                    store_compressed<J>(dst + tt, v[q], sq, cnt);
                    tt += (size_t)cnt;
                }
            }
            t += total;
            i -= total;
            in.advance(used);
        }
    }
    for (; i >= 1; --i) {  // the tail, one output at a time
        const uint32_t mask = bitmask(i);
        uint32_t r;
        while ((r = (in.next() & mask)) > i) {
        }
        if (STORE) dst[t] = (J)r;
        ++t;
    }
}
#endif

#if PHET_X86
// Count-only scan of 16-bit outputs (the scanner thread of the parallel replay, classes <= 65536 cells): the same
// trips on 32 lanes per vector -- no widening, 4 compares + 2 mask merges per 64 outputs instead of 8 compares and
// a dozen mask moves -- 128 outputs per trip where an octave is wide enough for the extra uncertainty of i to cost
// little.  v + k saturates (vpaddusw), so the trips are only taken while i + 128 < 65535.
template <typename Src>
MT_AVX512 void count_draws16_avx512(Src &in, uint32_t n) {
    const __m512i lane32 = _mm512_set_epi16(31, 30, 29, 28, 27, 26, 25, 24, 23, 22, 21, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9,
                                            8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m512i lanek[4] = {lane32, _mm512_add_epi16(lane32, _mm512_set1_epi16(32)), _mm512_add_epi16(lane32, _mm512_set1_epi16(64)),
                              _mm512_add_epi16(lane32, _mm512_set1_epi16(96))};
    uint32_t i = n - 1;
    while (i >= kScalarBelow) {
        const uint32_t mask = bitmask(i), lo = (mask >> 1) + 1;
        const uint32_t floor_i = lo > kScalarBelow ? lo : kScalarBelow;
        if (i + 160u >= 65535u) {   // (only the first ~160 draws of a 65,536-cell class) one output at a time
            uint32_t r;
            while ((r = (in.next() & mask)) > i) {
            }
            --i;
            continue;
        }
        const __m512i vm = _mm512_set1_epi16((short)mask);
        if (mask >= 8191u) {
            // 128 outputs per trip, SOFTWARE-PIPELINED: the vector work of trip t + 1 does not wait for trip t to be
            // settled.  As soon as trip t is classified (sure / ambiguous lanes against its exact start i), the start of
            // trip t + 1 is known to lie in [i - #not-rejected, i - #sure] -- a window as wide as trip t has ambiguous
            // lanes, usually 0 -- so trip t + 1 is classified against those two bounds right away: v + k <= lower bound is
            // accepted whatever happens, v > upper bound is rejected whatever happens, the lanes in between are settled
            // in lane order with the exact start once trip t is done.  The chain from one trip to the next is then a
            // popcount and a subtraction instead of broadcast -> compare -> mask moves -> popcount.
            auto classify = [&](const uint16_t *q0, uint32_t ihi, uint32_t ilo, uint64_t &s0, uint64_t &s1, uint64_t &a0, uint64_t &a1) MT_AVX512 {
                const __m512i vhi = _mm512_set1_epi16((short)ihi), vlo = _mm512_set1_epi16((short)ilo);
                __mmask32 A[4], S[4];
#pragma GCC unroll 4
                for (int q = 0; q < 4; ++q) {
                    const __m512i vq = _mm512_and_si512(_mm512_loadu_si512(q0 + 32 * q), vm);
                    A[q] = _mm512_cmple_epu16_mask(vq, vhi);                                   // not rejected: v <= largest possible start
                    S[q] = _mm512_cmple_epu16_mask(_mm512_adds_epu16(vq, lanek[q]), vlo);      // accepted for sure: v + k <= smallest possible start
                }
                s0 = _cvtmask64_u64(_mm512_kunpackd(S[1], S[0]));
                s1 = _cvtmask64_u64(_mm512_kunpackd(S[3], S[2]));
                a0 = _cvtmask64_u64(_mm512_kunpackd(A[1], A[0])) & ~s0;
                a1 = _cvtmask64_u64(_mm512_kunpackd(A[3], A[2])) & ~s1;
            };
            uint64_t sure0 = 0, sure1 = 0, amb0 = 0, amb1 = 0;
            bool have = false;   // (sure, amb) hold the classification of the trip at the current position
            while (i >= floor_i) {
                in.need(256);
                const uint16_t *p = reinterpret_cast<const uint16_t *>(in.p);
                if (!have) classify(p, i, i, sure0, sure1, amb0, amb1);
                // the next trip, from the bounds on where it will start (i >= 8192 > 128: no wrap)
                const uint32_t n_sure = (uint32_t)(__builtin_popcountll(sure0) + __builtin_popcountll(sure1));
                const uint32_t n_open = (uint32_t)(__builtin_popcountll(amb0) + __builtin_popcountll(amb1));
                uint64_t nsure0, nsure1, namb0, namb1;
                classify(p + 128, i - n_sure, i - n_sure - n_open, nsure0, nsure1, namb0, namb1);
                if (kScanPrefetch > 0) {   // the outputs were written by another core: pull the lines of a later trip closer
                    _mm_prefetch(reinterpret_cast<const char *>(p) + kScanPrefetch, _MM_HINT_T0);
                    _mm_prefetch(reinterpret_cast<const char *>(p) + kScanPrefetch + 64, _MM_HINT_T0);
                    _mm_prefetch(reinterpret_cast<const char *>(p) + kScanPrefetch + 128, _MM_HINT_T0);
                    _mm_prefetch(reinterpret_cast<const char *>(p) + kScanPrefetch + 192, _MM_HINT_T0);
                }
                while (__builtin_expect(amb0 != 0, 0)) {   // in lane order: everything before lane k is decided
                    const int k = __builtin_ctzll(amb0);
                    const uint32_t ik = i - (uint32_t)__builtin_popcountll(sure0 & ((1ull << k) - 1ull));
                    if (((uint32_t)p[k] & mask) <= ik) sure0 |= 1ull << k;
                    amb0 &= amb0 - 1;
                }
                const uint32_t t0 = (uint32_t)__builtin_popcountll(sure0);
                while (__builtin_expect(amb1 != 0, 0)) {
                    const int k = __builtin_ctzll(amb1);
                    const uint32_t ik = i - t0 - (uint32_t)__builtin_popcountll(sure1 & ((1ull << k) - 1ull));
                    if (((uint32_t)p[64 + k] & mask) <= ik) sure1 |= 1ull << k;
                    amb1 &= amb1 - 1;
                }
                uint32_t total = t0 + (uint32_t)__builtin_popcountll(sure1);
                const uint32_t room = i - lo + 1;   // accepts left in this octave
                int used = 128;
                if (__builtin_expect(total >= room, 0)) {   // stop right behind the room-th accepted lane
                    if (room <= t0) used = __builtin_ctzll(_pdep_u64(1ull << (room - 1), sure0)) + 1;
                    else used = 64 + __builtin_ctzll(_pdep_u64(1ull << (room - t0 - 1), sure1)) + 1;
                    total = room;
                }
                i -= total;
                in.advance(used);
                have = used == 128;   // (a trip cut short moves the next one: classify afresh)
                sure0 = nsure0;
                sure1 = nsure1;
                amb0 = namb0;
                amb1 = namb1;
            }
        } else {
            // narrow octaves: 64 outputs per trip, pipelined the same way
            auto classify64 = [&](const uint16_t *q0, uint32_t ihi, uint32_t ilo, uint64_t &s_, uint64_t &a_) MT_AVX512 {
                const __m512i vhi = _mm512_set1_epi16((short)ihi), vlo = _mm512_set1_epi16((short)ilo);
                __mmask32 A[2], S[2];
#pragma GCC unroll 2
                for (int q = 0; q < 2; ++q) {
                    const __m512i vq = _mm512_and_si512(_mm512_loadu_si512(q0 + 32 * q), vm);
                    A[q] = _mm512_cmple_epu16_mask(vq, vhi);
                    S[q] = _mm512_cmple_epu16_mask(_mm512_adds_epu16(vq, lanek[q]), vlo);
                }
                s_ = _cvtmask64_u64(_mm512_kunpackd(S[1], S[0]));
                a_ = _cvtmask64_u64(_mm512_kunpackd(A[1], A[0])) & ~s_;
            };
            uint64_t sure = 0, amb = 0;
            bool have = false;
            while (i >= floor_i) {
                in.need(128);
                const uint16_t *p = reinterpret_cast<const uint16_t *>(in.p);
                if (!have) classify64(p, i, i, sure, amb);
                const uint32_t n_sure = (uint32_t)__builtin_popcountll(sure), n_open = (uint32_t)__builtin_popcountll(amb);
                uint64_t nsure, namb;
                // (i >= 96 > 64 >= n_sure + n_open: no wrap; a lower bound below lane k just makes every lane ambiguous)
                classify64(p + 64, i - n_sure, i - n_sure - n_open, nsure, namb);
                while (__builtin_expect(amb != 0, 0)) {
                    const int k = __builtin_ctzll(amb);
                    const uint32_t ik = i - (uint32_t)__builtin_popcountll(sure & ((1ull << k) - 1ull));
                    if (((uint32_t)p[k] & mask) <= ik) sure |= 1ull << k;
                    amb &= amb - 1;
                }
                uint32_t total = (uint32_t)__builtin_popcountll(sure);
                const uint32_t room = i - lo + 1;
                int used = 64;
                if (__builtin_expect(total >= room, 0)) {
                    used = __builtin_ctzll(_pdep_u64(1ull << (room - 1), sure)) + 1;
                    total = room;
                }
                i -= total;
                in.advance(used);
                have = used == 64;
                sure = nsure;
                amb = namb;
            }
        }
    }
    for (; i >= 1; --i) {   // the tail, one output at a time
        const uint32_t mask = bitmask(i);
        uint32_t r;
        while ((r = (in.next() & mask)) > i) {
        }
    }
}
#endif

template <typename J, typename Src>
void shuffle_draws_from(Src &src, bool avx512, uint32_t n, J *dst) {
    if (n < 2) return;
#if PHET_X86
    if (avx512) {
        if (dst) shuffle_draws_avx512<J, true>(src, n, dst);
        else shuffle_draws_avx512<J, false>(src, n, dst);
        return;
    }
#endif
    if (dst) shuffle_draws_scalar<J, true>(src, n, dst);
    else shuffle_draws_scalar<J, false>(src, n, dst);
}

template <typename J>
void shuffle_draws(phet_mt &e, uint32_t n, J *dst) {
    SeqSrc src{e};
    shuffle_draws_from<J>(src, e.avx512, n, dst);
}

phet_mt *mt_alloc() {
    void *mem = nullptr;
    if (posix_memalign(&mem, 64, sizeof(phet_mt)) != 0 || !mem) return nullptr;
    phet_mt *e = new (mem) phet_mt();
    e->avx512 = cpu_has_avx512();
    return e;
}

void mt_free(phet_mt *e) {
    if (e) {
        e->~phet_mt();
        free(e);
    }
}

inline void backoff(unsigned &spins) {
    cpu_relax();
    if ((++spins & 1023u) == 0) std::this_thread::yield();
}

struct PermDesc {
    uint32_t n;
    int32_t eb;
    void *dst;
    int64_t ticket;
    int32_t pos;              // generator position at the start of this permutation ...
    uint32_t key[kN];         // ... and its state (filled by the scanner)
};

}  // namespace

// ---- parallel replay: one stream, three kinds of threads -------------------------------------------------------
//   generators  each walks the WHOLE stream with the twist alone (no tempering: ~1/3 of a generator's work) and
//               tempers + stores every G-th region of 64 blocks into a ring that stays cache-resident (16 bits per
//               output when every permutation has <= 65536 cells) -- no jump-ahead polynomial needed
//   scanner     the one job that is sequential by nature: walks the permutations in order with the count-only scan
//               (no generation, no compaction, no stores) and records where in the stream each one STARTS
//   workers     take permutations in order and redo the scan from the recorded start, this time compacting and
//               storing the accepted targets
// A fit is then bounded by the count-only scan of the stream (measured on the GPU box's Xeon: the ring moves
// 30 GB/s between two cores, the scan needs ~9 GB/s of 16-bit outputs).
constexpr int kRegionBlocks = 64;
constexpr int kSnapEvery = 4096;   // blocks between the generator states kept for the final (key, pos)

struct phet_mt_par {
    static constexpr size_t kQ = 1024;
    phet_mt *mt = nullptr;
    int64_t max_n = 0;
    bool use_ring = false;
    std::vector<PermDesc> desc = std::vector<PermDesc>(kQ);
    std::vector<uint64_t> start = std::vector<uint64_t>(kQ);   // ring mode: stream position where permutation k starts
    std::unique_ptr<std::atomic<uint8_t>[]> done{new std::atomic<uint8_t>[kQ]};
    alignas(64) std::atomic<uint64_t> submitted{0};
    alignas(64) std::atomic<uint64_t> started{0};
    alignas(64) std::atomic<uint64_t> claimed{0};
    alignas(64) std::atomic<uint64_t> finished_prefix{0};
    alignas(64) std::atomic<uint64_t> scan_pos{0};
    std::atomic<bool> closed{false};
    std::mutex mu, prefix_mu;
    std::condition_variable cv;
    std::vector<int64_t> ticket_left;
    std::thread tscan;
    std::vector<std::thread> workers, generators;
    // ring mode
    OutRing ring;
    uint32_t key0[kN];
    uint64_t abs0 = 0, end_pos = 0;
    std::vector<uint32_t> snaps;   // generator 0: its state at every kSnapEvery-th block (block 0 first)

    void publish_low() {  // outputs before the oldest unfinished permutation (or before the scanner) are dead
        const uint64_t f = finished_prefix.load(std::memory_order_acquire);
        const uint64_t st = started.load(std::memory_order_acquire);
        const uint64_t lo = f < st ? start[f & (kQ - 1)] : scan_pos.load(std::memory_order_acquire);
        uint64_t cur = ring.low.load(std::memory_order_relaxed);
        while (lo > cur && !ring.low.compare_exchange_weak(cur, lo, std::memory_order_release)) {
        }
    }

    template <typename E>
    void generator_main(int g, int G) {
        alignas(64) uint32_t key[kN + 32];
        alignas(64) uint32_t out[kN + 32];
        memcpy(key, key0, sizeof(key0));
        memset(key + kN, 0, sizeof(uint32_t) * 32);
        const bool avx = mt->avx512;
        uint64_t blk = 0;   // block `key` is the state of
        unsigned spins = 0;
        E *w = static_cast<E *>(ring.w);
        for (uint64_t reg = (uint64_t)g;; reg += (uint64_t)G) {
            // room: region `reg` overwrites the slot of region reg - nreg
            while (!ring.stop.load(std::memory_order_relaxed) &&
                   (reg + 1) * ring.region_elems > ring.low.load(std::memory_order_acquire) + ring.cap) {
                backoff(spins);
                if ((spins & 1023u) == 0) ring.writer_spins.fetch_add(1024, std::memory_order_relaxed);
            }
            if (ring.stop.load(std::memory_order_relaxed)) return;
            for (int b = 0; b < kRegionBlocks; ++b) {
                const uint64_t want = reg * kRegionBlocks + (uint64_t)b;
                while (blk < want) {   // run ahead with the twist alone
#if PHET_X86
                    if (avx) twist_block_avx512(key);
                    else
#endif
                        twist_scalar(key);
                    ++blk;
                    if (g == 0 && blk % kSnapEvery == 0) snaps.insert(snaps.end(), key, key + kN);
                }
                E *dst = w + (size_t)((want * kN) % ring.cap);
                if (sizeof(E) == 4) {
#if PHET_X86
                    if (avx) temper_avx512(key, reinterpret_cast<uint32_t *>(dst));
                    else
#endif
                        temper_scalar(key, reinterpret_cast<uint32_t *>(dst));
                } else {
#if PHET_X86
                    if (avx) temper16_avx512(key, reinterpret_cast<uint16_t *>(dst));
                    else
#endif
                    {
                        temper_scalar(key, out);
                        for (int i = 0; i < kN; ++i) dst[i] = (E)out[i];
                    }
                }
                if (dst == w) memcpy(w + ring.cap, w, sizeof(E) * kRingMirror);   // the wrap-around mirror
            }
            ring.ready[reg % ring.nreg].store(reg + 1, std::memory_order_release);
        }
    }

    template <typename E>
    void ring_scanner_main() {
        RingSrc<E> src(&ring, abs0);
        unsigned spins = 0;
        for (uint64_t k = 0;; ++k) {
            while (k >= submitted.load(std::memory_order_acquire)) {
                if (closed.load(std::memory_order_acquire) && k >= submitted.load(std::memory_order_acquire)) {
                    end_pos = src.abs;
                    return;
                }
                backoff(spins);
            }
            const PermDesc &d = desc[k & (kQ - 1)];
            start[k & (kQ - 1)] = src.abs;
            scan_pos.store(src.abs, std::memory_order_release);
            started.store(k + 1, std::memory_order_release);
            publish_low();
            count_only(src, d.n);
        }
    }

    void count_only(RingSrc<uint16_t> &src, uint32_t n) {
        if (n < 2) return;
#if PHET_X86
        if (mt->avx512) {
            count_draws16_avx512(src, n);
            return;
        }
#endif
        shuffle_draws_from<uint16_t>(src, false, n, (uint16_t *)nullptr);
    }
    void count_only(RingSrc<uint32_t> &src, uint32_t n) { shuffle_draws_from<uint16_t>(src, mt->avx512, n, (uint16_t *)nullptr); }

    void scanner_main() {  // plain mode: scans on the caller's generator, records its state at each start
        unsigned spins = 0;
        for (uint64_t k = 0;; ++k) {
            while (k >= submitted.load(std::memory_order_acquire)) {
                if (closed.load(std::memory_order_acquire) && k >= submitted.load(std::memory_order_acquire)) return;
                backoff(spins);
            }
            PermDesc &d = desc[k & (kQ - 1)];
            if (d.dst) {
                int p = 0;
                mt->state(d.key, &p);
                d.pos = p;
            }
            started.store(k + 1, std::memory_order_release);
            shuffle_draws<uint16_t>(*mt, d.n, (uint16_t *)nullptr);
        }
    }

    void finish(uint64_t k, int64_t ticket) {
        done[k & (kQ - 1)].store(1, std::memory_order_release);
        {
            std::lock_guard<std::mutex> lk(prefix_mu);
            uint64_t f = finished_prefix.load(std::memory_order_relaxed);
            while (done[f & (kQ - 1)].load(std::memory_order_acquire)) {
                done[f & (kQ - 1)].store(0, std::memory_order_relaxed);
                ++f;
            }
            finished_prefix.store(f, std::memory_order_release);
        }
        if (use_ring) publish_low();
        bool last;
        {
            std::lock_guard<std::mutex> lk(mu);
            last = --ticket_left[(size_t)ticket] == 0;
        }
        if (last) cv.notify_all();
    }

    template <typename E>
    void worker_main() {
        phet_mt *local = use_ring ? nullptr : mt_alloc();
        unsigned spins = 0;
        for (;;) {
            uint64_t k = claimed.load(std::memory_order_relaxed);
            if (k >= started.load(std::memory_order_acquire)) {
                if (closed.load(std::memory_order_acquire) && k >= submitted.load(std::memory_order_acquire)) break;
                backoff(spins);
                continue;
            }
            if (!claimed.compare_exchange_weak(k, k + 1, std::memory_order_acq_rel)) continue;
            const PermDesc &d = desc[k & (kQ - 1)];
            const int64_t ticket = d.ticket;
            if (d.dst && d.n >= 2) {
                if (use_ring) {
                    RingSrc<E> src(&ring, start[k & (kQ - 1)]);
                    if (d.eb == 2) shuffle_draws_from<uint16_t>(src, mt->avx512, d.n, static_cast<uint16_t *>(d.dst));
                    else shuffle_draws_from<uint32_t>(src, mt->avx512, d.n, static_cast<uint32_t *>(d.dst));
                } else if (local) {
                    local->load(d.key, d.pos);
                    if (d.eb == 2) shuffle_draws<uint16_t>(*local, d.n, static_cast<uint16_t *>(d.dst));
                    else shuffle_draws<uint32_t>(*local, d.n, static_cast<uint32_t *>(d.dst));
                }
            }
            finish(k, ticket);
        }
        mt_free(local);
    }
};

namespace {

#define MT_REQUIRE(cond, msg)                                    \
    do {                                                         \
        if (!(cond)) {                                           \
            ::phet::set_error(std::string("invalid: ") + (msg)); \
            return PHET_ERR_INVALID;                             \
        }                                                        \
    } while (0)

int check_jobs(const phet_mt_job *jobs, int32_t n_jobs, int64_t max_n) {
    for (int k = 0; k < n_jobs; ++k) {
        MT_REQUIRE(jobs[k].n >= 1 && jobs[k].n <= max_n, "permutation length out of range");
        MT_REQUIRE(jobs[k].elem_bytes == 2 || jobs[k].elem_bytes == 4, "elem_bytes must be 2 or 4");
        MT_REQUIRE(jobs[k].elem_bytes == 4 || jobs[k].n <= 65536, "16-bit draws need n <= 65536");
        MT_REQUIRE(jobs[k].out == nullptr || jobs[k].stride_bytes >= (jobs[k].n - 1) * jobs[k].elem_bytes,
                   "stride shorter than one permutation's draws");
    }
    return PHET_OK;
}

}  // namespace

extern "C" {

int phet_mt_create(const uint32_t *key, int32_t pos, phet_mt **out) {
    MT_REQUIRE(key != nullptr && out != nullptr, "NULL argument");
    MT_REQUIRE(pos >= 0 && pos <= kN, "MT19937 position must be in [0, 624]");
    phet_mt *e = mt_alloc();
    if (!e) {
        phet::set_error("out of host memory");
        return PHET_ERR_NOMEM;
    }
    e->load(key, pos);
    *out = e;
    return PHET_OK;
}

int phet_mt_state(const phet_mt *e, uint32_t *key_out, int32_t *pos_out) {
    MT_REQUIRE(e != nullptr && key_out != nullptr && pos_out != nullptr, "NULL argument");
    int p = 0;
    e->state(key_out, &p);
    *pos_out = p;
    return PHET_OK;
}

int phet_mt_destroy(phet_mt *e) {
    mt_free(e);
    return PHET_OK;
}

int phet_mt_simd(const phet_mt *e) { return e ? (e->avx512 ? 512 : 0) : (cpu_has_avx512() ? 512 : 0); }

int phet_mt_shuffle_rounds(phet_mt *e, const phet_mt_job *jobs, int32_t n_jobs, int64_t rounds) {
    MT_REQUIRE(e != nullptr && jobs != nullptr && n_jobs >= 1 && rounds >= 0, "bad arguments");
    if (int rc = check_jobs(jobs, n_jobs, ((int64_t)1 << 31) - 1)) return rc;
    for (int64_t r = 0; r < rounds; ++r) {
        for (int k = 0; k < n_jobs; ++k) {
            const phet_mt_job &j = jobs[k];
            unsigned char *dst = j.out ? static_cast<unsigned char *>(j.out) + (size_t)r * (size_t)j.stride_bytes : nullptr;
            if (j.elem_bytes == 2) shuffle_draws<uint16_t>(*e, (uint32_t)j.n, reinterpret_cast<uint16_t *>(dst));
            else shuffle_draws<uint32_t>(*e, (uint32_t)j.n, reinterpret_cast<uint32_t *>(dst));
        }
    }
    return PHET_OK;
}

int phet_mt_bounded(phet_mt *e, uint32_t max, int64_t count, uint32_t *out) {
    MT_REQUIRE(e != nullptr && out != nullptr && count >= 0, "bad arguments");
    if (max == 0) {  // random_interval(0) returns 0 without touching the stream
        for (int64_t k = 0; k < count; ++k) out[k] = 0;
        return PHET_OK;
    }
    const uint32_t mask = bitmask(max);
    for (int64_t k = 0; k < count; ++k) {
        uint32_t v;
        while ((v = (e->next32() & mask)) > max) {
        }
        out[k] = v;
    }
    return PHET_OK;
}

int phet_host_fisher_yates(const void *draws, int32_t elem_bytes, int64_t n, int64_t take, uint32_t *out) {
    MT_REQUIRE(out != nullptr && n >= 1 && take >= 0 && take <= n, "bad arguments");
    MT_REQUIRE(n == 1 || draws != nullptr, "draws are NULL");
    MT_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "elem_bytes must be 2 or 4");
    uint32_t *x = static_cast<uint32_t *>(malloc(sizeof(uint32_t) * (size_t)n));
    if (!x) {
        phet::set_error("out of host memory");
        return PHET_ERR_NOMEM;
    }
    for (int64_t q = 0; q < n; ++q) x[q] = (uint32_t)q;
    for (int64_t t = 0; t + 1 < n; ++t) {
        const int64_t i = n - 1 - t;
        const uint32_t j = elem_bytes == 2 ? static_cast<const uint16_t *>(draws)[t] : static_cast<const uint32_t *>(draws)[t];
        if ((int64_t)j > i) {
            free(x);
            phet::set_error("invalid: a draw exceeds its position (not a rk_interval stream)");
            return PHET_ERR_INVALID;
        }
        std::swap(x[i], x[j]);
    }
    memcpy(out, x, sizeof(uint32_t) * (size_t)take);
    free(x);
    return PHET_OK;
}

// ---- parallel replay ------------------------------------------------------------------------------------------
int phet_mt_par_begin(phet_mt *mt, int32_t workers, int64_t max_n, phet_mt_par **out) {
    MT_REQUIRE(mt != nullptr && out != nullptr, "NULL argument");
    MT_REQUIRE(max_n >= 1 && max_n < ((int64_t)1 << 31), "max_n must be in [1, 2^31)");
    *out = nullptr;
    unsigned hw = std::thread::hardware_concurrency();
    // several ranks of one job replay the same stream side by side on one host (one process per GPU): each takes its
    // share of the cores.  All these threads spin, so oversubscription is what hurts: measured on the 8-GPU box (32
    // cores), 8 replays with 4 + 1 + 4 threads each took 0.5 - 1.3 s instead of 0.45 s.  PHET_REPLAY_PROCS, else
    // torchrun's LOCAL_WORLD_SIZE, says how many processes share the host.
    int procs = 1;
    if (const char *e = getenv("PHET_REPLAY_PROCS")) procs = atoi(e) > 0 ? atoi(e) : 1;
    else if (const char *e2 = getenv("LOCAL_WORLD_SIZE")) procs = atoi(e2) > 0 ? atoi(e2) : 1;
    const unsigned budget = hw ? (hw / (unsigned)procs > 0 ? hw / (unsigned)procs : 1u) : 0u;
    // 3 generators + scanner + 3 workers when the cores are there (measured end to end at config 4 on the 16-core
    // 1-GPU box: 0.36 s; 4 + 1 + 4: 0.39 s; 2 + 1 + 2: 0.39 s); 2 + 1 + 1 on a 4-core share (a rank of 8 stores 1/8 only)
    int gens = 3;
    if (budget && budget < 8) gens = budget >= 4 ? 2 : (budget >= 3 ? 1 : 0);
    if (const char *w = getenv("PHET_REPLAY_GENERATORS")) gens = atoi(w) >= 0 ? atoi(w) : gens;
    if (workers <= 0) {
        workers = 3;
        if (budget && budget < 8) workers = (int)budget - gens - 1 >= 1 ? (int)budget - gens - 1 : 1;
        if (workers > 3) workers = 3;
        if (const char *w = getenv("PHET_REPLAY_WORKERS")) workers = atoi(w) > 0 ? atoi(w) : workers;
    }
    if (budget) hw = budget;
    if (hw && (unsigned)(workers + gens + 1) > hw) {   // not enough cores for the ring pipeline: scanner generates for itself
        gens = 0;
        if ((unsigned)workers + 1 > hw) workers = hw > 1 ? (int)hw - 1 : 1;
    }
    if (max_n < 2048) gens = 0;   // short permutations: the stream is consumed faster than threads can hand it over
    phet_mt_par *p = new (std::nothrow) phet_mt_par();
    if (!p) {
        phet::set_error("out of host memory");
        return PHET_ERR_NOMEM;
    }
    p->mt = mt;
    p->max_n = max_n;
    p->use_ring = gens > 0;
    for (size_t i = 0; i < phet_mt_par::kQ; ++i) p->done[i].store(0, std::memory_order_relaxed);
    const bool narrow = max_n <= 65536;
    if (p->use_ring) {
        OutRing &r = p->ring;
        r.ebytes = narrow ? 2 : 4;
        r.region_elems = (size_t)kRegionBlocks * kN;
        size_t ring_mb = 8;
        if (const char *e = getenv("PHET_REPLAY_RING_MB")) ring_mb = atoi(e) > 0 ? (size_t)atoi(e) : 8;
        size_t elems = (ring_mb << 20) / (size_t)r.ebytes;
        const size_t need = (size_t)max_n * 12;   // several permutations' outputs (~1.4 n each) in flight
        if (elems < need) elems = need;
        r.nreg = (elems + r.region_elems - 1) / r.region_elems;
        if (r.nreg < (size_t)(2 * gens + 2)) r.nreg = (size_t)(2 * gens + 2);
        r.cap = r.nreg * r.region_elems;
        void *mem = nullptr;
        if (posix_memalign(&mem, 64, (r.cap + kRingMirror + 64) * (size_t)r.ebytes) != 0 || !mem) {
            delete p;
            phet::set_error("out of host memory");
            return PHET_ERR_NOMEM;
        }
        r.w = mem;
        r.ready.reset(new std::atomic<uint64_t>[r.nreg]);
        for (size_t i = 0; i < r.nreg; ++i) r.ready[i].store(0, std::memory_order_relaxed);
        int pos = 0;
        mt->state(p->key0, &pos);
        p->abs0 = (uint64_t)pos;
        p->snaps.assign(p->key0, p->key0 + kN);
        p->scan_pos.store(p->abs0, std::memory_order_release);
        for (int g = 0; g < gens; ++g) {
            if (narrow) p->generators.emplace_back(&phet_mt_par::generator_main<uint16_t>, p, g, gens);
            else p->generators.emplace_back(&phet_mt_par::generator_main<uint32_t>, p, g, gens);
        }
        if (narrow) p->tscan = std::thread(&phet_mt_par::ring_scanner_main<uint16_t>, p);
        else p->tscan = std::thread(&phet_mt_par::ring_scanner_main<uint32_t>, p);
    } else {
        p->tscan = std::thread(&phet_mt_par::scanner_main, p);
    }
    for (int w = 0; w < workers; ++w) {
        if (narrow) p->workers.emplace_back(&phet_mt_par::worker_main<uint16_t>, p);
        else p->workers.emplace_back(&phet_mt_par::worker_main<uint32_t>, p);
    }
    *out = p;
    return PHET_OK;
}

int phet_mt_par_threads(const phet_mt_par *p, int32_t out[3]) {
    MT_REQUIRE(p != nullptr && out != nullptr, "NULL argument");
    out[0] = (int32_t)p->generators.size();
    out[1] = 1;
    out[2] = (int32_t)p->workers.size();
    return PHET_OK;
}

int phet_mt_par_submit(phet_mt_par *p, const phet_mt_job *jobs, int32_t n_jobs, int64_t rounds, int64_t *ticket_out) {
    MT_REQUIRE(p != nullptr && jobs != nullptr && n_jobs >= 1 && rounds >= 0 && ticket_out != nullptr, "bad arguments");
    MT_REQUIRE(!p->closed.load(), "the replay is already closed");

    if (int rc = check_jobs(jobs, n_jobs, p->max_n)) return rc;
    int64_t ticket;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        ticket = (int64_t)p->ticket_left.size();
        p->ticket_left.push_back(rounds * n_jobs);
    }
    *ticket_out = ticket;
    uint64_t k = p->submitted.load(std::memory_order_relaxed);
    unsigned spins = 0;
    for (int64_t r = 0; r < rounds; ++r) {
        for (int j = 0; j < n_jobs; ++j) {
            while (k - p->finished_prefix.load(std::memory_order_acquire) >= phet_mt_par::kQ - 1) backoff(spins);
            PermDesc &d = p->desc[k & (phet_mt_par::kQ - 1)];
            d.n = (uint32_t)jobs[j].n;
            d.eb = jobs[j].elem_bytes;
            d.dst = jobs[j].out ? static_cast<unsigned char *>(jobs[j].out) + (size_t)r * (size_t)jobs[j].stride_bytes : nullptr;
            d.ticket = ticket;
            p->submitted.store(++k, std::memory_order_release);
        }
    }
    return PHET_OK;
}

int phet_mt_par_wait(phet_mt_par *p, int64_t ticket) {
    MT_REQUIRE(p != nullptr, "NULL argument");
    std::unique_lock<std::mutex> lk(p->mu);
    MT_REQUIRE(ticket >= 0 && (size_t)ticket < p->ticket_left.size(), "unknown ticket");
    p->cv.wait(lk, [&] { return p->ticket_left[(size_t)ticket] == 0; });
    return PHET_OK;
}

int phet_mt_par_end(phet_mt_par *p) {
    if (!p) return PHET_OK;
    p->closed.store(true, std::memory_order_release);
    p->tscan.join();  // plain mode: the caller's generator is now where the reference's stream would be
    for (std::thread &w : p->workers) w.join();
    if (p->use_ring) {
        p->ring.stop.store(true, std::memory_order_release);
        for (std::thread &g : p->generators) g.join();
        // the stream position the scan ended at -> NumPy's (key, pos), re-twisted from the nearest kept state
        if (p->end_pos != p->abs0) {
            uint64_t block = p->end_pos / kN;
            int off = (int)(p->end_pos % kN);
            if (off == 0) {  // NumPy regenerates lazily: the last block fully consumed, pos = 624
                --block;
                off = kN;
            }
            alignas(64) uint32_t key[kN + 32];
            uint64_t s0 = block / kSnapEvery;
            const uint64_t have = p->snaps.size() / kN;   // generator 0 may not have reached the last multiple
            if (s0 >= have) s0 = have - 1;
            memcpy(key, p->snaps.data() + (size_t)s0 * kN, sizeof(uint32_t) * kN);
            memset(key + kN, 0, sizeof(uint32_t) * 32);
            for (uint64_t b2 = s0 * kSnapEvery; b2 < block; ++b2) {
#if PHET_X86
                if (p->mt->avx512) twist_block_avx512(key);
                else
#endif
                    twist_scalar(key);
            }
            p->mt->load(key, off);
        }
        if (getenv("PHET_REPLAY_DEBUG"))
            fprintf(stderr, "replay: readers spun %llu (starved by the generators), generators spun %llu (ring full)\n",
                    (unsigned long long)p->ring.reader_spins.load(), (unsigned long long)p->ring.writer_spins.load());
        free(p->ring.w);
    }
    delete p;
    return PHET_OK;
}

}  // extern "C"
