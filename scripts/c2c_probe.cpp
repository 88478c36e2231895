// Probe: how fast can one core stream data that another core has just written (through L2 -> L3)?  Sizing input for a
// generator-thread / scanner-thread split of the np.random replay.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
#include <thread>
#include <vector>
int main(int argc, char **argv) {
    const size_t ring_bytes = (argc > 1 ? atol(argv[1]) : 8) << 20;
    const size_t total = (size_t)4 << 30;  // bytes streamed
    const size_t chunk = 2496;             // one MT block of 16-bit outputs ... 32-bit: 2496 B
    uint8_t *ring = (uint8_t *)aligned_alloc(64, ring_bytes + 4096);
    memset(ring, 1, ring_bytes + 4096);
    std::atomic<uint64_t> wpos{0}, rpos{0};
    uint64_t sink = 0;
    auto t0 = std::chrono::steady_clock::now();
    std::thread prod([&] {
        uint64_t w = 0;
        __m512i v = _mm512_set1_epi32(7);
        while (w < total) {
            while (w + chunk - rpos.load(std::memory_order_acquire) > ring_bytes) _mm_pause();
            uint8_t *d = ring + (w % ring_bytes);
            for (size_t o = 0; o < chunk; o += 64) { v = _mm512_add_epi32(v, _mm512_set1_epi32(3)); _mm512_storeu_si512(d + o, v); }
            w += chunk;
            wpos.store(w, std::memory_order_release);
        }
    });
    std::thread cons([&] {
        uint64_t r = 0;
        __m512i acc = _mm512_setzero_si512();
        while (r < total) {
            uint64_t w;
            while ((w = wpos.load(std::memory_order_acquire)) < r + chunk) _mm_pause();
            while (r + chunk <= w) {
                const uint8_t *s = ring + (r % ring_bytes);
                for (size_t o = 0; o < chunk; o += 64) acc = _mm512_add_epi32(acc, _mm512_loadu_si512(s + o));
                r += chunk;
            }
            rpos.store(r, std::memory_order_release);
        }
        sink = _mm512_reduce_add_epi32(acc);
    });
    prod.join();
    cons.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("ring %zu MB: %.2f GB/s producer->consumer (%llu)\n", ring_bytes >> 20, total / dt / 1e9, (unsigned long long)sink);
}
