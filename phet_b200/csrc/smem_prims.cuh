// Explicit shared-window accesses and the monotone bucket map shared by the thread-per-sample kernels
// (k_step1p, k_step3p): their scratch areas are addressed by byte arithmetic on 32-bit shared addresses.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace phet {

// ---- explicit shared-window accesses (the scratch areas are addressed by byte arithmetic) ----
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v));
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w));
}
__device__ __forceinline__ void sts_val(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v));
}
__device__ __forceinline__ void sts_val(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v));
}
__device__ __forceinline__ float lds_val(uint32_t addr, float) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_val(uint32_t addr, double) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t hist_addr(uint32_t hbase, uint32_t b) {
    // word b/4 of this lane (words interleaved by lane, 128 bytes apart), byte b%4:
    // 128*(b/4) + b%4 == 32*b - 31*(b%4); two IMADs (FMA pipe) and one LOP3
    return (b & 3u) * 0xFFFFFFE1u + (b * 32u + hbase);
}

// Bucket of v and the shared-memory address of its counter.  t = v*scale + off, clamped to
// top = 4*NBW - 1.5, is monotone non-decreasing in v (fma with scale >= 0, min, truncation and the
// saturation of the conversion all are); below the window and NaN -> 0.  The counter of bucket b is
// byte b%4 of word b/4, and the words of a thread sit 128 bytes apart (bank == lane: conflict-free).
// (One conversion only: F2I runs on the quarter-rate XU pipe, which the fp64 conversions also use.)
__device__ __forceinline__ uint32_t pair_bucket(float v, float scale, float off, float top, uint32_t hbase,
                                                uint32_t &addr) {
    const float t = fminf(fmaf(v, scale, off), top);
    uint32_t b;
    asm("cvt.rzi.u32.f32 %0, %1;" : "=r"(b) : "f"(t));
    addr = hist_addr(hbase, b);
    return b;
}

// vals[pos] with the address formed by one IMAD (FMA pipe) instead of LEA (ALU pipe)
__device__ __forceinline__ float lds_pos(uint32_t base, uint32_t pos, float) {
    float v;
    asm volatile("{\n.reg .u32 a;\nmad.lo.u32 a, %1, 4, %2;\nld.shared.f32 %0, [a];\n}" : "=f"(v) : "r"(pos), "r"(base));
    return v;
}
__device__ __forceinline__ double lds_pos(uint32_t base, uint32_t pos, double) {
    double v;
    asm volatile("{\n.reg .u32 a;\nmad.lo.u32 a, %1, 8, %2;\nld.shared.f64 %0, [a];\n}" : "=d"(v) : "r"(pos), "r"(base));
    return v;
}


}  // namespace phet
