// staging.cuh — row staging global -> shared with the bulk async-copy engine
// (cp.async.bulk, SASS UBLKCP) + mbarrier completion, and unaligned 16-byte
// extraction from the staged window.
//
// Why: interleaved 3-byte pixels, arbitrary crop offsets (x0*3 bytes) and pitches
// make source rows start at any byte alignment.  The bulk copy engine moves the
// 16-byte-aligned window that covers the wanted bytes into shared memory without
// occupying load/store instruction slots or registers; threads then cut aligned
// 16-byte output vectors out of the window with two LDS.128 + a byte funnel shift.
#pragma once
#include "common.cuh"

namespace ipp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}

// global -> shared bulk copy; src and dst 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// The 16-byte-aligned window covering global bytes [p, p + nbytes).
struct Window {
    const uint8_t* base;  // p rounded down to 16
    uint32_t bytes;       // multiple of 16
    uint32_t shift;       // p - base, 0..15
};
__device__ __forceinline__ Window make_window(const uint8_t* p, uint32_t nbytes) {
    Window w;
    const uintptr_t a = (uintptr_t)p;
    const uintptr_t lo = a & ~(uintptr_t)15;
    const uintptr_t hi = (a + nbytes + 15) & ~(uintptr_t)15;
    w.base = (const uint8_t*)lo;
    w.bytes = (uint32_t)(hi - lo);
    w.shift = (uint32_t)(a - lo);
    return w;
}

// 16 bytes starting at byte offset `off` (any alignment) of a 16-byte-aligned
// shared-memory window.  Reads the two aligned vectors that cover it; the window
// must therefore be padded by 16 bytes (stage with smem row stride >= bytes + 16).
__device__ __forceinline__ uint4 window_load16(const uint8_t* win, uint32_t off) {
    const uint32_t a = off & ~15u;
    const uint4 lo = *reinterpret_cast<const uint4*>(win + a);
    const uint32_t sh = off & 15u;
    if (sh == 0) return lo;
    const uint4 hi = *reinterpret_cast<const uint4*>(win + a + 16);
    const uint32_t bs = (sh & 3u) * 8u;
    uint32_t v0, v1, v2, v3, v4;
    switch (sh >> 2) {
        case 0:  v0 = lo.x; v1 = lo.y; v2 = lo.z; v3 = lo.w; v4 = hi.x; break;
        case 1:  v0 = lo.y; v1 = lo.z; v2 = lo.w; v3 = hi.x; v4 = hi.y; break;
        case 2:  v0 = lo.z; v1 = lo.w; v2 = hi.x; v3 = hi.y; v4 = hi.z; break;
        default: v0 = lo.w; v1 = hi.x; v2 = hi.y; v3 = hi.z; v4 = hi.w; break;
    }
    uint4 r;
    r.x = __funnelshift_r(v0, v1, bs);
    r.y = __funnelshift_r(v1, v2, bs);
    r.z = __funnelshift_r(v2, v3, bs);
    r.w = __funnelshift_r(v3, v4, bs);
    return r;
}

// Shared-memory stride of one staged row of `rowbytes` payload bytes: worst-case window
// (15 bytes of leading misalignment, rounded up to 16) plus slack so that
// window_load16() of any offset inside the row may read its second vector.
__host__ __device__ inline uint32_t staged_row_stride(uint32_t rowbytes) {
    return ((rowbytes + 30u) & ~15u) + 48u;
}

// Split a destination row of `rowbytes` bytes starting at address d into
// head (< 16 unaligned bytes), nb aligned 16-byte chunks, tail (< 16 bytes).
struct RowSplit {
    uint32_t head, nb, tail;
};
__device__ __forceinline__ RowSplit split_row(const uint8_t* d, uint32_t rowbytes) {
    RowSplit s;
    uint32_t head = (uint32_t)((16 - ((uintptr_t)d & 15)) & 15);
    if (head > rowbytes) head = rowbytes;
    s.head = head;
    s.nb = (rowbytes - head) >> 4;
    s.tail = rowbytes - head - (s.nb << 4);
    return s;
}

}  // namespace ipp
