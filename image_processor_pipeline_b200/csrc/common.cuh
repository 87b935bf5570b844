// common.cuh — shared helpers for the sm_100a kernels of libippgpu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/ippgpu.h"

namespace ipp {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
int sm_count();

#define IPP_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            ipp::set_error(__VA_ARGS__);        \
            return (code);                      \
        }                                       \
    } while (0)

#define IPP_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t _e = (call);                                                    \
        if (_e != cudaSuccess) {                                                    \
            ipp::set_error("%s failed: %s", #call, cudaGetErrorString(_e));         \
            return (int)_e;                                                         \
        }                                                                           \
    } while (0)

#define IPP_LAUNCH_CHECK()                                                          \
    do {                                                                            \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) {                                                    \
            ipp::set_error("kernel launch failed: %s", cudaGetErrorString(_e));     \
            return (int)_e;                                                         \
        }                                                                           \
    } while (0)

// RAII hook around every kernel launch of the library: counts launches and, when
// ipp_timing_enable(1) was called, brackets the launch with CUDA events on the launching
// stream so that ipp_timing_report() can give per-kernel device time.
struct LaunchTimer {
    LaunchTimer(const char* name, cudaStream_t st);
    ~LaunchTimer();
    const char* name_;
    cudaStream_t st_;
    int slot_;
};

// output tile of rotate_batch_kernel (batch.cu); the host planner (planner.cu) counts the same tiles
#ifndef IPP_ROT_TILE_H
#define IPP_ROT_TILE_H 256
#endif
constexpr int kRotTileW = 64, kRotTileH = IPP_ROT_TILE_H;

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- device helpers
// Streaming global accesses: data is touched once, keep it out of L1.
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream16(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t ldg_stream4(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// libImaging DIV255 / MULDIV255: t += 128; ((t >> 8) + t) >> 8
__device__ __forceinline__ uint32_t div255(uint32_t t) {
    t += 128u;
    return ((t >> 8) + t) >> 8;
}

// warp reductions
__device__ __forceinline__ int warp_min(int v) {
    return __reduce_min_sync(0xffffffffu, v);
}
__device__ __forceinline__ int warp_max(int v) {
    return __reduce_max_sync(0xffffffffu, v);
}

}  // namespace ipp
