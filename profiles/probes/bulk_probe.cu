// bulk_probe.cu — how fast can cp.async.bulk (UBLKCP) alone stream HBM into shared memory?
// One thread per CTA keeps `stages` bulk copies of `chunk` bytes in flight over its share of a big buffer.
// Also a plain LDG.128 reader for comparison.   nvcc -arch=sm_100a -O3 -o bulk_probe bulk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../image_processor_pipeline_b200/csrc/staging.cuh"
namespace ipp { void set_error(const char*, ...) {} int sm_count() { return 148; } LaunchTimer::LaunchTimer(const char*, cudaStream_t) {} LaunchTimer::~LaunchTimer() {} }
using namespace ipp;

__global__ void bulk_kernel(const uint8_t* src, size_t total, uint32_t chunk, int stages) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
        const size_t nchunks = total / chunk;
        int it = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
            const int s = it % stages, use = it / stages;
            if (use > 0) mbar_wait(&bars[s], (uint32_t)(use - 1) & 1u);
            mbar_expect_tx(&bars[s], chunk);
            bulk_g2s(smem + (size_t)s * chunk, src + c * chunk, chunk, &bars[s]);
        }
        for (int s = 0; s < stages; ++s) {          // drain
            const int uses = (it - s + stages - 1) / stages;
            if (uses > 0) mbar_wait(&bars[s], (uint32_t)(uses - 1) & 1u);
        }
    }
}

__global__ void ldg_kernel(const uint4* src, size_t n16, uint32_t* sink) {
    uint32_t acc = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        const uint4 a = ldg_stream16(src + i), b = ldg_stream16(src + i + stride), c = ldg_stream16(src + i + 2 * stride), d = ldg_stream16(src + i + 3 * stride);
        acc |= a.x | b.y | c.z | d.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main() {
    const size_t total = (size_t)1 << 31;     // 2 GiB
    uint8_t* buf; uint32_t* sink;
    cudaMalloc(&buf, total); cudaMalloc(&sink, 4); cudaMemset(buf, 1, total);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto launch, const char* name) {
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-60s %8.3f ms  %8.1f GB/s  (%s)\n", name, ms, total / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    char name[128];
    for (int grid : {148, 296, 592})
        for (uint32_t chunk : {5760u, 16384u, 46080u})
            for (int stages : {2, 3, 4, 8}) {
                const size_t smem = (size_t)chunk * stages;
                if (smem * (grid / 148) > 200 * 1024) continue;
                snprintf(name, sizeof(name), "bulk grid=%d chunk=%u stages=%d (%.0f KB in flight/SM)", grid, chunk, stages, smem * (grid / 148) / 1024.0);
                timeit([&] { bulk_kernel<<<grid, 32, smem>>>(buf, total, chunk, stages); }, name);
            }
    for (int mult : {4, 8, 16})
        { snprintf(name, sizeof(name), "LDG.128 x4 unrolled, grid=%d x 256", 148 * mult);
          timeit([&] { ldg_kernel<<<148 * mult, 256>>>((const uint4*)buf, total / 16, sink); }, name); }
    return 0;
}
