// Issue-rate probe: legacy mma.sync IMMA (m16n8k32 u8/s8 -> s32) vs HMMA (m16n8k16 f16 -> f32) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) probe(int iters, int* out) {
    int c[8][4];
    float f[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) { c[j][i] = threadIdx.x + i; f[j][i] = (float)(threadIdx.x + i); }
    uint32_t a0 = threadIdx.x * 0x01010101u, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 ^ 0x5a5a5a5au, b1 = b0 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {            // 8 independent accumulators per warp
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[j][0]), "+f"(f[j][1]), "+f"(f[j][2]), "+f"(f[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 3)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(b0));
            else if (KIND == 4)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(f[j][0]), "+f"(f[j][1]), "+f"(f[j][2]), "+f"(f[j][3]) : "r"(a0), "r"(a1), "r"(b0));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[j][0]), "+f"(f[j][1]), "+f"(f[j][2]), "+f"(f[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    int s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[j][i] + (int)f[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, double macs_per_mma) {
    int* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(int));
    const int iters = 4096, grid = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<KIND><<<grid, 256>>>(16, out);
    cudaEventRecord(e0);
    probe<KIND><<<grid, 256>>>(iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)grid * 8 * iters * 8;        // warps x iters x 8
    printf("%s: %.3f ms, %.1f G warp-MMA/s, %.2f cycles per MMA per SMSP @1.965 GHz, %.1f T MAC/s\n", name, ms, mmas / ms / 1e6,
           1.965e9 * 592 / (mmas / (ms * 1e-3)), mmas * macs_per_mma / ms / 1e9);
    cudaFree(out);
}

int main() {
    run<0>("IMMA m16n8k32 u8.s8", 16.0 * 8 * 32);
    run<1>("HMMA m16n8k16 f16  ", 16.0 * 8 * 16);
    run<2>("HMMA m16n8k16 bf16 ", 16.0 * 8 * 16);
    run<3>("IMMA m16n8k16 u8.s8", 16.0 * 8 * 16);
    run<4>("HMMA m16n8k8 f16   ", 16.0 * 8 * 8);
    return 0;
}
