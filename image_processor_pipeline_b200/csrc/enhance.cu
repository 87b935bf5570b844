// enhance.cu — photometric jitter of a batch of RGB images (reference transforms/tranfo.py:37-51, the last step
// of the dataset workflow, driven by the 'sample' pairing of pipeline.py:209-225).
//
// Reference behaviour replaced, per image (Pillow):
//   ImageEnhance.Brightness / Contrast / Color .enhance(f)   = ImagingBlend(degenerate, image, f)  (Blend.c: single-precision
//       a + f * (b - a), truncated to 8 bits, clipped when f is outside [0, 1]); degenerate = black / the mean gray
//       int(mean(L) + 0.5) / the gray image L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16
//   ImageFilter.GaussianBlur(r)  = ImagingGaussianBlur: three box passes per axis in 8.24 fixed point, every pass rounded to
//       8 bits (BoxBlur.c; host computes the box radius / weights exactly like _gaussian_blur_radius and
//       ImagingHorizontalBoxBlur do)
//   band.point(lambda)           = a 256-entry table per channel (the host draws and rounds it like Image.point)
// Everything is integer or single-precision arithmetic evaluated as libImaging does: bit-exact.
//
// Passes over a batch [n][h][w][3] (dense rows):
//   1. enh_luma_sum_kernel   one read: sum of L(brightness(px)) per image (the Contrast step needs the image mean)
//   2. enh_point_kernel      one read + one write: brightness -> contrast (one 256-entry table per image, built per CTA) ->
//                            colour (per pixel) -> channel tables when the image is not blurred
//   3. enh_blur_h / enh_blur_v   only for the images drawn for the blur: the three horizontal passes of a row group and the
//                            three vertical passes of a 64-byte column strip run from shared memory (one read + one write per
//                            axis instead of three); the channel tables are applied in the last store.
#include "common.cuh"

namespace ipp {

struct EnhParams {
    const uint8_t* src;
    uint8_t* dst;
    uint8_t* tmp;
    int64_t stride;            // bytes per image (dense: h * w * 3)
    int w, h;
    const ipp_enhance_item* items;
    const uint8_t* luts;       // [n][3][256]
    unsigned long long* sums;  // [n]
    const int32_t* blur_list;  // indices of the images to blur
};

__device__ __forceinline__ uint32_t blend_byte(float a, float b, float alpha, bool inside) {
    // (UINT8)(a + alpha * (b - a)) in single precision, no contraction; extrapolation is clipped first (Blend.c)
    const float t = __fadd_rn(a, __fmul_rn(alpha, __fsub_rn(b, a)));
    if (inside) return (uint32_t)(int)t;
    return t <= 0.0f ? 0u : (t >= 255.0f ? 255u : (uint32_t)(int)t);
}
__device__ __forceinline__ uint32_t luma(uint32_t r, uint32_t g, uint32_t b) {
    return (r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16;
}

// pixels 4t .. 4t+3 of a dense RGB image as 12 bytes (three aligned words when possible)
__device__ __forceinline__ void load_px4(const uint8_t* img, int64_t t, int64_t npx, uint8_t (&c)[12]) {
    const uint8_t* s = img + 12 * t;
    if (4 * t + 3 < npx && ((((uintptr_t)s) & 3) == 0)) {
        const uint32_t a = ldg_stream4(s), b = ldg_stream4(s + 4), d = ldg_stream4(s + 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) { c[k] = (uint8_t)(a >> (8 * k)); c[4 + k] = (uint8_t)(b >> (8 * k)); c[8 + k] = (uint8_t)(d >> (8 * k)); }
    } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) c[k] = (4 * t + k / 3 < npx) ? s[k] : (uint8_t)0;
    }
}

constexpr int kEnhCtas = 96;      // CTAs per image in the two streaming passes

__global__ void __launch_bounds__(256) enh_luma_sum_kernel(const EnhParams p) {
    __shared__ uint8_t lb[256];
    const int img = blockIdx.y;
    const ipp_enhance_item it = p.items[img];
    {
        const int c = threadIdx.x;
        lb[c] = (uint8_t)blend_byte(0.0f, (float)c, it.brightness, it.brightness >= 0.0f && it.brightness <= 1.0f);
    }
    __syncthreads();
    const int64_t npx = (int64_t)p.w * p.h, ngrp = (npx + 3) >> 2;
    const uint8_t* src = p.src + (int64_t)img * p.stride;
    unsigned long long sum = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ngrp; t += (int64_t)gridDim.x * blockDim.x) {
        uint8_t c[12];
        load_px4(src, t, npx, c);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (4 * t + k < npx) sum += luma(lb[c[3 * k]], lb[c[3 * k + 1]], lb[c[3 * k + 2]]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&p.sums[img], sum);
}

__global__ void __launch_bounds__(256) enh_point_kernel(const EnhParams p) {
    __shared__ uint8_t l1[256];            // contrast(brightness(c))
    __shared__ uint8_t lc[768];            // channel tables (only when they are applied here)
    const int img = blockIdx.y;
    const ipp_enhance_item it = p.items[img];
    const int64_t npx = (int64_t)p.w * p.h, ngrp = (npx + 3) >> 2;
    const bool lut_here = it.lut && !it.blur;
    {
        // Contrast: degenerate = int(mean(L) + 0.5), mean = sum / count as Python divides two ints (correctly rounded double)
        const double mean = (double)p.sums[img] / (double)npx;
        const float mg = (float)(int)(mean + 0.5);
        const int c = threadIdx.x;
        const uint32_t b = blend_byte(0.0f, (float)c, it.brightness, it.brightness >= 0.0f && it.brightness <= 1.0f);
        l1[c] = (uint8_t)blend_byte(mg, (float)b, it.contrast, it.contrast >= 0.0f && it.contrast <= 1.0f);
        if (lut_here)
            for (int i = c; i < 768; i += 256) lc[i] = p.luts[(int64_t)img * 768 + i];
    }
    __syncthreads();
    const bool col_in = it.color >= 0.0f && it.color <= 1.0f;
    const uint8_t* src = p.src + (int64_t)img * p.stride;
    uint8_t* dst = p.dst + (int64_t)img * p.stride;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ngrp; t += (int64_t)gridDim.x * blockDim.x) {
        uint8_t c[12];
        load_px4(src, t, npx, c);
        uint8_t o[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t r = l1[c[3 * k]], g = l1[c[3 * k + 1]], b = l1[c[3 * k + 2]];
            const float L = (float)luma(r, g, b);
            uint32_t r2 = blend_byte(L, (float)r, it.color, col_in), g2 = blend_byte(L, (float)g, it.color, col_in),
                     b2 = blend_byte(L, (float)b, it.color, col_in);
            if (lut_here) { r2 = lc[r2]; g2 = lc[256 + g2]; b2 = lc[512 + b2]; }
            o[3 * k] = (uint8_t)r2; o[3 * k + 1] = (uint8_t)g2; o[3 * k + 2] = (uint8_t)b2;
        }
        uint8_t* d = dst + 12 * t;
        if (4 * t + 3 < npx && ((((uintptr_t)d) & 3) == 0)) {
            uint32_t w[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) w[j] = o[4 * j] | (o[4 * j + 1] << 8) | (o[4 * j + 2] << 16) | ((uint32_t)o[4 * j + 3] << 24);
            reinterpret_cast<uint32_t*>(d)[0] = w[0]; reinterpret_cast<uint32_t*>(d)[1] = w[1]; reinterpret_cast<uint32_t*>(d)[2] = w[2];
        } else {
            for (int k = 0; k < 12 && 4 * t + k / 3 < npx; ++k) d[k] = o[k];
        }
    }
}

// one box pass along a line of `n` elements `step` bytes apart inside shared memory (ImagingLineBoxBlur: window sum over
// [x - r, x + r] with clamped edges times ww, the two far elements times fw, (v + 2^23) >> 24 in unsigned 32-bit arithmetic)
__device__ __forceinline__ uint8_t box_at(const uint8_t* line, int x, int n, int step, int r, uint32_t ww, uint32_t fw) {
    uint32_t acc = 0;
    for (int d = -r; d <= r; ++d) acc += line[min(max(x + d, 0), n - 1) * step];
    const uint32_t far = (uint32_t)line[min(max(x - r - 1, 0), n - 1) * step] + (uint32_t)line[min(max(x + r + 1, 0), n - 1) * step];
    return (uint8_t)((acc * ww + far * fw + (1u << 23)) >> 24);
}

// three horizontal passes of `rows` image rows per CTA: element = one byte of a pixel, neighbours 3 bytes apart
__global__ void __launch_bounds__(256) enh_blur_h_kernel(const EnhParams p, int rows) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int img = p.blur_list[blockIdx.y];
    const ipp_enhance_item it = p.items[img];
    const int rowbytes = p.w * 3;
    const int y0 = blockIdx.x * rows, nr = min(rows, p.h - y0);
    uint8_t* a = sm;
    uint8_t* b = sm + (size_t)rows * rowbytes;
    const uint8_t* src = p.dst + (int64_t)img * p.stride + (int64_t)y0 * rowbytes;      // the point pass wrote dst
    const int total = nr * rowbytes;
    for (int i = threadIdx.x; i < total; i += 256) a[i] = src[i];
    __syncthreads();
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = threadIdx.x; i < total; i += 256) {
            const int r = i / rowbytes, xb = i - r * rowbytes;
            const int x = xb / 3, ch = xb - 3 * x;
            b[i] = box_at(a + r * rowbytes + ch, x, p.w, 3, it.box_radius, it.ww, it.fw);
        }
        __syncthreads();
        uint8_t* t = a; a = b; b = t;
    }
    uint8_t* dst = p.tmp + (int64_t)blockIdx.y * p.stride + (int64_t)y0 * rowbytes;
    for (int i = threadIdx.x; i < total; i += 256) dst[i] = a[i];
}

// three vertical passes of a strip of `sw` bytes x all rows per CTA; channel tables in the last store
__global__ void __launch_bounds__(256) enh_blur_v_kernel(const EnhParams p, int sw) {
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ uint8_t lc[768];
    const int img = p.blur_list[blockIdx.y];
    const ipp_enhance_item it = p.items[img];
    const int rowbytes = p.w * 3;
    const int b0 = blockIdx.x * sw, nb = min(sw, rowbytes - b0);
    if (it.lut)
        for (int i = threadIdx.x; i < 768; i += 256) lc[i] = p.luts[(int64_t)img * 768 + i];
    uint8_t* a = sm;
    uint8_t* b = sm + (size_t)p.h * sw;
    const uint8_t* src = p.tmp + (int64_t)blockIdx.y * p.stride + b0;
    const int total = p.h * sw;
    for (int i = threadIdx.x; i < total; i += 256) {
        const int y = i / sw, xb = i - y * sw;
        if (xb < nb) a[i] = src[(int64_t)y * rowbytes + xb];
    }
    __syncthreads();
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = threadIdx.x; i < total; i += 256) {
            const int y = i / sw, xb = i - y * sw;
            if (xb < nb) b[i] = box_at(a + xb, y, p.h, sw, it.box_radius, it.ww, it.fw);
        }
        __syncthreads();
        uint8_t* t = a; a = b; b = t;
    }
    uint8_t* dst = p.dst + (int64_t)img * p.stride + b0;
    for (int i = threadIdx.x; i < total; i += 256) {
        const int y = i / sw, xb = i - y * sw;
        if (xb < nb) {
            uint8_t v = a[i];
            if (it.lut) v = lc[256 * ((b0 + xb) % 3) + v];
            dst[(int64_t)y * rowbytes + xb] = v;
        }
    }
}

}  // namespace ipp

extern "C" {

int64_t ipp_enhance_workspace_bytes(int w, int h, int n, int n_blur) {
    if (w <= 0 || h <= 0 || n <= 0 || n_blur < 0) return 0;
    return 256 + (((int64_t)8 * n + 255) & ~(int64_t)255) + (((int64_t)4 * n + 255) & ~(int64_t)255) + (int64_t)n_blur * h * w * 3;
}

int ipp_enhance_batch_u8(const uint8_t* src, uint8_t* dst, int w, int h, int n, const ipp_enhance_item* items,
                         const uint8_t* luts, const int32_t* blur_list, int n_blur,
                         void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    IPP_REQUIRE(src && dst && items && luts && workspace && n >= 1 && n_blur >= 0 && n_blur <= n && (blur_list || n_blur == 0), IPP_ERR_BAD_ARG,
                "ipp_enhance_batch_u8: bad arguments");
    IPP_REQUIRE(w > 0 && h > 0, IPP_ERR_BAD_SHAPE, "ipp_enhance_batch_u8: bad geometry");
    IPP_REQUIRE(workspace_bytes >= ipp_enhance_workspace_bytes(w, h, n, n_blur) && (((uintptr_t)workspace) & 255) == 0, IPP_ERR_WORKSPACE,
                "ipp_enhance_batch_u8: workspace too small or not 256-byte aligned (need %lld bytes)",
                (long long)ipp_enhance_workspace_bytes(w, h, n, n_blur));
    EnhParams p{};
    p.src = src; p.dst = dst; p.stride = (int64_t)h * w * 3; p.w = w; p.h = h; p.items = items; p.luts = luts; p.blur_list = blur_list;
    uint8_t* ws = (uint8_t*)workspace;
    p.sums = (unsigned long long*)ws;
    ws += ((int64_t)8 * n + 255) & ~(int64_t)255;
    ws += ((int64_t)4 * n + 255) & ~(int64_t)255;
    p.tmp = ws;
    IPP_CUDA(cudaMemsetAsync(p.sums, 0, (size_t)8 * n, st));
    dim3 grid(kEnhCtas, n);
    { ipp::LaunchTimer _lt("enh_luma_sum_kernel", st); enh_luma_sum_kernel<<<grid, 256, 0, st>>>(p); }
    IPP_LAUNCH_CHECK();
    { ipp::LaunchTimer _lt("enh_point_kernel", st); enh_point_kernel<<<grid, 256, 0, st>>>(p); }
    IPP_LAUNCH_CHECK();
    if (n_blur > 0) {
        const int rowbytes = w * 3;
        int rows = 24576 / rowbytes;
        if (rows < 1) rows = 1;
        if (rows > h) rows = h;
        const size_t smem_h = 2 * (size_t)rows * rowbytes;
        IPP_REQUIRE(smem_h <= 200 * 1024, IPP_ERR_UNSUPPORTED, "ipp_enhance_batch_u8: row too long for the blur");
        int sw = 64;
        while (sw > 4 && 2 * (size_t)h * sw > 200 * 1024) sw >>= 1;
        const size_t smem_v = 2 * (size_t)h * sw;
        IPP_REQUIRE(smem_v <= 200 * 1024, IPP_ERR_UNSUPPORTED, "ipp_enhance_batch_u8: image too tall for the blur");
        static size_t set_h[64] = {}, set_v[64] = {};
        int dev_id = 0;
        IPP_CUDA(cudaGetDevice(&dev_id));
        if (dev_id < 0 || dev_id >= 64 || smem_h > set_h[dev_id]) {
            IPP_CUDA(cudaFuncSetAttribute(enh_blur_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_h, 48 * 1024)));
            if (dev_id >= 0 && dev_id < 64) set_h[dev_id] = smem_h;
        }
        if (dev_id < 0 || dev_id >= 64 || smem_v > set_v[dev_id]) {
            IPP_CUDA(cudaFuncSetAttribute(enh_blur_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_v, 48 * 1024)));
            if (dev_id >= 0 && dev_id < 64) set_v[dev_id] = smem_v;
        }
        { ipp::LaunchTimer _lt("enh_blur_h_kernel", st); enh_blur_h_kernel<<<dim3(ceil_div(h, rows), n_blur), 256, smem_h, st>>>(p, rows); }
        IPP_LAUNCH_CHECK();
        { ipp::LaunchTimer _lt("enh_blur_v_kernel", st); enh_blur_v_kernel<<<dim3(ceil_div(rowbytes, sw), n_blur), 256, smem_v, st>>>(p, sw); }
        IPP_LAUNCH_CHECK();
    }
    return IPP_OK;
}

}  // extern "C"
