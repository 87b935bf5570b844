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
//   3. enh_blur_h / enh_blur_vwalk   only for the images drawn for the blur: the three horizontal passes of a group of 8 rows run
//                            from shared memory; the three vertical passes are chained sliding windows in registers (a thread
//                            walks down one 4-byte column word): one read + one write per axis instead of three.  The channel
//                            tables are applied in the last store.  (enh_blur_v: generic radius / odd row lengths, shared memory.)
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
    const int64_t npx = (int64_t)p.w * p.h;
    const uint8_t* src = p.src + (int64_t)img * p.stride;
    // whole 4-pixel groups (12 aligned bytes) without per-pixel bounds tests, 32-bit indices and a 32-bit partial sum (a thread
    // sees a few dozen pixels); the ragged tail / an unaligned image goes through the checked loader
    const uint32_t nfull = ((((uintptr_t)src) & 3) == 0 && npx < (1ll << 31)) ? (uint32_t)(npx >> 2) : 0u;
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
    uint32_t part = 0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nfull; t += gridDim.x * blockDim.x) {
        const uint32_t a = ldg_stream4(src32 + 3 * t), b = ldg_stream4(src32 + 3 * t + 1), d = ldg_stream4(src32 + 3 * t + 2);
        // a = r0 g0 b0 r1 | b = g1 b1 r2 g2 | d = b2 r3 g3 b3
        part += luma(lb[a & 255u], lb[(a >> 8) & 255u], lb[(a >> 16) & 255u]);
        part += luma(lb[a >> 24], lb[b & 255u], lb[(b >> 8) & 255u]);
        part += luma(lb[(b >> 16) & 255u], lb[b >> 24], lb[d & 255u]);
        part += luma(lb[(d >> 8) & 255u], lb[(d >> 16) & 255u], lb[d >> 24]);
    }
    unsigned long long sum = part;
    const int64_t ngrp = (npx + 3) >> 2;
    for (int64_t t = (int64_t)nfull + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ngrp; t += (int64_t)gridDim.x * blockDim.x) {
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
    auto pixel = [&](uint32_t c0, uint32_t c1, uint32_t c2, uint32_t& o0, uint32_t& o1, uint32_t& o2) {
        const uint32_t r = l1[c0], g = l1[c1], b = l1[c2];
        const float L = (float)luma(r, g, b);
        o0 = blend_byte(L, (float)r, it.color, col_in);
        o1 = blend_byte(L, (float)g, it.color, col_in);
        o2 = blend_byte(L, (float)b, it.color, col_in);
        if (lut_here) { o0 = lc[o0]; o1 = lc[256 + o1]; o2 = lc[512 + o2]; }
    };
    // whole 4-pixel groups: three aligned words in, three out, 32-bit indices, no per-pixel bounds tests
    const uint32_t nfull = ((((uintptr_t)src | (uintptr_t)dst) & 3) == 0 && npx < (1ll << 31)) ? (uint32_t)(npx >> 2) : 0u;
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
    uint32_t* dst32 = reinterpret_cast<uint32_t*>(dst);
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nfull; t += gridDim.x * blockDim.x) {
        const uint32_t a = ldg_stream4(src32 + 3 * t), b = ldg_stream4(src32 + 3 * t + 1), d = ldg_stream4(src32 + 3 * t + 2);
        uint32_t o[12];
        pixel(a & 255u, (a >> 8) & 255u, (a >> 16) & 255u, o[0], o[1], o[2]);
        pixel(a >> 24, b & 255u, (b >> 8) & 255u, o[3], o[4], o[5]);
        pixel((b >> 16) & 255u, b >> 24, d & 255u, o[6], o[7], o[8]);
        pixel((d >> 8) & 255u, (d >> 16) & 255u, d >> 24, o[9], o[10], o[11]);
        dst32[3 * t] = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        dst32[3 * t + 1] = o[4] | (o[5] << 8) | (o[6] << 16) | (o[7] << 24);
        dst32[3 * t + 2] = o[8] | (o[9] << 8) | (o[10] << 16) | (o[11] << 24);
    }
    for (int64_t t = (int64_t)nfull + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ngrp; t += (int64_t)gridDim.x * blockDim.x) {
        uint8_t c[12];
        load_px4(src, t, npx, c);
        uint8_t o[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t r2, g2, b2;
            pixel(c[3 * k], c[3 * k + 1], c[3 * k + 2], r2, g2, b2);
            o[3 * k] = (uint8_t)r2; o[3 * k + 1] = (uint8_t)g2; o[3 * k + 2] = (uint8_t)b2;
        }
        uint8_t* d = dst + 12 * t;
        for (int k = 0; k < 12 && 4 * t + k / 3 < npx; ++k) d[k] = o[k];
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

// three horizontal passes of 8 image rows per CTA from shared memory: blockDim = (128, 8), thread (tx, ty) walks the pixels
// tx, tx + 128, ... of row ty.  Rows are staged as one 32-bit word per pixel (R | G << 8 | B << 16): a neighbour is ONE shared
// load and the three channels are summed at once in two 16-bit x 2 accumulators (the kernel was bound by byte loads before).
constexpr int kBlurRows = 8;
__global__ void __launch_bounds__(1024) enh_blur_h_kernel(const EnhParams p, int walker) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int img = p.blur_list[blockIdx.y];
    if (walker && p.items[img].box_radius <= 3) return;               // handled by the register walker (enh_blur_hwalk_kernel)
    const ipp_enhance_item it = p.items[img];
    const int rowbytes = p.w * 3;
    const int y0 = blockIdx.x * kBlurRows, nr = min(kBlurRows, p.h - y0);
    const int tid = threadIdx.y * 128 + threadIdx.x;
    uint32_t* a = reinterpret_cast<uint32_t*>(sm);
    uint32_t* b = a + (size_t)kBlurRows * p.w;
    const uint8_t* src = p.dst + (int64_t)img * p.stride + (int64_t)y0 * rowbytes;      // the point pass wrote dst
    const int npx = nr * p.w;
    for (int i = tid; i < npx; i += 1024) {
        const uint8_t* q = src + 3 * i;
        a[i] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16);
    }
    __syncthreads();
    const int r = it.box_radius, last = p.w - 1;
    const uint32_t ww = it.ww, fw = it.fw;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        if ((int)threadIdx.y < nr) {
            const uint32_t* row = a + threadIdx.y * p.w;
            uint32_t* out = b + threadIdx.y * p.w;
            for (int x = threadIdx.x; x <= last; x += 128) {
                uint32_t lo = 0, hi = 0;
                for (int d = -r; d <= r; ++d) {
                    const uint32_t v = row[min(max(x + d, 0), last)];
                    lo += v & 0x00ff00ffu; hi += (v >> 8) & 0xffu;
                }
                const uint32_t fa = row[min(max(x - r - 1, 0), last)], fb = row[min(max(x + r + 1, 0), last)];
                const uint32_t flo = (fa & 0x00ff00ffu) + (fb & 0x00ff00ffu), fhi = ((fa >> 8) & 0xffu) + ((fb >> 8) & 0xffu);
                const uint32_t c0 = ((lo & 0xffffu) * ww + (flo & 0xffffu) * fw + (1u << 23)) >> 24;
                const uint32_t c2 = ((lo >> 16) * ww + (flo >> 16) * fw + (1u << 23)) >> 24;
                const uint32_t c1 = (hi * ww + fhi * fw + (1u << 23)) >> 24;
                out[x] = (c0 & 255u) | ((c1 & 255u) << 8) | ((c2 & 255u) << 16);
            }
        }
        __syncthreads();
        uint32_t* t = a; a = b; b = t;
    }
    uint8_t* dst = p.tmp + (int64_t)blockIdx.y * p.stride + (int64_t)y0 * rowbytes;
    const int total = nr * rowbytes;
    for (int i = tid; i < total; i += 1024) {
        const int px = i / 3;                      // constant divisor: a multiply
        dst[i] = (uint8_t)(a[px] >> (8 * (i - 3 * px)));
    }
}

// three vertical passes, generic radius: a strip of `sw` bytes x all rows per CTA; channel tables in the last store
__global__ void __launch_bounds__(256) enh_blur_v_kernel(const EnhParams p, int sw, int walker) {
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ uint8_t lc[768];
    const int img = p.blur_list[blockIdx.y];
    const ipp_enhance_item it = p.items[img];
    if (walker && it.box_radius <= 3) return;                        // handled by the register walker below
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

// three vertical passes, radius <= 3 (every GaussianBlur radius tranfo.py draws): a thread owns one 4-byte column word and
// walks down the image through three chained sliding windows held in registers — every byte is read once and written once,
// a warp reads 128 contiguous bytes per row.  Four bytes are summed at once in two 16-bit x 2 accumulators.
struct Walk {       // one box pass over a column word: ring of the last 2r + 3 inputs (rows y-r-1 .. y+r+1 around the output row)
    uint32_t ring[9];
    uint32_t lo, hi;    // window sums of bytes {0, 2} and {1, 3}
};
__device__ __forceinline__ uint32_t walk_emit(const Walk& w, int r, uint32_t ww, uint32_t fw) {
    // ring[0] = row y-r-1 (far above), ring[2r+2] = row y+r+1 (far below); lo / hi = sums over ring[1 .. 2r+1]
    const uint32_t fa = w.ring[0], fb = w.ring[2 * r + 2];
    const uint32_t flo = (fa & 0x00ff00ffu) + (fb & 0x00ff00ffu), fhi = ((fa >> 8) & 0x00ff00ffu) + ((fb >> 8) & 0x00ff00ffu);
    const uint32_t b0 = ((w.lo & 0xffffu) * ww + (flo & 0xffffu) * fw + (1u << 23)) >> 24;
    const uint32_t b2 = ((w.lo >> 16) * ww + (flo >> 16) * fw + (1u << 23)) >> 24;
    const uint32_t b1 = ((w.hi & 0xffffu) * ww + (fhi & 0xffffu) * fw + (1u << 23)) >> 24;
    const uint32_t b3 = ((w.hi >> 16) * ww + (fhi >> 16) * fw + (1u << 23)) >> 24;
    return (b0 & 255u) | ((b1 & 255u) << 8) | ((b2 & 255u) << 16) | (b3 << 24);
}
template <int R>
__device__ __forceinline__ void walk_push(Walk& w, uint32_t v) {
    // the window moves down one row: ring[1] leaves the sum, the old far-below element enters it, v becomes far-below
    const uint32_t out = w.ring[1], in = w.ring[2 * R + 2];
    w.lo += (in & 0x00ff00ffu) - (out & 0x00ff00ffu);
    w.hi += ((in >> 8) & 0x00ff00ffu) - ((out >> 8) & 0x00ff00ffu);
#pragma unroll
    for (int i = 0; i < 2 * R + 2; ++i) w.ring[i] = w.ring[i + 1];
    w.ring[2 * R + 2] = v;
}
template <int R>
__device__ __forceinline__ void walk_init(Walk& w, uint32_t first) {
    // state for output row -1 ... every ring entry = row 0 (clamped), sums over 2R+1 copies
#pragma unroll
    for (int i = 0; i < 2 * R + 3; ++i) w.ring[i] = first;
    w.lo = (first & 0x00ff00ffu) * (2 * R + 1);
    w.hi = ((first >> 8) & 0x00ff00ffu) * (2 * R + 1);
}
__device__ __forceinline__ uint32_t lut_word(const uint8_t* lc, uint32_t v, int ch0) {
    // channel of byte k of this word = (ch0 + k) % 3
    const int c1 = ch0 == 2 ? 0 : ch0 + 1, c2 = c1 == 2 ? 0 : c1 + 1;
    return (uint32_t)lc[256 * ch0 + (v & 255u)] | ((uint32_t)lc[256 * c1 + ((v >> 8) & 255u)] << 8) |
           ((uint32_t)lc[256 * c2 + ((v >> 16) & 255u)] << 16) | ((uint32_t)lc[256 * ch0 + (v >> 24)] << 24);
}
// Rows [y0, y1) are produced by a walk that starts 3 (R + 1) rows above y0 (as if that row were the top: the outputs that are
// wrong for it are exactly the ones not stored — pass k is correct from k (R + 1) rows below the start) and clamps at the true
// bottom.  The caller passes src / dst already offset to the start row s; h = rows from s to the bottom, first = y0 - s.
template <int R>
__device__ __forceinline__ void walk_column(const uint32_t* src, uint32_t* dst, int pitch_w, int h, int first, int count, uint32_t ww, uint32_t fw,
                                            const uint8_t* lc, bool lut, int ch0) {
    // Pass k sees the output rows of pass k-1 as its input rows.  A pass can emit output row y once its input row
    // y + R + 1 (clamped to h - 1) has been pushed; feeding the last row again beyond the end reproduces the clamped edge.
    Walk w1, w2, w3;
    const uint32_t f0 = src[0];
    walk_init<R>(w1, f0);
    // prime pass 1 with input rows 0 .. R (ring then holds rows -R-1 .. R+1 around output row 0 after R + 2 pushes in total)
    int in1 = 0;                       // next input row of pass 1
    auto next1 = [&]() { const uint32_t v = src[(int64_t)min(in1, h - 1) * pitch_w]; ++in1; return v; };
    // after walk_init the ring models output row -1 - ... we push R + 2 rows (rows 0 .. R+1) to reach output row 0:
    // ring = {row -R-1 .. row R+1} = {0 x (R+2) clamped, 1 .. R+1}
    // walk_init put row 0 everywhere = state "output row -(R+2)+... "; pushing rows 1 .. R+1 yields output row 0.
    next1();                           // row 0 is already everywhere
    for (int i = 0; i < R + 1; ++i) walk_push<R>(w1, next1());
    int out1 = 0;                      // next output row of pass 1
    auto emit1 = [&]() { const uint32_t v = walk_emit(w1, R, ww, fw); walk_push<R>(w1, next1()); ++out1; return v; };
    // pass 2 primed from pass 1's outputs 0 .. R+1
    const uint32_t g0 = emit1();
    walk_init<R>(w2, g0);
    for (int i = 0; i < R + 1; ++i) walk_push<R>(w2, out1 < h ? emit1() : w2.ring[2 * R + 2]);
    int out2 = 0;
    auto next2 = [&]() { return out1 < h ? emit1() : w2.ring[2 * R + 2]; };        // beyond the end: the last row again
    auto emit2 = [&]() { const uint32_t v = walk_emit(w2, R, ww, fw); walk_push<R>(w2, next2()); ++out2; return v; };
    const uint32_t h0 = emit2();
    walk_init<R>(w3, h0);
    for (int i = 0; i < R + 1; ++i) walk_push<R>(w3, out2 < h ? emit2() : w3.ring[2 * R + 2]);
    auto next3 = [&]() { return out2 < h ? emit2() : w3.ring[2 * R + 2]; };
    for (int y = 0; y < first + count; ++y) {
        uint32_t v = walk_emit(w3, R, ww, fw);
        walk_push<R>(w3, next3());
        if (y >= first) {
            if (lut) v = lut_word(lc, v, ch0);
            dst[(int64_t)y * pitch_w] = v;
        }
    }
}
__global__ void __launch_bounds__(128) enh_blur_vwalk_kernel(const EnhParams p) {
    __shared__ uint8_t lc[768];
    const int img = p.blur_list[blockIdx.y];
    const ipp_enhance_item it = p.items[img];
    const int pitch_w = (p.w * 3) >> 2;
    if (it.box_radius > 3) return;                                  // generic kernel
    if (it.lut)
        for (int i = threadIdx.x; i < 768; i += 128) lc[i] = p.luts[(int64_t)img * 768 + i];
    __syncthreads();
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col >= pitch_w) return;
    // the column is cut into gridDim.z segments (more threads in flight: a walk is a chain of dependent loads)
    const int seg_h = (p.h + (int)gridDim.z - 1) / (int)gridDim.z;
    const int y0 = (int)blockIdx.z * seg_h, y1 = min(y0 + seg_h, p.h);
    if (y0 >= y1) return;
    const int s0 = max(y0 - 3 * (it.box_radius + 1), 0);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.tmp + (int64_t)blockIdx.y * p.stride) + col + (int64_t)s0 * pitch_w;
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.dst + (int64_t)img * p.stride) + col + (int64_t)s0 * pitch_w;
    const int ch0 = (4 * col) % 3;
    const int hh = p.h - s0, first = y0 - s0, count = y1 - y0;
    switch (it.box_radius) {
        case 0: walk_column<0>(src, dst, pitch_w, hh, first, count, it.ww, it.fw, lc, it.lut != 0, ch0); break;
        case 1: walk_column<1>(src, dst, pitch_w, hh, first, count, it.ww, it.fw, lc, it.lut != 0, ch0); break;
        case 2: walk_column<2>(src, dst, pitch_w, hh, first, count, it.ww, it.fw, lc, it.lut != 0, ch0); break;
        default: walk_column<3>(src, dst, pitch_w, hh, first, count, it.ww, it.fw, lc, it.lut != 0, ch0); break;
    }
}

// three horizontal passes, radius <= 3: the vertical walker turned by 90 degrees.  Eight image rows are staged as one 32-bit word
// per pixel (R | G << 8 | B << 16: the byte lanes of a word are the channels, exactly what walk_column sums in parallel); one warp
// per row, lane = a segment of the row, which it walks through the three chained sliding windows in registers — one shared load
// and one shared store per pixel and pass chain, where the windowed kernel above re-reads 2r + 3 neighbours per pixel and pass.
// A segment starts 3 (r + 1) pixels early (walk_column's warm-up); its length is odd so that the 32 lanes of a warp sit in
// different banks.
constexpr int kHwRows = 8;
__global__ void __launch_bounds__(32 * kHwRows) enh_blur_hwalk_kernel(const EnhParams p) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int img = p.blur_list[blockIdx.y];
    const ipp_enhance_item it = p.items[img];
    if (it.box_radius > 3) return;                                    // generic kernel
    const int rowbytes = p.w * 3;
    const int y0 = blockIdx.x * kHwRows, nr = min(kHwRows, p.h - y0);
    const int tid = threadIdx.x;
    uint32_t* a = reinterpret_cast<uint32_t*>(sm);
    uint32_t* b = a + (size_t)kHwRows * p.w;
    const uint8_t* src = p.dst + (int64_t)img * p.stride + (int64_t)y0 * rowbytes;      // the point pass wrote dst
    const int npx = nr * p.w;
    for (int i = tid; i < npx; i += 32 * kHwRows) {
        const uint8_t* q = src + 3 * i;
        a[i] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16);
    }
    __syncthreads();
    {
        const int row = tid >> 5, seg = tid & 31;
        const int S = ((p.w + 31) >> 5) | 1;
        const int x0 = seg * S, x1 = min(x0 + S, p.w);
        if (row < nr && x0 < x1) {
            const int s0 = max(x0 - 3 * (it.box_radius + 1), 0);
            const uint32_t* ws = a + (size_t)row * p.w + s0;
            uint32_t* wd = b + (size_t)row * p.w + s0;
            const int hh = p.w - s0, first = x0 - s0, count = x1 - x0;
            switch (it.box_radius) {
                case 0: walk_column<0>(ws, wd, 1, hh, first, count, it.ww, it.fw, nullptr, false, 0); break;
                case 1: walk_column<1>(ws, wd, 1, hh, first, count, it.ww, it.fw, nullptr, false, 0); break;
                case 2: walk_column<2>(ws, wd, 1, hh, first, count, it.ww, it.fw, nullptr, false, 0); break;
                default: walk_column<3>(ws, wd, 1, hh, first, count, it.ww, it.fw, nullptr, false, 0); break;
            }
        }
    }
    __syncthreads();
    uint8_t* dst = p.tmp + (int64_t)blockIdx.y * p.stride + (int64_t)y0 * rowbytes;
    const int total = nr * rowbytes;
    if ((((uintptr_t)dst) & 3) == 0 && (total & 3) == 0) {
        // word j of the packed rows = bytes 4j .. 4j+3 = channels (4j + t) % 3 of pixels (4j + t) / 3
        for (int j = tid; j < (total >> 2); j += 32 * kHwRows) {
            const int px = (4 * j) / 3, ph = 4 * j - 3 * px;          // ph = 0, 1, 2
            const uint32_t v0 = b[px], v1 = b[min(px + 1, npx - 1)];
            // 6 candidate bytes: v0 channels 0..2 then v1 channels 0..2; the word takes 4 of them starting at ph
            const uint64_t six = (uint64_t)(v0 & 0x00ffffffu) | ((uint64_t)(v1 & 0x00ffffffu) << 24);
            reinterpret_cast<uint32_t*>(dst)[j] = (uint32_t)(six >> (8 * ph));
        }
    } else {
        for (int i = tid; i < total; i += 32 * kHwRows) {
            const int px = i / 3;                      // constant divisor: a multiply
            dst[i] = (uint8_t)(b[px] >> (8 * (i - 3 * px)));
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
        const size_t smem_h = 2 * (size_t)kBlurRows * w * 4;
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
        // horizontal passes: the register walker takes the images with a box radius <= 3, the windowed kernel the others
        static const bool h_walker = getenv("IPP_ENH_NO_HWALK") == nullptr;
        if (h_walker) {
            const size_t smem_w = 2 * (size_t)kHwRows * w * 4;
            static size_t set_w[64] = {};
            if (dev_id < 0 || dev_id >= 64 || smem_w > set_w[dev_id]) {
                IPP_CUDA(cudaFuncSetAttribute(enh_blur_hwalk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_w, 48 * 1024)));
                if (dev_id >= 0 && dev_id < 64) set_w[dev_id] = smem_w;
            }
            { ipp::LaunchTimer _lt("enh_blur_hwalk_kernel", st); enh_blur_hwalk_kernel<<<dim3(ceil_div(h, kHwRows), n_blur), 32 * kHwRows, smem_w, st>>>(p); }
            IPP_LAUNCH_CHECK();
        }
        { ipp::LaunchTimer _lt("enh_blur_h_kernel", st); enh_blur_h_kernel<<<dim3(ceil_div(h, kBlurRows), n_blur), dim3(128, kBlurRows), smem_h, st>>>(p, h_walker ? 1 : 0); }
        IPP_LAUNCH_CHECK();
        // vertical passes: the register walker takes the images with a box radius <= 3 (and 4-byte-multiple rows), the generic
        // shared-memory kernel the others; each kernel returns at once for the images of the other
        const int walker = rowbytes % 4 == 0 && (((uintptr_t)dst | (uintptr_t)p.tmp) & 3) == 0 && (p.stride & 3) == 0;
        if (walker) {
            { ipp::LaunchTimer _lt("enh_blur_vwalk_kernel", st); enh_blur_vwalk_kernel<<<dim3(ceil_div(rowbytes / 4, 128), n_blur, h >= 256 ? 8 : 1), 128, 0, st>>>(p); }
            IPP_LAUNCH_CHECK();
        }
        { ipp::LaunchTimer _lt("enh_blur_v_kernel", st); enh_blur_v_kernel<<<dim3(ceil_div(rowbytes, sw), n_blur), 256, smem_v, st>>>(p, sw, walker); }
        IPP_LAUNCH_CHECK();
    }
    return IPP_OK;
}

}  // extern "C"
