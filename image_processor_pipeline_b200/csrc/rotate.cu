// rotate.cu — inverse-mapped rotation warp with fused alpha-bbox reduction, and
// the stand-alone non-zero bounding box.
//
// Reference behaviour replaced: PIL Image.rotate(angle, expand=True) — libImaging
// affine_fixed (NEAREST, 16.16 fixed point; transforms/rotations.py:96) — and
// PIL getbbox() (rotations.py:99, recadrages.py:73) / cv2.findNonZero+boundingRect
// (pixels_isolés.py:77-79).
//
// Output tiles are 32x32 pixels per CTA (8 lanes x 4 px wide, 4 rows per warp): the
// rotated source footprint of a square tile is compact, so the gathers of one CTA
// share L1 lines, every source byte is fetched from HBM once (L2 holds the image)
// and each lane stores one aligned 16-byte vector.
#include "common.cuh"

namespace ipp {

struct RotParams {
    const uint8_t* src;
    int64_t src_pitch;
    int src_w, src_h;
    uint8_t* dst;
    int64_t dst_pitch;
    int dst_w, dst_h;
    int a0, a1, a2, a3, a4, a5;
    double m[6];
    int32_t* bbox;
};

__device__ __forceinline__ void bbox_block_reduce(int32_t* bbox, int minx, int miny, int maxx1, int maxy1) {
    // block-wide reduction of (min x, min y, max x+1, max y+1) -> 4 atomics per CTA at most
    __shared__ int s[4];
    if (threadIdx.x == 0) {
        s[0] = INT_MAX; s[1] = INT_MAX; s[2] = 0; s[3] = 0;
    }
    __syncthreads();
    minx = warp_min(minx);
    miny = warp_min(miny);
    maxx1 = warp_max(maxx1);
    maxy1 = warp_max(maxy1);
    if ((threadIdx.x & 31) == 0 && maxx1 > 0) {
        atomicMin(&s[0], minx);
        atomicMin(&s[1], miny);
        atomicMax(&s[2], maxx1);
        atomicMax(&s[3], maxy1);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s[2] > 0) {
        atomicMin(&bbox[0], s[0]);
        atomicMin(&bbox[1], s[1]);
        atomicMax(&bbox[2], s[2]);
        atomicMax(&bbox[3], s[3]);
    }
}

__device__ __forceinline__ uint32_t load_px4(const uint8_t* p) {
    if ((((uintptr_t)p) & 3) == 0) return __ldg(reinterpret_cast<const uint32_t*>(p));
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// mode 0: nearest, fixed point
__global__ void __launch_bounds__(256) rotate_nearest_kernel(const RotParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x = blockIdx.x * 32 + (lane & 7) * 4;
    const int y = blockIdx.y * 32 + warp * 4 + (lane >> 3);
    int minx = INT_MAX, miny = INT_MAX, maxx1 = 0, maxy1 = 0;
    if (y < p.dst_h && x < p.dst_w) {
        uint32_t px[4];
        int xx = p.a2 + y * p.a1 + x * p.a0;
        int yy = p.a5 + y * p.a4 + x * p.a3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xin = xx >> 16, yin = yy >> 16;
            uint32_t v = 0;
            if (xin >= 0 && xin < p.src_w && yin >= 0 && yin < p.src_h)
                v = load_px4(p.src + (int64_t)yin * p.src_pitch + (int64_t)xin * 4);
            px[k] = v;
            xx += p.a0;
            yy += p.a3;
        }
        uint8_t* d = p.dst + (int64_t)y * p.dst_pitch + (int64_t)x * 4;
        if (x + 3 < p.dst_w && (((uintptr_t)d) & 15) == 0) {
            *reinterpret_cast<uint4*>(d) = make_uint4(px[0], px[1], px[2], px[3]);
        } else {
            for (int k = 0; k < 4 && x + k < p.dst_w; ++k) {
                uint8_t* dk = d + 4 * k;
                dk[0] = (uint8_t)px[k]; dk[1] = (uint8_t)(px[k] >> 8);
                dk[2] = (uint8_t)(px[k] >> 16); dk[3] = (uint8_t)(px[k] >> 24);
            }
        }
        if (p.bbox) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (x + k < p.dst_w && (px[k] >> 24) != 0) {
                    minx = min(minx, x + k);
                    maxx1 = max(maxx1, x + k + 1);
                    miny = y;
                    maxy1 = y + 1;
                }
            }
        }
    }
    if (p.bbox) bbox_block_reduce(p.bbox, minx, miny, maxx1, maxy1);
}

// premultiplied sample for the bilinear mode (libImaging rgbA2rgba: MULDIV255)
__device__ __forceinline__ void premult_px(uint32_t v, double out[4]) {
    const uint32_t a = v >> 24;
    out[0] = (double)div255((v & 255u) * a);
    out[1] = (double)div255(((v >> 8) & 255u) * a);
    out[2] = (double)div255(((v >> 16) & 255u) * a);
    out[3] = (double)a;
}

// mode 1: bilinear on premultiplied alpha — libImaging ImagingGenericTransform +
// affine_transform + bilinear_filter32RGB in double precision, then rgba2rgbA.
__global__ void __launch_bounds__(256) rotate_bilinear_kernel(const RotParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32 + (lane & 7) * 4;
    const int y = blockIdx.y * 32 + warp * 4 + (lane >> 3);
    int minx = INT_MAX, miny = INT_MAX, maxx1 = 0, maxy1 = 0;
    if (y < p.dst_h) {
        for (int k = 0; k < 4; ++k) {
            const int x = x0 + k;
            if (x >= p.dst_w) break;
            // xin = a0*(x+.5) + a1*(y+.5) + a2, evaluated as C evaluates it (no fma contraction)
            const double xo = (double)x + 0.5, yo = (double)y + 0.5;
            double xin = __dadd_rn(__dadd_rn(__dmul_rn(p.m[0], xo), __dmul_rn(p.m[1], yo)), p.m[2]);
            double yin = __dadd_rn(__dadd_rn(__dmul_rn(p.m[3], xo), __dmul_rn(p.m[4], yo)), p.m[5]);
            uint32_t outpx = 0;
            if (xin >= 0.0 && xin < (double)p.src_w && yin >= 0.0 && yin < (double)p.src_h) {
                xin -= 0.5;
                yin -= 0.5;
                const int xi = (int)floor(xin), yi = (int)floor(yin);
                const double dx = xin - (double)xi, dy = yin - (double)yi;
                const int xa = min(max(xi, 0), p.src_w - 1), xb = min(max(xi + 1, 0), p.src_w - 1);
                const int ya = min(max(yi, 0), p.src_h - 1), yb = min(max(yi + 1, 0), p.src_h - 1);
                double s00[4], s01[4], s10[4], s11[4];
                premult_px(load_px4(p.src + (int64_t)ya * p.src_pitch + (int64_t)xa * 4), s00);
                premult_px(load_px4(p.src + (int64_t)ya * p.src_pitch + (int64_t)xb * 4), s01);
                premult_px(load_px4(p.src + (int64_t)yb * p.src_pitch + (int64_t)xa * 4), s10);
                premult_px(load_px4(p.src + (int64_t)yb * p.src_pitch + (int64_t)xb * 4), s11);
                uint32_t c[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const double v1 = __dadd_rn(s00[b], __dmul_rn(__dsub_rn(s01[b], s00[b]), dx));
                    const double v2 = __dadd_rn(s10[b], __dmul_rn(__dsub_rn(s11[b], s10[b]), dx));
                    const double v = __dadd_rn(v1, __dmul_rn(__dsub_rn(v2, v1), dy));
                    c[b] = (uint32_t)(int)v;   // (UINT8) truncation of a value in [0,255]
                }
                const uint32_t a = c[3];
                if (a != 0 && a != 255) {       // libImaging rgba2rgbA: CLIP8(255*c/a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) c[b] = min(255u, (255u * c[b]) / a);
                }
                outpx = (c[0] & 255u) | ((c[1] & 255u) << 8) | ((c[2] & 255u) << 16) | (a << 24);
            }
            uint8_t* d = p.dst + (int64_t)y * p.dst_pitch + (int64_t)x * 4;
            if ((((uintptr_t)d) & 3) == 0) {
                *reinterpret_cast<uint32_t*>(d) = outpx;
            } else {
                d[0] = (uint8_t)outpx; d[1] = (uint8_t)(outpx >> 8);
                d[2] = (uint8_t)(outpx >> 16); d[3] = (uint8_t)(outpx >> 24);
            }
            if (p.bbox && (outpx >> 24) != 0) {
                minx = min(minx, x);
                maxx1 = max(maxx1, x + 1);
                miny = y;
                maxy1 = y + 1;
            }
        }
    }
    if (p.bbox) bbox_block_reduce(p.bbox, minx, miny, maxx1, maxy1);
}

// ------------------------------------------------------------------ bbox of non-zero
__global__ void bbox_reset_kernel(int32_t* bbox, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        bbox[4 * i + 0] = INT_MAX;
        bbox[4 * i + 1] = INT_MAX;
        bbox[4 * i + 2] = 0;
        bbox[4 * i + 3] = 0;
    }
}

// one warp per row segment of 32 lanes x 16 bytes; rows strided over blockIdx.y
__global__ void __launch_bounds__(256) alpha_bbox_kernel(const uint8_t* __restrict__ img, int64_t pitch, int w, int h,
                                                         int ch, int alpha_only, int32_t* bbox) {
    const int rowbytes = w * ch;
    int minx = INT_MAX, miny = INT_MAX, maxx1 = 0, maxy1 = 0;
    const int y = blockIdx.y;
    const uint8_t* row = img + (int64_t)y * pitch;
    const bool vec = ((((uintptr_t)row) & 15) == 0);
    for (int b0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16; b0 < rowbytes; b0 += gridDim.x * blockDim.x * 16) {
        uint32_t wds[4] = {0, 0, 0, 0};
        const int nbytes = min(16, rowbytes - b0);
        if (vec && nbytes == 16) {
            const uint4 v = ldg_stream16(row + b0);
            wds[0] = v.x; wds[1] = v.y; wds[2] = v.z; wds[3] = v.w;
        } else {
            for (int k = 0; k < nbytes; ++k) wds[k >> 2] |= (uint32_t)row[b0 + k] << (8 * (k & 3));
        }
        if ((wds[0] | wds[1] | wds[2] | wds[3]) == 0) continue;
        for (int k = 0; k < nbytes; ++k) {
            const uint32_t byte = (wds[k >> 2] >> (8 * (k & 3))) & 255u;
            if (byte == 0) continue;
            const int bi = b0 + k;
            const int px = bi / ch;
            if (alpha_only && (bi - px * ch) != ch - 1) continue;
            minx = min(minx, px);
            maxx1 = max(maxx1, px + 1);
            miny = y;
            maxy1 = y + 1;
        }
    }
    bbox_block_reduce(bbox, minx, miny, maxx1, maxy1);
}

}  // namespace ipp

extern "C" {

int ipp_bbox_reset(int32_t* bbox, int n, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(bbox && n >= 1, IPP_ERR_BAD_ARG, "ipp_bbox_reset: bad arguments");
    { ipp::LaunchTimer _lt("bbox_reset_kernel", (cudaStream_t)stream); bbox_reset_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(bbox, n); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

int ipp_alpha_bbox_u8(const uint8_t* img, int64_t pitch, int w, int h, int channels,
                      int alpha_only, int32_t* bbox_out, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(img && bbox_out, IPP_ERR_BAD_ARG, "ipp_alpha_bbox_u8: NULL pointer");
    IPP_REQUIRE(channels >= 1 && channels <= 4 && w >= 0 && h >= 0 && pitch >= (int64_t)w * channels,
                IPP_ERR_BAD_SHAPE, "ipp_alpha_bbox_u8: bad geometry");
    cudaStream_t st = (cudaStream_t)stream;
    { ipp::LaunchTimer _lt("bbox_reset_kernel", st); bbox_reset_kernel<<<1, 32, 0, st>>>(bbox_out, 1); }
    IPP_LAUNCH_CHECK();
    if (w == 0 || h == 0) return IPP_OK;
    const int rowbytes = w * channels;
    dim3 grid(min(8, ceil_div(rowbytes, 256 * 16)), h);
    { ipp::LaunchTimer _lt("alpha_bbox_kernel", st); alpha_bbox_kernel<<<grid, 256, 0, st>>>(img, pitch, w, h, channels, alpha_only, bbox_out); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

int ipp_rotate_rgba_u8(const uint8_t* src, int64_t src_pitch, int src_w, int src_h,
                       int mode, const int32_t* coef_fixed_host, const double* coef_f64_host,
                       uint8_t* dst, int64_t dst_pitch, int dst_w, int dst_h,
                       int32_t* bbox_out, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(src && dst, IPP_ERR_BAD_ARG, "ipp_rotate_rgba_u8: NULL image pointer");
    IPP_REQUIRE(mode == 0 || mode == 1, IPP_ERR_BAD_ARG, "ipp_rotate_rgba_u8: mode=%d", mode);
    IPP_REQUIRE((mode == 0 && coef_fixed_host) || (mode == 1 && coef_f64_host), IPP_ERR_BAD_ARG,
                "ipp_rotate_rgba_u8: coefficients missing for mode %d", mode);
    IPP_REQUIRE(src_w > 0 && src_h > 0 && dst_w >= 0 && dst_h >= 0 && src_pitch >= (int64_t)src_w * 4 &&
                    dst_pitch >= (int64_t)dst_w * 4,
                IPP_ERR_BAD_SHAPE, "ipp_rotate_rgba_u8: bad geometry");
    if (dst_w == 0 || dst_h == 0) return IPP_OK;
    RotParams p{};
    p.src = src; p.src_pitch = src_pitch; p.src_w = src_w; p.src_h = src_h;
    p.dst = dst; p.dst_pitch = dst_pitch; p.dst_w = dst_w; p.dst_h = dst_h;
    p.bbox = bbox_out;
    dim3 grid(ceil_div(dst_w, 32), ceil_div(dst_h, 32));
    if (mode == 0) {
        p.a0 = coef_fixed_host[0]; p.a1 = coef_fixed_host[1]; p.a2 = coef_fixed_host[2];
        p.a3 = coef_fixed_host[3]; p.a4 = coef_fixed_host[4]; p.a5 = coef_fixed_host[5];
        { ipp::LaunchTimer _lt("rotate_nearest_kernel", (cudaStream_t)stream); rotate_nearest_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p); }
    } else {
        for (int i = 0; i < 6; ++i) p.m[i] = coef_f64_host[i];
        { ipp::LaunchTimer _lt("rotate_bilinear_kernel", (cudaStream_t)stream); rotate_bilinear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p); }
    }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

}  // extern "C"
