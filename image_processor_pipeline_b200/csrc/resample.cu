// resample.cu — Pillow-compatible 8bpc separable resample of RGBA sprites.
//
// Reference behaviour replaced: Image.resize((w, h), LANCZOS) on RGBA
// (transforms/overlays.py:129): RGBA -> RGBa (MULDIV255), libImaging
// ImagingResampleHorizontal_8bpc then ImagingResampleVertical_8bpc
// (clip8((2^21 + sum px*k) >> 22), 8-bit intermediate), RGBa -> RGBA.
// The integer coefficient tables come from the host (exact libImaging
// precompute_coeffs / normalize_coeffs_8bpc arithmetic in double precision).
#include "common.cuh"

namespace ipp {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint32_t clip8(int v) {
    v >>= kPrecisionBits;
    return (uint32_t)min(max(v, 0), 255);
}

__device__ __forceinline__ uint32_t premultiply_px(uint32_t v) {
    const uint32_t a = v >> 24;
    return div255((v & 255u) * a) | (div255(((v >> 8) & 255u) * a) << 8) |
           (div255(((v >> 16) & 255u) * a) << 16) | (a << 24);
}

__device__ __forceinline__ uint32_t unpremultiply_px(uint32_t v) {
    const uint32_t a = v >> 24;
    if (a == 0u || a == 255u) return v;
    const uint32_t c0 = min(255u, (255u * (v & 255u)) / a);
    const uint32_t c1 = min(255u, (255u * ((v >> 8) & 255u)) / a);
    const uint32_t c2 = min(255u, (255u * ((v >> 16) & 255u)) / a);
    return c0 | (c1 << 8) | (c2 << 16) | (a << 24);
}

// horizontal pass: out[y][X] for y in [0, src_h), X in [0, dst_w).
// PREMUL: source is straight-alpha RGBA and is premultiplied on the fly.
// UNPREMUL: this is the last pass, convert back to straight alpha on store.
template <bool PREMUL, bool UNPREMUL>
__global__ void __launch_bounds__(256) resample_h_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, int src_h,
                                                         uint8_t* __restrict__ dst, int64_t dst_pitch, int dst_w,
                                                         const int32_t* __restrict__ bounds,
                                                         const int32_t* __restrict__ kk, int ksize) {
    const int X = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (X >= dst_w || y >= src_h) return;
    const int xmin = bounds[2 * X], n = bounds[2 * X + 1];
    const int32_t* k = kk + (int64_t)X * ksize;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(src + (int64_t)y * src_pitch) + xmin;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0, s3 = s0;
    for (int i = 0; i < n; ++i) {
        uint32_t v = __ldg(row + i);
        if (PREMUL) v = premultiply_px(v);
        const int c = __ldg(k + i);
        s0 += (int)__byte_perm(v, 0u, 0x4440) * c;      // PRMT: one alu op per byte instead of shift + mask
        s1 += (int)__byte_perm(v, 0u, 0x4441) * c;
        s2 += (int)__byte_perm(v, 0u, 0x4442) * c;
        s3 += (int)__byte_perm(v, 0u, 0x4443) * c;
    }
    uint32_t o = clip8(s0) | (clip8(s1) << 8) | (clip8(s2) << 16) | (clip8(s3) << 24);
    if (UNPREMUL) o = unpremultiply_px(o);
    reinterpret_cast<uint32_t*>(dst + (int64_t)y * dst_pitch)[X] = o;
}

// vertical pass: out[Y][x] for Y in [0, dst_h), x in [0, w).
template <bool PREMUL, bool UNPREMUL>
__global__ void __launch_bounds__(256) resample_v_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, int w,
                                                         uint8_t* __restrict__ dst, int64_t dst_pitch, int dst_h,
                                                         const int32_t* __restrict__ bounds,
                                                         const int32_t* __restrict__ kk, int ksize) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || Y >= dst_h) return;
    const int ymin = bounds[2 * Y], n = bounds[2 * Y + 1];
    const int32_t* k = kk + (int64_t)Y * ksize;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0, s3 = s0;
    for (int i = 0; i < n; ++i) {
        uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + (int64_t)(ymin + i) * src_pitch) + x);
        if (PREMUL) v = premultiply_px(v);
        const int c = __ldg(k + i);
        s0 += (int)__byte_perm(v, 0u, 0x4440) * c;      // PRMT: one alu op per byte instead of shift + mask
        s1 += (int)__byte_perm(v, 0u, 0x4441) * c;
        s2 += (int)__byte_perm(v, 0u, 0x4442) * c;
        s3 += (int)__byte_perm(v, 0u, 0x4443) * c;
    }
    uint32_t o = clip8(s0) | (clip8(s1) << 8) | (clip8(s2) << 16) | (clip8(s3) << 24);
    if (UNPREMUL) o = unpremultiply_px(o);
    reinterpret_cast<uint32_t*>(dst + (int64_t)Y * dst_pitch)[x] = o;
}

}  // namespace ipp

// ------------------------------------------------------------------ host coefficient tables
// libImaging Resample.c precompute_coeffs() + normalize_coeffs_8bpc() for the LANCZOS
// filter (support 3) over the full-image box, in the same double-precision operation
// order and with the C library's sin(), so the tables are the ones Pillow computes.
#include <cmath>
namespace ipp {
static inline double sinc_filter(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return std::sin(x) / x;
}
static inline double lanczos_filter(double x) {
    if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
    return 0.0;
}
static inline void lanczos_scale(int in_size, int out_size, double* scale, double* filterscale, double* support, int* ksize) {
    *filterscale = *scale = (double)((float)in_size - 0.0f) / out_size;
    if (*filterscale < 1.0) *filterscale = 1.0;
    *support = 3.0 * *filterscale;
    *ksize = (int)std::ceil(*support) * 2 + 1;
}
}  // namespace ipp

extern "C" int ipp_lanczos_ksize(int in_size, int out_size) {
    if (in_size <= 0 || out_size <= 0) return 0;
    double scale, fs, support;
    int ksize;
    ipp::lanczos_scale(in_size, out_size, &scale, &fs, &support, &ksize);
    return ksize;
}

extern "C" int ipp_lanczos_coeffs_host(int in_size, int out_size, int32_t* bounds_host, int32_t* kk_host) {
    using namespace ipp;
    IPP_REQUIRE(in_size > 0 && out_size > 0 && bounds_host && kk_host, IPP_ERR_BAD_ARG, "ipp_lanczos_coeffs_host: bad arguments");
    double scale, filterscale, support;
    int ksize;
    lanczos_scale(in_size, out_size, &scale, &filterscale, &support, &ksize);
    const double ss = 1.0 / filterscale;
    double wbuf[4096];
    IPP_REQUIRE(ksize <= 4096, IPP_ERR_UNSUPPORTED, "ipp_lanczos_coeffs_host: ksize %d too large", ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        int32_t* k = kk_host + (int64_t)xx * ksize;
        for (int x = 0; x < xmax; ++x) {
            const double w = lanczos_filter((x + xmin - center + 0.5) * ss);
            wbuf[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            double w = wbuf[x];
            if (ww != 0.0) w /= ww;
            k[x] = (w < 0) ? (int)(-0.5 + w * (1 << kPrecisionBits)) : (int)(0.5 + w * (1 << kPrecisionBits));
        }
        for (int x = xmax; x < ksize; ++x) k[x] = 0;
        bounds_host[2 * xx] = xmin;
        bounds_host[2 * xx + 1] = xmax;
    }
    return IPP_OK;
}

extern "C" int ipp_resample_rgba_u8(const uint8_t* src, int64_t src_pitch, int src_w, int src_h,
                                    uint8_t* dst, int64_t dst_pitch, int dst_w, int dst_h,
                                    const int32_t* h_bounds, const int32_t* h_kk, int h_ksize,
                                    const int32_t* v_bounds, const int32_t* v_kk, int v_ksize,
                                    uint8_t* tmp, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    IPP_REQUIRE(src && dst, IPP_ERR_BAD_ARG, "ipp_resample_rgba_u8: NULL image pointer");
    IPP_REQUIRE(src_w > 0 && src_h > 0 && dst_w > 0 && dst_h > 0, IPP_ERR_BAD_SHAPE,
                "ipp_resample_rgba_u8: empty image (%dx%d -> %dx%d)", src_w, src_h, dst_w, dst_h);
    IPP_REQUIRE((src_pitch & 3) == 0 && (dst_pitch & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 3) == 0,
                IPP_ERR_BAD_ARG, "ipp_resample_rgba_u8: RGBA images must be 4-byte aligned");
    const bool do_h = h_ksize > 0, do_v = v_ksize > 0;
    IPP_REQUIRE(do_h || src_w == dst_w, IPP_ERR_BAD_ARG, "ipp_resample_rgba_u8: horizontal tables missing");
    IPP_REQUIRE(do_v || src_h == dst_h, IPP_ERR_BAD_ARG, "ipp_resample_rgba_u8: vertical tables missing");
    IPP_REQUIRE(!do_h || (h_bounds && h_kk), IPP_ERR_BAD_ARG, "ipp_resample_rgba_u8: NULL horizontal tables");
    IPP_REQUIRE(!do_v || (v_bounds && v_kk), IPP_ERR_BAD_ARG, "ipp_resample_rgba_u8: NULL vertical tables");
    if (!do_h && !do_v) {   // Image.resize returns a copy when the size is unchanged
        IPP_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, (size_t)src_w * 4, src_h, cudaMemcpyDeviceToDevice, st));
        return IPP_OK;
    }
    if (do_h && do_v) {
        IPP_REQUIRE(tmp && (((uintptr_t)tmp) & 3) == 0, IPP_ERR_WORKSPACE, "ipp_resample_rgba_u8: tmp scratch missing");
        const int64_t tpitch = (int64_t)dst_w * 4;
        dim3 g1(ceil_div(dst_w, 32), ceil_div(src_h, 8));
        { ipp::LaunchTimer _lt("resample_h_kernel", st); resample_h_kernel<true, false><<<g1, 256, 0, st>>>(src, src_pitch, src_h, tmp, tpitch, dst_w, h_bounds, h_kk, h_ksize); }
        IPP_LAUNCH_CHECK();
        dim3 g2(ceil_div(dst_w, 32), ceil_div(dst_h, 8));
        { ipp::LaunchTimer _lt("resample_v_kernel", st); resample_v_kernel<false, true><<<g2, 256, 0, st>>>(tmp, tpitch, dst_w, dst, dst_pitch, dst_h, v_bounds, v_kk, v_ksize); }
        IPP_LAUNCH_CHECK();
    } else if (do_h) {
        dim3 g1(ceil_div(dst_w, 32), ceil_div(src_h, 8));
        { ipp::LaunchTimer _lt("resample_h_kernel", st); resample_h_kernel<true, true><<<g1, 256, 0, st>>>(src, src_pitch, src_h, dst, dst_pitch, dst_w, h_bounds, h_kk, h_ksize); }
        IPP_LAUNCH_CHECK();
    } else {
        dim3 g2(ceil_div(dst_w, 32), ceil_div(dst_h, 8));
        { ipp::LaunchTimer _lt("resample_v_kernel", st); resample_v_kernel<true, true><<<g2, 256, 0, st>>>(src, src_pitch, dst_w, dst, dst_pitch, dst_h, v_bounds, v_kk, v_ksize); }
        IPP_LAUNCH_CHECK();
    }
    return IPP_OK;
}
