// hsv.cu — HSV filter-list keying: BGR -> BGRA with alpha = ~OR_k(inRange_k & zone_k).
//
// Reference behaviour replaced: transforms/filtres_liste.py:90-134 (cvtColor(BGR2HSV),
// per filter inRange & zone, OR, bitwise_not, merge(b, g, r, alpha)); bit-exact.
//
// One CTA owns a few rows: the bulk-copy engine stages the BGR row windows in shared
// memory (3-byte pixels start at any alignment); each thread converts 4 pixels
// (12 staged bytes) and stores one aligned 16-byte BGRA vector.
#include "hsv.cuh"
#include "staging.cuh"

#include <cmath>
#include <mutex>

namespace ipp {

int ensure_hsv_tables() {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return 0;
    int sdiv[256], hdiv[256];
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
        sdiv[i] = (int)std::nearbyint((double)(255 << 12) / (1.0 * i));      // cvRound: half to even
        hdiv[i] = (int)std::nearbyint((double)(180 << 12) / (6.0 * i));
    }
    e = cudaMemcpyToSymbol(g_hsv_sdiv, sdiv, sizeof(sdiv));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(g_hsv_hdiv, hdiv, sizeof(hdiv));
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) done[dev] = true;
    return 0;
}

struct HsvParams {
    const uint8_t* src;
    int64_t src_pitch, src_img_stride;
    uint8_t* dst;
    int64_t dst_pitch, dst_img_stride;
    int w, h;
    int rows_per_cta;
    uint32_t smem_row_stride;
    HsvFilterSet f;
};

constexpr int kHsvThreads = 256;

__global__ void __launch_bounds__(kHsvThreads) hsv_alpha_kernel(const HsvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ int sdiv[256], hdiv[256];
    const int tid = threadIdx.x;
    const int img = blockIdx.y;
    const int row0 = blockIdx.x * p.rows_per_cta;
    const int nrows = min(p.rows_per_cta, p.h - row0);
    const uint32_t rowbytes = (uint32_t)p.w * 3u;
    const uint8_t* src_img = p.src + (int64_t)img * p.src_img_stride;

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        uint32_t total = 0;
        for (int r = 0; r < nrows; ++r) total += make_window(src_img + (int64_t)(row0 + r) * p.src_pitch, rowbytes).bytes;
        mbar_expect_tx(&bar, total);
        for (int r = 0; r < nrows; ++r) {
            const Window w = make_window(src_img + (int64_t)(row0 + r) * p.src_pitch, rowbytes);
            bulk_g2s(smem + (size_t)r * p.smem_row_stride, w.base, w.bytes, &bar);
        }
    }
    fill_hsv_tables(sdiv, hdiv, tid, kHsvThreads);   // overlaps the bulk copy
    __syncthreads();
    mbar_wait(&bar, 0);

    const int gpr = (p.w + 3) >> 2;    // groups of 4 pixels per row
    const int items = nrows * gpr;
    for (int it = tid; it < items; it += kHsvThreads) {
        const int r = it / gpr;
        const int gi = it - r * gpr;
        const int y = row0 + r;
        const int x = gi * 4;
        const uint8_t* s = src_img + (int64_t)y * p.src_pitch;
        const uint32_t shift = (uint32_t)((uintptr_t)s & 15);
        const uint4 raw = window_load16(smem + (size_t)r * p.smem_row_stride, shift + (uint32_t)x * 3u);
        // 12 bytes = 4 BGR pixels: b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3
        const uint32_t w0 = raw.x, w1 = raw.y, w2 = raw.z;
        uint32_t bgr[4];
        bgr[0] = w0 & 0x00ffffffu;
        bgr[1] = __funnelshift_r(w0, w1, 24) & 0x00ffffffu;
        bgr[2] = __funnelshift_r(w1, w2, 16) & 0x00ffffffu;
        bgr[3] = w2 >> 8;
        uint32_t out[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int b = bgr[k] & 255u, g = (bgr[k] >> 8) & 255u, rr = bgr[k] >> 16;
            const uint32_t X = bgr_to_hsv_packed(b, g, rr, sdiv, hdiv);
            const bool ex = hsv_excluded(X, x + k, y, p.f);
            out[k] = bgr[k] | (ex ? 0u : 0xff000000u);
        }
        uint8_t* d = p.dst + (int64_t)img * p.dst_img_stride + (int64_t)y * p.dst_pitch + (int64_t)x * 4;
        if (x + 3 < p.w && (((uintptr_t)d) & 15) == 0) {
            stg_stream16(d, make_uint4(out[0], out[1], out[2], out[3]));
        } else {
            for (int k = 0; k < 4 && x + k < p.w; ++k) {
                uint8_t* dk = d + 4 * k;
                dk[0] = (uint8_t)out[k]; dk[1] = (uint8_t)(out[k] >> 8);
                dk[2] = (uint8_t)(out[k] >> 16); dk[3] = (uint8_t)(out[k] >> 24);
            }
        }
    }
}

}  // namespace ipp

extern "C" int ipp_hsv_ranges_to_alpha_u8(const uint8_t* bgr, int64_t src_pitch, int w, int h,
                                          const ipp_hsv_range* ranges_host, int n_ranges,
                                          uint8_t* bgra, int64_t dst_pitch,
                                          int n, int64_t src_image_stride, int64_t dst_image_stride, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(bgr && bgra && ranges_host, IPP_ERR_BAD_ARG, "ipp_hsv_ranges_to_alpha_u8: NULL pointer");
    IPP_REQUIRE(n_ranges >= 1 && n_ranges <= IPP_MAX_HSV_RANGES, IPP_ERR_BAD_ARG,
                "ipp_hsv_ranges_to_alpha_u8: n_ranges=%d not in 1..%d", n_ranges, IPP_MAX_HSV_RANGES);
    IPP_REQUIRE(n >= 1 && w >= 0 && h >= 0 && src_pitch >= (int64_t)w * 3 && dst_pitch >= (int64_t)w * 4,
                IPP_ERR_BAD_SHAPE, "ipp_hsv_ranges_to_alpha_u8: bad geometry");
    if (w == 0 || h == 0) return IPP_OK;
    const int te = ensure_hsv_tables();
    IPP_REQUIRE(te == 0, te, "ipp_hsv_ranges_to_alpha_u8: table upload failed (%d)", te);
    HsvParams p{};
    p.src = bgr; p.src_pitch = src_pitch; p.src_img_stride = src_image_stride;
    p.dst = bgra; p.dst_pitch = dst_pitch; p.dst_img_stride = dst_image_stride;
    p.w = w; p.h = h;
    p.f = pack_filters(ranges_host, n_ranges, w, h);
    const uint32_t rowbytes = (uint32_t)w * 3u;
    p.smem_row_stride = staged_row_stride(rowbytes);
    IPP_REQUIRE(p.smem_row_stride <= 200 * 1024, IPP_ERR_UNSUPPORTED, "ipp_hsv_ranges_to_alpha_u8: row too long");
    int rows = (int)(16384 / p.smem_row_stride);
    if (rows < 1) rows = 1;
    if (rows > h) rows = h;
    const int sms = sm_count();
    while (rows > 1 && (int64_t)ceil_div(h, rows) * n < 4LL * sms) rows = (rows + 1) / 2;
    p.rows_per_cta = rows;
    const size_t smem = (size_t)rows * p.smem_row_stride;
    if (smem > 40 * 1024)
        IPP_CUDA(cudaFuncSetAttribute(hsv_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(h, rows), n);
    { ipp::LaunchTimer _lt("hsv_alpha_kernel", (cudaStream_t)stream); hsv_alpha_kernel<<<grid, kHsvThreads, smem, (cudaStream_t)stream>>>(p); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}
