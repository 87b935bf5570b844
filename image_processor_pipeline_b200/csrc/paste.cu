// paste.cu — alpha paste of an RGBA sprite onto an RGB background with the final
// crop folded into the same pass.
//
// Reference behaviour replaced: bg.copy().paste(sprite, (x, y), sprite)
// (transforms/overlays.py:138-139; libImaging paste_mask_RGBA:
// out = DIV255(bg*(255-a) + spr*a)) followed by image[y0:y0+cs, x0:x0+cs]
// (transforms/crop_square.py:196).
//
// Streaming kernel: one CTA owns a few output rows; the bulk-copy engine stages the
// background row windows (arbitrary byte alignment: crop_x0*3) into shared memory;
// each thread emits aligned 16-byte output vectors, blending only the vectors that
// intersect the sprite rectangle (the sprite is small and L2 resident).
#include "staging.cuh"

namespace ipp {

struct PasteParams {
    const uint8_t* bg;
    int64_t bg_pitch;
    const uint8_t* sprite;
    int64_t sprite_pitch;
    int sprite_w, sprite_h, pos_x, pos_y;
    int crop_x0, crop_y0, crop_w, crop_h;
    uint8_t* dst;
    int64_t dst_pitch;
    int rows_per_cta;
    uint32_t smem_row_stride;
};

constexpr int kPasteThreads = 256;

// 16 output bytes cut from two aligned vectors of the staged window; WS = word shift (0..3),
// bs = byte shift inside the word in bits — both uniform for a whole row.
template <int WS>
__device__ __forceinline__ uint4 cut16(const uint4& lo, const uint4& hi, uint32_t bs) {
    uint32_t v0, v1, v2, v3, v4;
    if (WS == 0) { v0 = lo.x; v1 = lo.y; v2 = lo.z; v3 = lo.w; v4 = hi.x; }
    else if (WS == 1) { v0 = lo.y; v1 = lo.z; v2 = lo.w; v3 = hi.x; v4 = hi.y; }
    else if (WS == 2) { v0 = lo.z; v1 = lo.w; v2 = hi.x; v3 = hi.y; v4 = hi.z; }
    else { v0 = lo.w; v1 = hi.x; v2 = hi.y; v3 = hi.z; v4 = hi.w; }
    return make_uint4(__funnelshift_r(v0, v1, bs), __funnelshift_r(v1, v2, bs), __funnelshift_r(v2, v3, bs),
                      __funnelshift_r(v3, v4, bs));
}

// aligned 16-byte body of one output row (background only), lanes striding over the chunks
template <int WS>
__device__ __forceinline__ void copy_row_body(const uint8_t* win, uint32_t off0, uint8_t* d, uint32_t head, uint32_t nb, int lane) {
    const uint32_t bs = (off0 & 3u) * 8u;
    const uint8_t* src = win + (off0 & ~15u);
    const bool aligned = (off0 & 15u) == 0u;
    for (uint32_t c = lane; c < nb; c += 32) {
        const uint4 lo = *reinterpret_cast<const uint4*>(src + (c << 4));
        uint4 val = lo;
        if (!aligned) {
            const uint4 hi = *reinterpret_cast<const uint4*>(src + (c << 4) + 16);
            val = cut16<WS>(lo, hi, bs);
        }
        stg_stream16(d + head + (c << 4), val);
    }
}

// emit rows [row0, row0+nrows) of the cropped composite from staged background windows.
// One warp per row (row-uniform alignment and pointers live in uniform registers), two passes:
//  1. the background row streams out as aligned 16-byte vectors (plus < 16 head / tail bytes, one lane each);
//  2. rows that cross the sprite: after a __syncwarp (orders the stores of pass 1 before those of pass 2 for
//     the whole warp) the lanes walk the sprite's pixels of this row — coalesced sprite loads, background bytes
//     from the staged window, libImaging's DIV255 blend, three byte stores.  Transparent pixels (alpha 0:
//     DIV255(255 * bg) = bg) are skipped, opaque ones (alpha 255: DIV255(255 * s) = s) copied.
// The sprite covers a few percent of the image, so pass 2 re-touches little; keeping the blend out of the
// vector loop removes the per-byte phase arithmetic (pixels are 3 bytes, vectors 16) that dominated it.
#ifndef IPP_SPR_FLY
#define IPP_SPR_FLY 4
#endif
constexpr int kSprFly = IPP_SPR_FLY;          // sprite pixels per lane in flight
__device__ __forceinline__ void paste_rows_from_smem(const PasteParams& p, const uint8_t* smem, uint32_t smem_row_stride,
                                                     int row0, int nrows, int tid) {
    const uint32_t rowbytes = (uint32_t)p.crop_w * 3u;
    const int lane = tid & 31, warp = tid >> 5;
    // output pixel range of a row that can intersect the sprite
    const int hit_x0 = max(p.pos_x - p.crop_x0, 0);
    const int hit_x1 = min(p.pos_x + p.sprite_w - p.crop_x0, p.crop_w);      // exclusive
    for (int r = warp; r < nrows; r += kPasteThreads / 32) {
        const int y = row0 + r;
        uint8_t* d = p.dst + (int64_t)y * p.dst_pitch;
        const uint8_t* win = smem + (size_t)r * smem_row_stride;
        const uint8_t* s = p.bg + (int64_t)(p.crop_y0 + y) * p.bg_pitch + (int64_t)p.crop_x0 * 3;
        const uint32_t shift = (uint32_t)((uintptr_t)s & 15);
        const RowSplit sp = split_row(d, rowbytes);
        // sprite pixels of this row: the first kSprFly * 32 are requested BEFORE the row is streamed out, so that their L2 round
        // trip overlaps the copy instead of heading a dependent load -> blend -> store loop (one round trip per 32 pixels)
        const int sy = p.crop_y0 + y - p.pos_y;
        const bool has_sprite = sy >= 0 && sy < p.sprite_h && hit_x1 > hit_x0;
        const uint32_t* srow = reinterpret_cast<const uint32_t*>(p.sprite + (int64_t)(has_sprite ? sy : 0) * p.sprite_pitch) + (p.crop_x0 - p.pos_x);
        uint32_t v[kSprFly];
#pragma unroll
        for (int k = 0; k < kSprFly; ++k) {
            const int x = hit_x0 + 32 * k + lane;
            v[k] = (has_sprite && x < hit_x1) ? __ldg(srow + x) : 0u;        // alpha 0: nothing to blend
        }
        const uint32_t off0 = shift + sp.head;            // window offset of the first aligned output vector
        switch ((off0 & 15u) >> 2) {
            case 0: copy_row_body<0>(win, off0, d, sp.head, sp.nb, lane); break;
            case 1: copy_row_body<1>(win, off0, d, sp.head, sp.nb, lane); break;
            case 2: copy_row_body<2>(win, off0, d, sp.head, sp.nb, lane); break;
            default: copy_row_body<3>(win, off0, d, sp.head, sp.nb, lane); break;
        }
        if ((uint32_t)lane < sp.head) d[lane] = win[shift + lane];
        if (lane >= 16 && (uint32_t)(lane - 16) < sp.tail) {
            const uint32_t b = sp.head + (sp.nb << 4) + (uint32_t)(lane - 16);
            d[b] = win[shift + b];
        }
        if (has_sprite) {
            __syncwarp();
            auto blend = [&](uint32_t spx, int x) {
                const uint32_t a = spx >> 24;
                if (a == 0u) return;
                uint32_t o0 = spx & 255u, o1 = (spx >> 8) & 255u, o2 = (spx >> 16) & 255u;
                if (a != 255u) {
                    const uint8_t* b = win + shift + 3 * x;
                    const uint32_t na = 255u - a;
                    o0 = div255((uint32_t)b[0] * na + o0 * a);
                    o1 = div255((uint32_t)b[1] * na + o1 * a);
                    o2 = div255((uint32_t)b[2] * na + o2 * a);
                }
                uint8_t* o = d + 3 * x;
                o[0] = (uint8_t)o0; o[1] = (uint8_t)o1; o[2] = (uint8_t)o2;
            };
            for (int x0 = hit_x0; x0 < hit_x1; x0 += 32 * kSprFly) {
                uint32_t nv[kSprFly];
#pragma unroll
                for (int k = 0; k < kSprFly; ++k) {                           // next batch in flight while this one is blended
                    const int x = x0 + 32 * (kSprFly + k) + lane;
                    nv[k] = x < hit_x1 ? __ldg(srow + x) : 0u;
                }
#pragma unroll
                for (int k = 0; k < kSprFly; ++k) blend(v[k], x0 + 32 * k + lane);
#pragma unroll
                for (int k = 0; k < kSprFly; ++k) v[k] = nv[k];
            }
        }
    }
}

__global__ void __launch_bounds__(kPasteThreads) paste_kernel(const PasteParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * p.rows_per_cta;
    const int nrows = min(p.rows_per_cta, p.crop_h - row0);
    const uint32_t rowbytes = (uint32_t)p.crop_w * 3u;

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        uint32_t total = 0;
        for (int r = 0; r < nrows; ++r) {
            const uint8_t* s = p.bg + (int64_t)(p.crop_y0 + row0 + r) * p.bg_pitch + (int64_t)p.crop_x0 * 3;
            total += make_window(s, rowbytes).bytes;
        }
        mbar_expect_tx(&bar, total);
        for (int r = 0; r < nrows; ++r) {
            const uint8_t* s = p.bg + (int64_t)(p.crop_y0 + row0 + r) * p.bg_pitch + (int64_t)p.crop_x0 * 3;
            const Window w = make_window(s, rowbytes);
            bulk_g2s(smem + (size_t)r * p.smem_row_stride, w.base, w.bytes, &bar);
        }
    }
    mbar_wait(&bar, 0);

    paste_rows_from_smem(p, smem, p.smem_row_stride, row0, nrows, tid);
}

}  // namespace ipp

extern "C" int ipp_paste_rgba_on_rgb_u8(const uint8_t* bg, int64_t bg_pitch, int bg_w, int bg_h,
                                        const uint8_t* sprite, int64_t sprite_pitch, int sprite_w, int sprite_h,
                                        int pos_x, int pos_y,
                                        int crop_x0, int crop_y0, int crop_w, int crop_h,
                                        uint8_t* dst, int64_t dst_pitch, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(bg && dst, IPP_ERR_BAD_ARG, "ipp_paste_rgba_on_rgb_u8: NULL image pointer");
    IPP_REQUIRE(sprite || sprite_w == 0 || sprite_h == 0, IPP_ERR_BAD_ARG, "ipp_paste_rgba_on_rgb_u8: NULL sprite");
    IPP_REQUIRE(bg_w > 0 && bg_h > 0 && crop_x0 >= 0 && crop_y0 >= 0 && crop_w >= 0 && crop_h >= 0 &&
                    crop_x0 + crop_w <= bg_w && crop_y0 + crop_h <= bg_h,
                IPP_ERR_BAD_SHAPE, "ipp_paste_rgba_on_rgb_u8: crop (%d,%d,%d,%d) outside %dx%d", crop_x0, crop_y0,
                crop_w, crop_h, bg_w, bg_h);
    IPP_REQUIRE(bg_pitch >= (int64_t)bg_w * 3 && dst_pitch >= (int64_t)crop_w * 3, IPP_ERR_BAD_SHAPE,
                "ipp_paste_rgba_on_rgb_u8: pitch smaller than a row");
    IPP_REQUIRE(sprite_w == 0 || sprite_h == 0 ||
                    ((sprite_pitch & 3) == 0 && (((uintptr_t)sprite) & 3) == 0 && sprite_pitch >= (int64_t)sprite_w * 4),
                IPP_ERR_BAD_ARG, "ipp_paste_rgba_on_rgb_u8: sprite must be 4-byte aligned RGBA");
    if (crop_w == 0 || crop_h == 0) return IPP_OK;
    PasteParams p{};
    p.bg = bg; p.bg_pitch = bg_pitch;
    p.sprite = sprite; p.sprite_pitch = sprite_pitch;
    p.sprite_w = sprite_w > 0 ? sprite_w : 0; p.sprite_h = sprite_h > 0 ? sprite_h : 0;
    p.pos_x = pos_x; p.pos_y = pos_y;
    p.crop_x0 = crop_x0; p.crop_y0 = crop_y0; p.crop_w = crop_w; p.crop_h = crop_h;
    p.dst = dst; p.dst_pitch = dst_pitch;
    const uint32_t rowbytes = (uint32_t)crop_w * 3u;
    p.smem_row_stride = staged_row_stride(rowbytes);
    IPP_REQUIRE(p.smem_row_stride <= 200 * 1024, IPP_ERR_UNSUPPORTED, "ipp_paste_rgba_on_rgb_u8: row too long");
    int rows = (int)(16384 / p.smem_row_stride);
    if (rows < 1) rows = 1;
    if (rows > crop_h) rows = crop_h;
    const int sms = sm_count();
    while (rows > 1 && ceil_div(crop_h, rows) < 4 * sms) rows = (rows + 1) / 2;
    p.rows_per_cta = rows;
    const size_t smem = (size_t)rows * p.smem_row_stride;
    if (smem > 48 * 1024)
        IPP_CUDA(cudaFuncSetAttribute(paste_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { ipp::LaunchTimer _lt("paste_kernel", (cudaStream_t)stream); paste_kernel<<<ceil_div(crop_h, rows), kPasteThreads, smem, (cudaStream_t)stream>>>(p); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}
