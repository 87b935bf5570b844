// batch.cu — batched entry points of the stand-alone operators: n images per call, per-item
// parameter records on the device, flat tile lists (one launch per batch whatever the sizes).
//
// Reference behaviour replaced, per image of a directory a ProcessingStep maps a transform over
// (pipeline.py:304-437):
//   * rotations.py:55,96-99   convert("RGBA") + Image.rotate(expand=True) + getbbox
//   * overlays.py:104-149     LANCZOS resize of an RGBA sprite + paste onto an RGB background
// The single-image entry points (rotate.cu, resample.cu, paste.cu) stay; a 1080p image is 2-3 us
// of HBM time, so only a batch can come near the memory roofline.
#include "staging.cuh"

namespace ipp {

// ===================================================================== rotation
// CTA = one output tile of a canvas: the rotated source footprint of a compact tile is compact, so the gathers of
// a CTA share L1 lines and every source byte leaves HBM once (ncu: DRAM read = the sources, once).  Canvas rows are
// padded to a multiple of four pixels: every lane stores one aligned 16-byte vector.
struct RotBatchParams {
    const uint8_t* src;
    int64_t src_pitch, src_stride;
    int src_w, src_h;
    const ipp_rot_item* items;
    int n_items;
    uint8_t* dst_pool;
    int32_t* bbox;
};

// pixel (xin, yin) of a CH-channel image as RGBA (CH == 3: alpha 255, bytes from the aligned word(s) holding the pixel)
template <int CH>
__device__ __forceinline__ uint32_t fetch_rgba(const uint8_t* img, int64_t pitch, int xin, int yin) {
    if (CH == 4) return __ldg(reinterpret_cast<const uint32_t*>(img + (int64_t)yin * pitch) + xin);
    const uintptr_t a = (uintptr_t)(img + (int64_t)yin * pitch + xin * 3);
    const uint32_t* aw = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const int bsh = (int)(a & 3) * 8;
    const uint32_t w0 = __ldg(aw), w1 = bsh > 8 ? __ldg(aw + 1) : 0u;      // every word read holds a byte of the pixel
    return (__funnelshift_r(w0, w1, bsh) & 0x00ffffffu) | 0xff000000u;
}

// arithmetic without fma contraction in either precision (libImaging evaluates v1 + (v2 - v1) * d as written)
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
template <typename T>
__device__ __forceinline__ void premult_t(uint32_t v, T out[4]) {
    const uint32_t a = v >> 24;
    out[0] = (T)div255((v & 255u) * a);
    out[1] = (T)div255(((v >> 8) & 255u) * a);
    out[2] = (T)div255(((v >> 16) & 255u) * a);
    out[3] = (T)a;
}
template <int CH> struct InterpType { typedef float type; };      // opaque sources: single precision, <= 1 LSB
template <> struct InterpType<4> { typedef double type; };        // real alpha: libImaging's double (un-premultiply amplifies errors)

// bilinear sample of canvas pixel (x, y) (MODE 1, affine): libImaging's affine_transform + bilinear_filter32RGB on premultiplied
// alpha, source coordinates in double like C evaluates them (no fma contraction)
template <int CH>
__device__ __forceinline__ uint32_t bilinear_px(const uint8_t* img, int64_t pitch, int src_w, int src_h, const double* m, int x, int y) {
    typedef typename InterpType<CH>::type T;
    const double xo = (double)x + 0.5, yo = (double)y + 0.5;
    double xin = __dadd_rn(__dadd_rn(__dmul_rn(m[0], xo), __dmul_rn(m[1], yo)), m[2]);
    double yin = __dadd_rn(__dadd_rn(__dmul_rn(m[3], xo), __dmul_rn(m[4], yo)), m[5]);
    if (!(xin >= 0.0 && xin < (double)src_w && yin >= 0.0 && yin < (double)src_h)) return 0u;
    xin -= 0.5;
    yin -= 0.5;
    const int xi = (int)floor(xin), yi = (int)floor(yin);
    const T dx = (T)(xin - (double)xi), dy = (T)(yin - (double)yi);
    const int xa = min(max(xi, 0), src_w - 1), xb = min(max(xi + 1, 0), src_w - 1);
    const int ya = min(max(yi, 0), src_h - 1), yb = min(max(yi + 1, 0), src_h - 1);
    T s00[4], s01[4], s10[4], s11[4];
    premult_t<T>(fetch_rgba<CH>(img, pitch, xa, ya), s00);
    premult_t<T>(fetch_rgba<CH>(img, pitch, xb, ya), s01);
    premult_t<T>(fetch_rgba<CH>(img, pitch, xa, yb), s10);
    premult_t<T>(fetch_rgba<CH>(img, pitch, xb, yb), s11);
    uint32_t c[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const T v1 = add_rn(s00[b], mul_rn(sub_rn(s01[b], s00[b]), dx));
        const T v2 = add_rn(s10[b], mul_rn(sub_rn(s11[b], s10[b]), dx));
        const T v = add_rn(v1, mul_rn(sub_rn(v2, v1), dy));
        c[b] = (uint32_t)min(max((int)v, 0), 255);
    }
    const uint32_t a = c[3];
    if (a != 0u && a != 255u) {
#pragma unroll
        for (int b = 0; b < 3; ++b) c[b] = min(255u, (255u * c[b]) / a);
    }
    return c[0] | (c[1] << 8) | (c[2] << 16) | (a << 24);
}

// CTA = 64 x 256 output tile of one canvas: 16 lanes x 4 px across, 16 rows per pass, 16 passes.  Tall tiles because the head of
// a CTA (record search, record fields: three dependent L2 round trips) cost as much as two passes of work: 64 x 32 tiles ran at
// 0.46 of the copy peak whatever the per-pixel path cost, 64 x 256 at 0.60.  The per-pixel path is kept minimal: tile classes
// (below), 32-bit source offsets, incremental fixed-point coordinates, the alpha bbox from a 4-bit mask per thread, one block
// reduction per tile whose global atomics are skipped when they cannot change the box.
#ifndef IPP_ROT_MINB
#define IPP_ROT_MINB 8
#endif
#ifndef IPP_ROT_UNROLL
#define IPP_ROT_UNROLL 2
#endif
#ifndef IPP_ROT_MINB_BIL
#define IPP_ROT_MINB_BIL 4
#endif
constexpr int kRotUnroll = IPP_ROT_UNROLL;
// nearest: 8 CTAs per SM (32 registers) — the kernel waits for its source loads, resident warps are what hides them
template <int MODE, int CH>
__global__ void __launch_bounds__(256, MODE == 0 ? IPP_ROT_MINB : IPP_ROT_MINB_BIL) rotate_batch_kernel(const RotBatchParams p) {
    __shared__ int s_rec;
    __shared__ int s_bb[4];
    if (threadIdx.x < 32) {
        const int found = find_record(p.items, p.n_items, (int)blockIdx.x, [](const ipp_rot_item& r) { return r.tile_start; });
        if (threadIdx.x == 0) {
            s_rec = found;
            s_bb[0] = INT_MAX; s_bb[1] = INT_MAX; s_bb[2] = 0; s_bb[3] = 0;
        }
    }
    __syncthreads();
    const int rec = s_rec;
    const ipp_rot_item& it = p.items[rec];
    const int kind = it.kind, a0 = it.a0, a1 = it.a1, a3 = it.a3, a4 = it.a4;
    const int dst_w = it.dst_w, dst_h = it.dst_h, tiles_x = it.tiles_x;
    const int tile = (int)blockIdx.x - it.tile_start;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int tid = threadIdx.x;
    const int x = tx * kRotTileW + (tid & 15) * 4;
    const int wpad = (dst_w + 3) & ~3;
    const uint8_t* img = p.src + (int64_t)it.src * p.src_stride;
    const uint32_t pitch32 = (uint32_t)p.src_pitch;
    const uint32_t img_lsb = (uint32_t)((uintptr_t)img & 3u);
    const uint8_t* img_al = img - img_lsb;
    uint8_t* dbase = p.dst_pool + it.dst_off;
    int minx = INT_MAX, miny = INT_MAX, maxx1 = 0, maxy1 = 0;
    // The fixed-point source coordinates are affine in (x, y): their extremes over the tile sit at its corners.  A tile whose
    // image lies beyond one edge of the source is transparent (43 % of an expanded canvas on average: zeros, no per-pixel work);
    // a tile whose image lies inside needs no bounds test.  0 = outside, 1 = mixed, 2 = inside.
    int tile_class = 1;
    if (MODE == 0 && kind == 0) {
        const int x_lo = tx * kRotTileW, y_lo = ty * kRotTileH;
        const int ex = min(x_lo + kRotTileW, dst_w) - 1 - x_lo, ey = min(y_lo + kRotTileH, dst_h) - 1 - y_lo;
        const int bx = it.a2 + y_lo * a1 + x_lo * a0, by = it.a5 + y_lo * a4 + x_lo * a3;
        const int xx_min = bx + min(ex * a0, 0) + min(ey * a1, 0), xx_max = bx + max(ex * a0, 0) + max(ey * a1, 0);
        const int yy_min = by + min(ex * a3, 0) + min(ey * a4, 0), yy_max = by + max(ex * a3, 0) + max(ey * a4, 0);
        const int lim_x = p.src_w << 16, lim_y = p.src_h << 16;       // (unsigned)(v >> 16) < n  <=>  0 <= v < n << 16
        if (xx_max < 0 || xx_min >= lim_x || yy_max < 0 || yy_min >= lim_y) tile_class = 0;
        else if (xx_min >= 0 && xx_max < lim_x && yy_min >= 0 && yy_max < lim_y) tile_class = 2;
    }
#pragma unroll kRotUnroll
    for (int pass = 0; pass < kRotTileH / 16; ++pass) {
        const int y = ty * kRotTileH + (tid >> 4) + 16 * pass;
        if (y >= dst_h || x >= wpad) continue;
        uint32_t px[4] = {0u, 0u, 0u, 0u};
        if (MODE == 0 && kind == 0 && tile_class == 0) {
            stg_stream16(dbase + ((int64_t)y * wpad + x) * 4, make_uint4(0u, 0u, 0u, 0u));
            continue;
        }
        if (MODE == 0 && kind == 0) {
            int xx = it.a2 + y * a1 + x * a0, yy = it.a5 + y * a4 + x * a3;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xin = xx >> 16, yin = yy >> 16;
                xx += a0;
                yy += a3;
                const bool inside = (tile_class == 2 || (((unsigned)xin < (unsigned)p.src_w) & ((unsigned)yin < (unsigned)p.src_h))) & (x + k < dst_w);
                if (inside) {
                    if (CH == 4) {
                        px[k] = __ldg(reinterpret_cast<const uint32_t*>(img + (uint32_t)yin * pitch32) + xin);
                    } else {
                        const uint32_t off = (uint32_t)yin * pitch32 + (uint32_t)xin * 3u + img_lsb;
                        const uint32_t* aw = reinterpret_cast<const uint32_t*>(img_al + (off & ~3u));
                        const uint32_t bsh = (off & 3u) * 8u;
                        const uint32_t w0 = __ldg(aw), w1 = bsh > 8u ? __ldg(aw + 1) : 0u;     // every word read holds a byte of the pixel
                        px[k] = __funnelshift_r(w0, w1, bsh) | 0xff000000u;
                    }
                }
            }
        } else if (MODE == 0 || kind != 0) {
            const SrcMap m{kind, a0, a1, it.a2, a3, a4, it.a5, p.src_w, p.src_h};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int xin, yin;
                if (x + k < dst_w && map_canvas(m, x + k, y, xin, yin)) px[k] = fetch_rgba<CH>(img, p.src_pitch, xin, yin);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (x + k < dst_w) px[k] = bilinear_px<CH>(img, p.src_pitch, p.src_w, p.src_h, it.m, x + k, y);
        }
        stg_stream16(dbase + ((int64_t)y * wpad + x) * 4, make_uint4(px[0], px[1], px[2], px[3]));
        const uint32_t nz = (px[0] >> 24 ? 1u : 0u) | (px[1] >> 24 ? 2u : 0u) | (px[2] >> 24 ? 4u : 0u) | (px[3] >> 24 ? 8u : 0u);
        if (nz) {
            minx = min(minx, x + __ffs(nz) - 1);
            maxx1 = max(maxx1, x + 32 - __clz(nz));
            miny = min(miny, y);
            maxy1 = max(maxy1, y + 1);
        }
    }
    // block reduction (s_bb was initialised before the first barrier); global atomics only where they can change the box
    minx = warp_min(minx); miny = warp_min(miny); maxx1 = warp_max(maxx1); maxy1 = warp_max(maxy1);
    if ((tid & 31) == 0 && maxx1 > 0) {
        atomicMin(&s_bb[0], minx); atomicMin(&s_bb[1], miny); atomicMax(&s_bb[2], maxx1); atomicMax(&s_bb[3], maxy1);
    }
    __syncthreads();
    if (tid == 0 && s_bb[2] > 0) {
        int32_t* bb = p.bbox + 4 * rec;
        if (s_bb[0] < __ldcg(bb + 0)) atomicMin(bb + 0, s_bb[0]);      // the box only grows: a stale read can only cause a redundant atomic
        if (s_bb[1] < __ldcg(bb + 1)) atomicMin(bb + 1, s_bb[1]);
        if (s_bb[2] > __ldcg(bb + 2)) atomicMax(bb + 2, s_bb[2]);
        if (s_bb[3] > __ldcg(bb + 3)) atomicMax(bb + 3, s_bb[3]);
    }
}

// ===================================================================== overlay: RGBA sprite -> premultiplied planes
// thread = four pixels of one sprite row (one 16-byte load, one 32-bit store per plane)
__global__ void __launch_bounds__(256) sprite_planes_kernel(const uint8_t* __restrict__ sprites, int64_t pitch, int64_t stride,
                                                            const ipp_chain_item* __restrict__ items, uint8_t* ws) {
    const ipp_chain_item& it = items[blockIdx.y];
    const int sw = it.spr_w, sh = it.spr_h;
    const int Wp = (sw + 15) & ~15, q = Wp >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= q * sh) return;
    const int y = idx / q, x = (idx - y * q) * 4;
    const uint8_t* row = sprites + (int64_t)it.frame * stride + (int64_t)y * pitch;
    uint32_t v[4] = {0u, 0u, 0u, 0u};
    if (x + 3 < sw && ((((uintptr_t)row) & 15) == 0)) {
        const uint4 t = ldg_stream16(row + 4 * x);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (x + k < sw) v[k] = __ldg(reinterpret_cast<const uint32_t*>(row) + x + k);
    }
    uint32_t pl[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t a = v[k] >> 24;
        pl[0] |= div255((v[k] & 255u) * a) << (8 * k);
        pl[1] |= div255(((v[k] >> 8) & 255u) * a) << (8 * k);
        pl[2] |= div255(((v[k] >> 16) & 255u) * a) << (8 * k);
        pl[3] |= a << (8 * k);
    }
    uint8_t* d = ws + it.rot_off + (int64_t)y * Wp + x;
    const int64_t plane = (int64_t)sh * Wp;
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint32_t*>(d + c * plane) = pl[c];
}

}  // namespace ipp

extern "C" {

int ipp_rotate_batch_u8(const uint8_t* src, int64_t src_pitch, int64_t src_image_stride, int src_w, int src_h,
                        int src_channels, int n_src, const ipp_rot_item* items, int n_items, int total_tiles, int mode,
                        uint8_t* dst_pool, int32_t* bbox_out, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    IPP_REQUIRE(src && items && dst_pool && bbox_out, IPP_ERR_BAD_ARG, "ipp_rotate_batch_u8: NULL pointer");
    IPP_REQUIRE((src_channels == 3 || src_channels == 4) && (mode == 0 || mode == 1) && n_src >= 1 && n_items >= 0 && total_tiles >= 0,
                IPP_ERR_BAD_ARG, "ipp_rotate_batch_u8: channels=%d mode=%d", src_channels, mode);
    IPP_REQUIRE(src_w > 0 && src_h > 0 && src_pitch >= (int64_t)src_w * src_channels, IPP_ERR_BAD_SHAPE, "ipp_rotate_batch_u8: bad geometry");
    IPP_REQUIRE(src_channels == 3 || (((src_pitch | src_image_stride) & 3) == 0 && (((uintptr_t)src) & 3) == 0), IPP_ERR_BAD_ARG,
                "ipp_rotate_batch_u8: RGBA sources must be 4-byte aligned");
    IPP_REQUIRE((((uintptr_t)dst_pool) & 255) == 0, IPP_ERR_BAD_ARG, "ipp_rotate_batch_u8: dst_pool must be 256-byte aligned");
    if (n_items == 0) return IPP_OK;
    { ipp::LaunchTimer _lt("bbox_reset_kernel", st); bbox_reset_kernel<<<ceil_div(n_items, 256), 256, 0, st>>>(bbox_out, n_items); }
    IPP_LAUNCH_CHECK();
    if (total_tiles == 0) return IPP_OK;
    RotBatchParams p{src, src_pitch, src_image_stride, src_w, src_h, items, n_items, dst_pool, bbox_out};
    {
        ipp::LaunchTimer _lt(mode == 0 ? "rotate_batch_nearest_kernel" : "rotate_batch_bilinear_kernel", st);
        if (mode == 0 && src_channels == 3) rotate_batch_kernel<0, 3><<<total_tiles, 256, 0, st>>>(p);
        else if (mode == 0) rotate_batch_kernel<0, 4><<<total_tiles, 256, 0, st>>>(p);
        else if (src_channels == 3) rotate_batch_kernel<1, 3><<<total_tiles, 256, 0, st>>>(p);
        else rotate_batch_kernel<1, 4><<<total_tiles, 256, 0, st>>>(p);
    }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

int ipp_sprite_planes_u8(const uint8_t* sprites, int64_t sprite_pitch, int64_t sprite_stride, int spr_w, int spr_h,
                         int n_sprites, const ipp_chain_item* items, int n_items, uint8_t* sprite_ws, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    IPP_REQUIRE(sprites && items && sprite_ws && n_sprites >= 1 && n_items >= 0, IPP_ERR_BAD_ARG, "ipp_sprite_planes_u8: bad arguments");
    IPP_REQUIRE(spr_w > 0 && spr_h > 0 && sprite_pitch >= (int64_t)spr_w * 4 && ((sprite_pitch | sprite_stride) & 3) == 0 &&
                    (((uintptr_t)sprites) & 3) == 0 && (((uintptr_t)sprite_ws) & 255) == 0,
                IPP_ERR_BAD_SHAPE, "ipp_sprite_planes_u8: sprites must be 4-byte aligned RGBA, the workspace 256-byte aligned");
    if (n_items == 0) return IPP_OK;
    const int q = ((spr_w + 15) & ~15) >> 2;
    dim3 grid(ceil_div((int64_t)q * spr_h, 256), n_items);
    { ipp::LaunchTimer _lt("sprite_planes_kernel", st); sprite_planes_kernel<<<grid, 256, 0, st>>>(sprites, sprite_pitch, sprite_stride, items, sprite_ws); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

int ipp_sprites_export_u8(const ipp_chain_item* items_host, int n_items, const uint8_t* sprite_ws,
                          uint8_t* dst, const int64_t* dst_offsets_host, void* stream) {
    using namespace ipp;
    IPP_REQUIRE((items_host && dst_offsets_host) || n_items == 0, IPP_ERR_BAD_ARG, "ipp_sprites_export_u8: NULL pointer");
    IPP_REQUIRE(sprite_ws && dst && n_items >= 0, IPP_ERR_BAD_ARG, "ipp_sprites_export_u8: bad arguments");
    for (int i = 0; i < n_items; ++i) {
        const ipp_chain_item& it = items_host[i];
        IPP_CUDA(cudaMemcpyAsync(dst + dst_offsets_host[i], sprite_ws + it.spr_off, (size_t)it.new_w * it.new_h * 4,
                                 cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return IPP_OK;
}

}  // extern "C"

// ===================================================================== HSV key + largest component, fused (BASELINE configs[3])
// filtres_liste.py:90-134 followed by pixels_isolés.py:29-58 without the BGRA image in between: stage A of the chain
// (ipp_chain_stage_a: frame read once -> 1-bit mask -> keep-mask of the largest component + its bbox) and this kernel, which
// writes the BGRA result: B, G, R of the (border-cropped) frame, alpha = 255 on the kept component, 0 elsewhere.
// (The key's alpha is 0 or 255; pixels_isolés zeroes every pixel outside the largest component, and an image without any
// foreground keeps its all-zero alpha.)  thread = 4 pixels: 12 source bytes, 4 mask bits, one 16-byte store.
namespace ipp {
#ifndef IPP_BGRA_QUADS
#define IPP_BGRA_QUADS 2
#endif
constexpr int kBgraQuads = IPP_BGRA_QUADS;     // 4-pixel groups per thread: all their loads are in flight before the first store
// grid = (groups of one image / (256 * kBgraQuads), images): 32-bit index arithmetic (a 64-bit division per thread made the
// one-group-per-thread form of this kernel issue-bound at 0.64 of the copy peak)
__global__ void __launch_bounds__(256) bgra_from_keep_kernel(const uint8_t* __restrict__ frames, int64_t pitch, int64_t stride,
                                                             int crop_l, int crop_t, int w, int h,
                                                             const uint32_t* __restrict__ bits, int64_t bits_stride, int wpr,
                                                             uint8_t* __restrict__ out, int64_t out_pitch, int64_t out_stride, int img0) {
    const uint32_t q = (uint32_t)(w + 3) >> 2, per = q * (uint32_t)h;
    const int img = img0 + (int)blockIdx.y;
    const uint32_t rem0 = blockIdx.x * (256u * kBgraQuads) + threadIdx.x;
    const uint8_t* fbase = frames + (int64_t)img * stride + (int64_t)crop_t * pitch + (int64_t)crop_l * 3;
    uint8_t* obase = out + (int64_t)img * out_stride;
    const uint32_t* bbase = bits + (int64_t)img * bits_stride;
    const uint8_t* sp[kBgraQuads];
    uint8_t* dp[kBgraQuads];
    uint32_t a[kBgraQuads], b[kBgraQuads], c[kBgraQuads], m[kBgraQuads];
    int left[kBgraQuads];                      // pixels of the group inside the row; 0 = no group; < 0 = slow path (unaligned / ragged)
#pragma unroll
    for (int k = 0; k < kBgraQuads; ++k) {
        const uint32_t rem = rem0 + 256u * k;
        left[k] = 0;
        if (rem >= per) continue;
        const uint32_t y = rem / q, x = (rem - y * q) * 4u;
        const uint8_t* s = fbase + (int64_t)y * pitch + x * 3u;
        uint8_t* d = obase + (int64_t)y * out_pitch + x * 4u;
        sp[k] = s; dp[k] = d;
        m[k] = __ldg(bbase + y * (uint32_t)wpr + (x >> 5)) >> (x & 31u);      // x is a multiple of 4: the four bits sit in one word
        left[k] = min(4, w - (int)x);
        if (left[k] == 4 && ((((uintptr_t)s) & 3) == 0) && ((((uintptr_t)d) & 15) == 0)) {
            // three loads of one lane sit 12 bytes apart: the requests of a warp touch the same sectors, the second and third hit L1
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s);
            a[k] = __ldg(s32); b[k] = __ldg(s32 + 1); c[k] = __ldg(s32 + 2);             // b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3
        } else {
            left[k] = -left[k];
        }
    }
#pragma unroll
    for (int k = 0; k < kBgraQuads; ++k) {
        if (left[k] == 4) {
            uint4 v;
            v.x = (a[k] & 0x00ffffffu) | ((m[k] & 1u) ? 0xff000000u : 0u);
            v.y = __funnelshift_r(a[k], b[k], 24) & 0x00ffffffu; v.y |= (m[k] & 2u) ? 0xff000000u : 0u;
            v.z = __funnelshift_r(b[k], c[k], 16) & 0x00ffffffu; v.z |= (m[k] & 4u) ? 0xff000000u : 0u;
            v.w = (c[k] >> 8) | ((m[k] & 8u) ? 0xff000000u : 0u);
            stg_stream16(dp[k], v);
        } else if (left[k] < 0) {
            const uint8_t* s = sp[k];
            uint8_t* d = dp[k];
            for (int j = 0; j < -left[k]; ++j) {
                d[4 * j + 0] = s[3 * j + 0]; d[4 * j + 1] = s[3 * j + 1]; d[4 * j + 2] = s[3 * j + 2];
                d[4 * j + 3] = ((m[k] >> j) & 1u) ? 255 : 0;
            }
        }
    }
}
}  // namespace ipp

extern "C" int ipp_bgra_from_keep_u8(const uint8_t* frames, int64_t frame_pitch, int64_t frame_stride, const ipp_chain_geom* g,
                                     const void* workspace, uint8_t* out, int64_t out_pitch, int64_t out_stride, int n, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = chain_check_geom(g);
    if (rc != IPP_OK) return rc;
    IPP_REQUIRE(frames && workspace && out && n >= 1, IPP_ERR_BAD_ARG, "ipp_bgra_from_keep_u8: bad arguments");
    IPP_REQUIRE(frame_pitch >= (int64_t)g->frame_w * 3 && out_pitch >= (int64_t)g->crop_w * 4, IPP_ERR_BAD_SHAPE, "ipp_bgra_from_keep_u8: pitch smaller than a row");
    const CclGeom cg = make_ccl_geom(g->crop_w, g->crop_h);
    const int64_t per = (int64_t)((g->crop_w + 3) >> 2) * g->crop_h;
    IPP_REQUIRE(per < (1LL << 31), IPP_ERR_UNSUPPORTED, "ipp_bgra_from_keep_u8: image too large");
    for (int img0 = 0; img0 < n; img0 += 65535) {
        const dim3 grid((unsigned)ceil_div(per, 256 * kBgraQuads), (unsigned)std::min(65535, n - img0));
        ipp::LaunchTimer _lt("bgra_from_keep_kernel", st);
        bgra_from_keep_kernel<<<grid, 256, 0, st>>>(frames, frame_pitch, frame_stride, g->crop_l, g->crop_t, g->crop_w, g->crop_h,
                                                    (const uint32_t*)workspace, cg.bits_stride, cg.wpr, out, out_pitch, out_stride, img0);
    }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}
