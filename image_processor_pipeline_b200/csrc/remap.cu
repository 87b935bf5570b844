// remap.cu — flip x crop x copy as a load-address remap; fused symmetries; to-RGBA.
//
// Reference behaviour replaced: cv2.flip / ndarray.copy (transforms/symmetry.py:114-119),
// ndarray slicing crops (recadrages.py:46, crop_square.py:196, pixels_isolés.py:81),
// PIL crop (rotations.py:101) and convert("RGBA") (rotations.py:55).
//
// One CTA owns a few consecutive source rows.  The bulk-copy engine stages the
// 16-byte-aligned window of each row that covers [x0*bpp, (x0+out_w)*bpp) into shared
// memory (one read of the source, whatever its alignment); the CTA then emits up to
// four variants from that single staging: o (as is), v (mirrored row index),
// h and hv (pixel order reversed inside the row).  Destination rows are written as
// aligned 16-byte vectors (head/tail bytes individually), so any pitch works.
#include "staging.cuh"

namespace ipp {

struct RowsParams {
    const uint8_t* src;
    int64_t src_pitch, src_img_stride;
    int bpp, x0, y0, out_w, out_h;
    uint8_t* dst[4];  // 0 = o, 1 = h, 2 = v, 3 = hv
    int64_t dst_pitch, dst_img_stride;
    int rows_per_cta;
    uint32_t smem_row_stride;
    const int32_t* xy;  // optional per-image {x0, y0} (device); nullptr: x0 / y0 above for every image
};

constexpr int kRowsThreads = 256;

// 16 output bytes of a horizontally mirrored row, starting at output byte ob0.
__device__ __forceinline__ uint4 gather_reversed16(const uint8_t* win, uint32_t shift, uint32_t ob0,
                                                   int out_w, int bpp) {
    uint32_t q = ob0 / (uint32_t)bpp;
    uint32_t rr = ob0 - q * (uint32_t)bpp;
    int sp = (int)shift + (out_w - 1 - (int)q) * bpp + (int)rr;
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc |= (uint32_t)win[sp] << (8 * k);
            ++rr;
            ++sp;
            if (rr == (uint32_t)bpp) {
                rr = 0;
                sp -= 2 * bpp;
            }
        }
        w[i] = acc;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// 3-byte pixels, mirrored: 16 pixels = 48 bytes at a time.  The 48 output bytes starting at pixel 16 u of the mirrored row are the
// 16 source pixels that END at pixel out_w - 1 - 16 u, in reverse pixel order: three unaligned 16-byte window loads, then the
// pixels of every 12-byte group swap places with four byte permutes (b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3 ->
// b3 g3 r3 b2 | g2 r2 b1 g1 | r1 b0 g0 r0) and the four groups leave in reverse order — ~50 instructions per 48 bytes where the
// byte-wise gather needs ~330 (the mirrored variants made the kernel issue-bound).
__device__ __forceinline__ void reverse_group3(uint32_t x, uint32_t y, uint32_t z, uint32_t& o0, uint32_t& o1, uint32_t& o2) {
    o0 = __byte_perm(y, z, 0x2765);
    o1 = __byte_perm(__byte_perm(y, z, 0x0043), x, 0x3710);
    o2 = __byte_perm(x, y, 0x2105);
}
__device__ __forceinline__ void store_reversed48(const uint8_t* win, uint32_t src_off, uint8_t* d) {
    const uint4 a = window_load16(win, src_off), b = window_load16(win, src_off + 16), c = window_load16(win, src_off + 32);
    // source groups g0 = a.xyz, g1 = a.w b.xy, g2 = b.zw c.x, g3 = c.yzw; output = rev(g3) rev(g2) rev(g1) rev(g0)
    uint4 o0, o1, o2;
    reverse_group3(c.y, c.z, c.w, o0.x, o0.y, o0.z);
    reverse_group3(b.z, b.w, c.x, o0.w, o1.x, o1.y);
    reverse_group3(a.w, b.x, b.y, o1.z, o1.w, o2.x);
    reverse_group3(a.x, a.y, a.z, o2.y, o2.z, o2.w);
    stg_stream16(d, o0);
    stg_stream16(d + 16, o1);
    stg_stream16(d + 32, o2);
}

__global__ void __launch_bounds__(kRowsThreads) rows_kernel(const RowsParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    const int img = blockIdx.y;
    const int row0 = blockIdx.x * p.rows_per_cta;
    const int nrows = min(p.rows_per_cta, p.out_h - row0);
    const uint32_t rowbytes = (uint32_t)p.out_w * (uint32_t)p.bpp;
    const uint8_t* src_img = p.src + (int64_t)img * p.src_img_stride;
    // per-image offsets are device data: clamped to [0, src - out] (p.x0 / p.y0 then hold those maxima), never read outside the image
    const int x0 = p.xy ? min(max(p.xy[2 * img], 0), p.x0) : p.x0, y0 = p.xy ? min(max(p.xy[2 * img + 1], 0), p.y0) : p.y0;

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        uint32_t total = 0;
        for (int r = 0; r < nrows; ++r) {
            const uint8_t* s = src_img + (int64_t)(y0 + row0 + r) * p.src_pitch + (int64_t)x0 * p.bpp;
            total += make_window(s, rowbytes).bytes;
        }
        mbar_expect_tx(&bar, total);
        for (int r = 0; r < nrows; ++r) {
            const uint8_t* s = src_img + (int64_t)(y0 + row0 + r) * p.src_pitch + (int64_t)x0 * p.bpp;
            const Window w = make_window(s, rowbytes);
            bulk_g2s(smem + (size_t)r * p.smem_row_stride, w.base, w.bytes, &bar);
        }
    }
    mbar_wait(&bar, 0);

    const uint32_t cpr = (rowbytes >> 4) + 2;  // body chunks (+1 slack), head, tail
    const uint32_t items = (uint32_t)nrows * cpr;
#pragma unroll 1
    for (int v = 0; v < 4; ++v) {
        uint8_t* dimg = p.dst[v];
        if (dimg == nullptr) continue;
        dimg += (int64_t)img * p.dst_img_stride;
        const bool rev = (v & 1) != 0;
        // mirrored 3-byte pixels with 16-byte aligned destination rows: 48-byte units (16 pixels); the generic loop below then
        // only sees the bytes behind the last whole unit
        const bool fast_rev = rev && p.bpp == 3 && ((((uintptr_t)dimg) | (uintptr_t)p.dst_pitch) & 15) == 0;
        const uint32_t upr = fast_rev ? rowbytes / 48u : 0u;     // units per row
        if (fast_rev) {
            for (uint32_t it = tid; it < (uint32_t)nrows * upr; it += kRowsThreads) {
                const uint32_t r = it / upr, u = it - r * upr;
                const int y = row0 + (int)r;
                const int drow = (v & 2) ? (p.out_h - 1 - y) : y;
                const uint8_t* s = src_img + (int64_t)(y0 + y) * p.src_pitch + (int64_t)x0 * 3;
                const uint32_t shift = (uint32_t)((uintptr_t)s & 15);
                store_reversed48(smem + (size_t)r * p.smem_row_stride, shift + (uint32_t)(p.out_w - 16 - 16 * (int)u) * 3u,
                                 dimg + (int64_t)drow * p.dst_pitch + 48u * u);
            }
        }
        const uint32_t c_first = upr * 3u;                       // first 16-byte chunk the generic loop still owns (head == 0 here)
        for (uint32_t it = tid; it < items; it += kRowsThreads) {
            const uint32_t r = it / cpr;
            const uint32_t c = it - r * cpr;
            const int y = row0 + (int)r;
            const int drow = (v & 2) ? (p.out_h - 1 - y) : y;
            uint8_t* d = dimg + (int64_t)drow * p.dst_pitch;
            const uint8_t* win = smem + (size_t)r * p.smem_row_stride;
            const uint8_t* s = src_img + (int64_t)(y0 + y) * p.src_pitch + (int64_t)x0 * p.bpp;
            const uint32_t shift = (uint32_t)((uintptr_t)s & 15);
            const RowSplit sp = split_row(d, rowbytes);
            if (c < sp.nb) {
                if (c < c_first) continue;
                const uint32_t ob = sp.head + (c << 4);
                uint4 val;
                if (!rev) {
                    val = window_load16(win, shift + ob);
                } else if (p.bpp == 4 && ((shift | ob) & 3u) == 0) {
                    const uint32_t q = ob >> 2;
                    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(win + shift);
                    const int last = p.out_w - 1 - (int)q;
                    val = make_uint4(w32[last], w32[last - 1], w32[last - 2], w32[last - 3]);
                } else {
                    val = gather_reversed16(win, shift, ob, p.out_w, p.bpp);
                }
                stg_stream16(d + ob, val);
            } else if (c == cpr - 2 || c == cpr - 1) {
                // head bytes (c == cpr-2) or tail bytes (c == cpr-1)
                const uint32_t b0 = (c == cpr - 2) ? 0u : sp.head + (sp.nb << 4);
                const uint32_t cnt = (c == cpr - 2) ? sp.head : sp.tail;
                for (uint32_t k = 0; k < cnt; ++k) {
                    const uint32_t ob = b0 + k;
                    uint32_t sb;
                    if (!rev) {
                        sb = shift + ob;
                    } else {
                        const uint32_t q = ob / (uint32_t)p.bpp;
                        const uint32_t rr = ob - q * (uint32_t)p.bpp;
                        sb = shift + (uint32_t)(p.out_w - 1 - (int)q) * (uint32_t)p.bpp + rr;
                    }
                    d[ob] = win[sb];
                }
            }
        }
    }
}

static int launch_rows(RowsParams& p, int n, cudaStream_t st) {
    const uint32_t rowbytes = (uint32_t)p.out_w * (uint32_t)p.bpp;
    p.smem_row_stride = staged_row_stride(rowbytes);
    IPP_REQUIRE(p.smem_row_stride <= 200 * 1024, IPP_ERR_UNSUPPORTED,
                "remap: a row of %u bytes does not fit the shared-memory staging buffer", rowbytes);
    // ~48 KB of rows per CTA (4 resident CTAs: ~190 KB in flight per SM).  Measured on C1 (1024 x 1024 x 3, one read + four
    // writes): 16 / 24 / 32 / 48 / 64 KB = 0.74 / 0.81 / 0.82 / 0.86 / 0.83 of the copy peak — a CTA loads, waits and then writes,
    // so the bytes in flight come from the other resident CTAs, and at 16 KB the register file allowed only 6 of them
    static const int budget_env = getenv("IPP_ROWS_SMEM") ? atoi(getenv("IPP_ROWS_SMEM")) : 0;
    int rows = (int)((budget_env > 0 ? budget_env : 49152) / p.smem_row_stride);
    if (rows < 1) rows = 1;
    if (rows > p.out_h) rows = p.out_h;
    // keep the grid at >= ~4 waves of CTAs when the image is small
    const int sms = sm_count();
    while (rows > 1 && (int64_t)ceil_div(p.out_h, rows) * n < 4LL * sms) rows = (rows + 1) / 2;
    p.rows_per_cta = rows;
    const size_t smem = (size_t)rows * p.smem_row_stride;
    if (smem > 48 * 1024) {
        IPP_CUDA(cudaFuncSetAttribute(rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(ceil_div(p.out_h, rows), n);
    { ipp::LaunchTimer _lt("rows_kernel", st); rows_kernel<<<grid, kRowsThreads, smem, st>>>(p); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

// ------------------------------------------------------------------ to RGBA
__global__ void __launch_bounds__(256) to_rgba_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, int w, int h,
                                                      int ch, int swap_rb, uint8_t* __restrict__ dst, int64_t dst_pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* s = src + (int64_t)y * src_pitch + (int64_t)x * ch;
    uint32_t c0, c1, c2, a = 255;
    if (ch == 1) {
        c0 = c1 = c2 = s[0];
    } else {
        c0 = s[0];
        c1 = s[1];
        c2 = s[2];
        if (ch == 4) a = s[3];
    }
    if (swap_rb) {
        const uint32_t t = c0;
        c0 = c2;
        c2 = t;
    }
    const uint32_t px = c0 | (c1 << 8) | (c2 << 16) | (a << 24);
    uint8_t* d = dst + (int64_t)y * dst_pitch + (int64_t)x * 4;
    if ((((uintptr_t)d) & 3) == 0) {
        *reinterpret_cast<uint32_t*>(d) = px;
    } else {
        d[0] = (uint8_t)c0; d[1] = (uint8_t)c1; d[2] = (uint8_t)c2; d[3] = (uint8_t)a;
    }
}

// ------------------------------------------------------------------ quarter turns (PIL transpose fast paths)
// ROTATE_90 (counter-clockwise): out[y][x] = in[x][w-1-y];  ROTATE_270: out[y][x] = in[h-1-x][y].
// 32x32 tiles through shared memory so that both the loads and the stores are row-contiguous.
__global__ void __launch_bounds__(256) rot90_rgba_kernel(const uint8_t* __restrict__ src, int64_t src_pitch, int w, int h,
                                                         uint8_t* __restrict__ dst, int64_t dst_pitch, int ccw) {
    __shared__ uint32_t tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // destination tile origin (dst is h wide, w high)
    const int dx0 = blockIdx.x * 32, dy0 = blockIdx.y * 32;
    // source tile origin
    const int sx0 = ccw ? (w - 32 - dy0) : dy0;
    const int sy0 = ccw ? dx0 : (h - 32 - dx0);
    for (int r = ty; r < 32; r += 8) {
        const int sx = sx0 + tx, sy = sy0 + r;
        if (sx >= 0 && sx < w && sy >= 0 && sy < h)
            tile[r][tx] = *reinterpret_cast<const uint32_t*>(src + (int64_t)sy * src_pitch + (int64_t)sx * 4);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int dx = dx0 + tx, dy = dy0 + r;
        if (dx < h && dy < w) {
            // ccw: in[x][w-1-y] -> source row = dx - sy0 ... expressed in tile coordinates
            const int srow = ccw ? (dx - sy0) : (h - 1 - dx - sy0);
            const int scol = ccw ? (w - 1 - dy - sx0) : (dy - sx0);
            *reinterpret_cast<uint32_t*>(dst + (int64_t)dy * dst_pitch + (int64_t)dx * 4) = tile[srow][scol];
        }
    }
}

}  // namespace ipp

extern "C" {

int ipp_rot90_rgba_u8(const uint8_t* src, int64_t src_pitch, int w, int h, int counter_clockwise,
                      uint8_t* dst, int64_t dst_pitch, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(src && dst, IPP_ERR_BAD_ARG, "ipp_rot90_rgba_u8: NULL image pointer");
    IPP_REQUIRE(w > 0 && h > 0 && src_pitch >= (int64_t)w * 4 && dst_pitch >= (int64_t)h * 4, IPP_ERR_BAD_SHAPE,
                "ipp_rot90_rgba_u8: bad geometry");
    IPP_REQUIRE(((src_pitch | dst_pitch) & 3) == 0 && ((((uintptr_t)src) | ((uintptr_t)dst)) & 3) == 0, IPP_ERR_BAD_ARG,
                "ipp_rot90_rgba_u8: RGBA images must be 4-byte aligned");
    dim3 grid(ceil_div(h, 32), ceil_div(w, 32));
    { ipp::LaunchTimer _lt("rot90_rgba_kernel", (cudaStream_t)stream); rot90_rgba_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_pitch, w, h, dst, dst_pitch, counter_clockwise); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

int ipp_remap_u8(const uint8_t* src, int64_t src_pitch, int src_w, int src_h, int bpp,
                 int x0, int y0, int out_w, int out_h, int flip_h, int flip_v,
                 uint8_t* dst, int64_t dst_pitch,
                 int n, int64_t src_image_stride, int64_t dst_image_stride, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(src && dst, IPP_ERR_BAD_ARG, "ipp_remap_u8: NULL image pointer");
    IPP_REQUIRE(bpp >= 1 && bpp <= 16, IPP_ERR_BAD_ARG, "ipp_remap_u8: bpp=%d not in 1..16", bpp);
    IPP_REQUIRE(n >= 1, IPP_ERR_BAD_ARG, "ipp_remap_u8: n=%d", n);
    IPP_REQUIRE(out_w >= 0 && out_h >= 0 && x0 >= 0 && y0 >= 0 && x0 + out_w <= src_w && y0 + out_h <= src_h,
                IPP_ERR_BAD_SHAPE, "ipp_remap_u8: crop (%d,%d,%d,%d) outside %dx%d", x0, y0, out_w, out_h, src_w, src_h);
    IPP_REQUIRE(src_pitch >= (int64_t)src_w * bpp && dst_pitch >= (int64_t)out_w * bpp, IPP_ERR_BAD_SHAPE,
                "ipp_remap_u8: pitch smaller than a row");
    if (out_w == 0 || out_h == 0) return IPP_OK;
    RowsParams p{};
    p.src = src; p.src_pitch = src_pitch; p.src_img_stride = src_image_stride;
    p.bpp = bpp; p.x0 = x0; p.y0 = y0; p.out_w = out_w; p.out_h = out_h;
    p.dst[(flip_h ? 1 : 0) | (flip_v ? 2 : 0)] = dst;
    p.dst_pitch = dst_pitch; p.dst_img_stride = dst_image_stride;
    return launch_rows(p, n, (cudaStream_t)stream);
}

int ipp_crop_batch_u8(const uint8_t* src, int64_t src_pitch, int64_t src_image_stride, int src_w, int src_h, int bpp,
                      const int32_t* xy_offsets, int out_w, int out_h,
                      uint8_t* dst, int64_t dst_pitch, int64_t dst_image_stride, int n, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(src && dst && xy_offsets, IPP_ERR_BAD_ARG, "ipp_crop_batch_u8: NULL pointer");
    IPP_REQUIRE(bpp >= 1 && bpp <= 16 && n >= 1, IPP_ERR_BAD_ARG, "ipp_crop_batch_u8: bpp=%d n=%d", bpp, n);
    IPP_REQUIRE(out_w >= 0 && out_h >= 0 && out_w <= src_w && out_h <= src_h, IPP_ERR_BAD_SHAPE,
                "ipp_crop_batch_u8: crop %dx%d larger than %dx%d", out_w, out_h, src_w, src_h);
    IPP_REQUIRE(src_pitch >= (int64_t)src_w * bpp && dst_pitch >= (int64_t)out_w * bpp, IPP_ERR_BAD_SHAPE,
                "ipp_crop_batch_u8: pitch smaller than a row");
    if (out_w == 0 || out_h == 0) return IPP_OK;
    RowsParams p{};
    p.src = src; p.src_pitch = src_pitch; p.src_img_stride = src_image_stride;
    p.bpp = bpp; p.out_w = out_w; p.out_h = out_h;
    p.dst[0] = dst;
    p.dst_pitch = dst_pitch; p.dst_img_stride = dst_image_stride;
    p.xy = xy_offsets;
    p.x0 = src_w - out_w; p.y0 = src_h - out_h;      // clamp limits of the per-image offsets
    return launch_rows(p, n, (cudaStream_t)stream);
}

int ipp_symmetries_u8(const uint8_t* src, int64_t src_pitch, int w, int h, int bpp,
                      uint8_t* dst_o, uint8_t* dst_h, uint8_t* dst_v, uint8_t* dst_hv,
                      int64_t dst_pitch,
                      int n, int64_t src_image_stride, int64_t dst_image_stride, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(src, IPP_ERR_BAD_ARG, "ipp_symmetries_u8: NULL source");
    IPP_REQUIRE(bpp >= 1 && bpp <= 16, IPP_ERR_BAD_ARG, "ipp_symmetries_u8: bpp=%d not in 1..16", bpp);
    IPP_REQUIRE(n >= 1 && w >= 0 && h >= 0, IPP_ERR_BAD_ARG, "ipp_symmetries_u8: bad n/size");
    IPP_REQUIRE(src_pitch >= (int64_t)w * bpp && dst_pitch >= (int64_t)w * bpp, IPP_ERR_BAD_SHAPE,
                "ipp_symmetries_u8: pitch smaller than a row");
    if (w == 0 || h == 0 || (!dst_o && !dst_h && !dst_v && !dst_hv)) return IPP_OK;
    RowsParams p{};
    p.src = src; p.src_pitch = src_pitch; p.src_img_stride = src_image_stride;
    p.bpp = bpp; p.x0 = 0; p.y0 = 0; p.out_w = w; p.out_h = h;
    p.dst[0] = dst_o; p.dst[1] = dst_h; p.dst[2] = dst_v; p.dst[3] = dst_hv;
    p.dst_pitch = dst_pitch; p.dst_img_stride = dst_image_stride;
    return launch_rows(p, n, (cudaStream_t)stream);
}

int ipp_to_rgba_u8(const uint8_t* src, int64_t src_pitch, int w, int h, int src_channels,
                   int swap_rb, uint8_t* dst, int64_t dst_pitch, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(src && dst, IPP_ERR_BAD_ARG, "ipp_to_rgba_u8: NULL image pointer");
    IPP_REQUIRE(src_channels == 1 || src_channels == 3 || src_channels == 4, IPP_ERR_BAD_ARG,
                "ipp_to_rgba_u8: channels=%d not in {1,3,4}", src_channels);
    IPP_REQUIRE(w >= 0 && h >= 0 && src_pitch >= (int64_t)w * src_channels && dst_pitch >= (int64_t)w * 4,
                IPP_ERR_BAD_SHAPE, "ipp_to_rgba_u8: bad size/pitch");
    if (w == 0 || h == 0) return IPP_OK;
    dim3 grid(ceil_div(w, 256), h);
    { ipp::LaunchTimer _lt("to_rgba_kernel", (cudaStream_t)stream); to_rgba_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_pitch, w, h, src_channels, swap_rb, dst, dst_pitch); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

}  // extern "C"
