// ccl.cuh — 8-connected component labelling on a 1-bit/pixel mask, reduced to what the
// reference needs: WHICH component is the largest, its bounding box, and a keep-mask.
//
// Reference behaviour replaced: cv2.connectedComponentsWithStats(binary, 8) + the
// argmax-area loop + clean_mask (transforms/pixels_isolés.py:35-55).  OpenCV numbers
// components in 2x2-block raster order; the reference keeps the FIRST label with the
// strictly greatest area, i.e. among equal areas the component whose smallest
// 2x2-block raster index is smallest.
//
// Design (bandwidth-lean: nothing per pixel except the bit itself):
//  * pixels -> 2x2 blocks (all pixels of a block are mutually 8-adjacent), the unit
//    OpenCV's own algorithm labels; block index = by*BW + bx IS the tie-break key.
//  * horizontal connectivity is resolved with bit arithmetic only: per block row a bit
//    vector C ("block is connected to its left neighbour"); a maximal chain of C bits is
//    a RUN and is represented by its first block (found with clz, no memory traffic).
//  * union-find (atomicMin, root = smallest block index) is done only between RUNS of
//    adjacent block rows, and only once per contact segment (bit-parallel dedup), so a
//    convex blob costs ~1 union per block row instead of one per pixel.
//  * areas are accumulated per run segment (popc) onto the root; argmax = one 64-bit
//    atomicMax of (area << 32 | ~root) per root.
//  * the keep-mask (pixels of the winning component) overwrites the pixel mask in place.
// Only run-start entries of the label/area arrays are ever touched.
#pragma once
#include "common.cuh"

namespace ipp {

// A warp of the group passes covers a PATCH of kCclPatchW x kCclPatchH groups (not 32 consecutive groups of one block row):
// a compact object then fills whole warps and the background leaves whole warps empty, where a row-wise mapping ran the
// long path of every pass at ~25 % lane efficiency.
#ifndef IPP_CCL_PATCH_W
#define IPP_CCL_PATCH_W 2
#endif
constexpr int kCclPatchW = IPP_CCL_PATCH_W, kCclPatchH = 32 / kCclPatchW;
struct CclGeom {
    int w, h;        // pixels
    int wpr;         // 32-bit words per pixel row
    int bw, bh;      // blocks
    int bww;         // 32-bit words per block row
    uint32_t bww_inv;    // floor(2^32 / bww) + 1: __umulhi(x, bww_inv) = x / bww or one more (x * bww < 2^32)
    int pw, npatch;      // warp patches (kCclPatchW x kCclPatchH groups) per band of block rows, and per image
    uint32_t pw_inv;     // floor(2^32 / pw) + 1
    // per-image strides (in elements of the respective arrays)
    int64_t bits_stride;   // uint32 words
    int64_t cf_stride;     // uint32 words (C and F have the same shape)
    int64_t lab_stride;    // int32 entries (L and A)
};

struct CclBuffers {
    uint32_t* bits;      // [n][h][wpr]       pixel mask in, keep-mask out
    uint32_t* cbits;     // [n][bh][bww]
    uint32_t* fbits;     // [n][bh][bww]
    int2* la;            // [n][bh*bw]  .x = union-find parent (label), .y = area (valid on roots);
                         //              interleaved: a run start costs one 32-byte sector, not two
    unsigned long long* best;   // [n]  (area << 32) | (0xffffffff - root)
    int32_t* acc;        // [n][8]  minx, miny, maxx1, maxy1, n_fg, -, -, -
};

static inline CclGeom make_ccl_geom(int w, int h) {
    CclGeom g;
    g.w = w; g.h = h;
    g.wpr = (w + 31) / 32;
    g.bw = (w + 1) / 2;
    g.bh = (h + 1) / 2;
    g.bww = (g.bw + 31) / 32;
    g.bww_inv = (uint32_t)((1ull << 32) / (uint64_t)g.bww) + 1u;
    g.pw = (g.bww + kCclPatchW - 1) / kCclPatchW;
    g.npatch = g.pw * ((g.bh + kCclPatchH - 1) / kCclPatchH);
    g.pw_inv = (uint32_t)((1ull << 32) / (uint64_t)g.pw) + 1u;
    g.bits_stride = (int64_t)h * g.wpr;
    g.cf_stride = (int64_t)g.bh * g.bww;
    g.lab_stride = (int64_t)g.bh * g.bw;
    return g;
}

static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

// bytes of workspace for n images (bits + C + F + label + area + best + acc)
static inline int64_t ccl_workspace_bytes(const CclGeom& g, int n) {
    int64_t t = 0;
    t += align256(g.bits_stride * 4 * n);
    t += 2 * align256(g.cf_stride * 4 * n);
    t += 2 * align256(g.lab_stride * 4 * n);
    t += align256((int64_t)8 * n);
    t += align256((int64_t)32 * n);
    return t;
}

static inline CclBuffers carve_ccl(void* ws, const CclGeom& g, int n) {
    CclBuffers b;
    uint8_t* p = (uint8_t*)ws;
    b.bits = (uint32_t*)p;   p += align256(g.bits_stride * 4 * n);
    b.cbits = (uint32_t*)p;  p += align256(g.cf_stride * 4 * n);
    b.fbits = (uint32_t*)p;  p += align256(g.cf_stride * 4 * n);
    b.la = (int2*)p;         p += 2 * align256(g.lab_stride * 4 * n);
    b.best = (unsigned long long*)p; p += align256((int64_t)8 * n);
    b.acc = (int32_t*)p;
    return b;
}

// host: run the CCL phases on an already filled pixel mask (stream-ordered)
int ccl_run(const CclGeom& g, const CclBuffers& b, int n, cudaStream_t st);

}  // namespace ipp
