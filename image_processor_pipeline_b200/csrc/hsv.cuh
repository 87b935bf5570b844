// hsv.cuh — OpenCV-exact 8-bit BGR->HSV (H in [0,180)) and the inRange test of a
// filter list, shared by the stand-alone keying kernel and the fused chain.
//
// Reference behaviour: cv2.cvtColor(BGR2HSV) + cv2.inRange per filter + zone mask
// (transforms/filtres_liste.py:90-117).  OpenCV RGB2HSV_b: v=max, d=v-min,
// s=(d*sdiv[v]+2048)>>12, h from the max channel, h=(h*hdiv180[d]+2048)>>12, +180 if <0,
// sdiv[i]=cvRound(255*4096/i), hdiv180[i]=cvRound(180*4096/(6*i)).
#pragma once
#include "common.cuh"

namespace ipp {

// inRange of one HSV triple against one filter, all three bands at once: bands sit
// in 10-bit fields (8 value bits + 2 guard bits) so that
//   field(X + clo) has bit 9 set  <=>  x >= lo      (clo field = 512 - lo)
//   field(chi - X) has bit 9 set  <=>  x <= hi      (chi field = 512 + hi)
// with no carry/borrow between fields for lo in [0,256], hi in [-1,255].
constexpr uint32_t kGuard = (1u << 9) | (1u << 19) | (1u << 29);

struct HsvFilterSet {
    int n;
    uint32_t clo[IPP_MAX_HSV_RANGES];
    uint32_t chi[IPP_MAX_HSV_RANGES];
    int zx0[IPP_MAX_HSV_RANGES], zx1[IPP_MAX_HSV_RANGES];   // zone rect [x0,x1) x [y0,y1)
    int zy0[IPP_MAX_HSV_RANGES], zy1[IPP_MAX_HSV_RANGES];
    int all_zones_full;
};

// host: pack the caller's ranges for an image of w x h (zone margins resolved like
// numpy resolves mask[top : h-bottom, left : w-right], negative stops included).
static inline void slice_indices(int start, int stop, int len, int* a, int* b) {
    if (start < 0) { start += len; if (start < 0) start = 0; } else if (start > len) start = len;
    if (stop < 0) { stop += len; if (stop < 0) stop = 0; } else if (stop > len) stop = len;
    if (stop < start) stop = start;
    *a = start; *b = stop;
}

static inline HsvFilterSet pack_filters(const ipp_hsv_range* r, int n, int w, int h) {
    HsvFilterSet f{};
    f.n = n;
    f.all_zones_full = 1;
    for (int k = 0; k < n; ++k) {
        uint32_t clo = 0, chi = 0;
        for (int b = 0; b < 3; ++b) {
            int lo = r[k].lo[b], hi = r[k].hi[b];
            if (lo < 0) lo = 0;
            if (lo > 256) lo = 256;
            if (hi > 255) hi = 255;
            if (hi < -1) hi = -1;
            clo |= (uint32_t)(512 - lo) << (10 * b);
            chi |= (uint32_t)(512 + hi) << (10 * b);
        }
        f.clo[k] = clo;
        f.chi[k] = chi;
        slice_indices(r[k].zone[0], h - r[k].zone[1], h, &f.zy0[k], &f.zy1[k]);
        slice_indices(r[k].zone[2], w - r[k].zone[3], w, &f.zx0[k], &f.zx1[k]);
        if (!(f.zx0[k] == 0 && f.zy0[k] == 0 && f.zx1[k] == w && f.zy1[k] == h)) f.all_zones_full = 0;
    }
    return f;
}

// The two 256-entry division tables live in device global memory (filled once per
// device by ensure_hsv_tables(), host-computed with cvRound = round-half-to-even) and
// are copied into shared memory by every CTA (coalesced, L2 resident).
// (libippgpu is built as ONE translation unit — ippgpu_all.cu — so these are plain definitions.)
__device__ int g_hsv_sdiv[256];
__device__ int g_hsv_hdiv[256];
int ensure_hsv_tables();   // host (hsv.cu); returns 0 or a cudaError_t

__device__ __forceinline__ void fill_hsv_tables(int* sdiv, int* hdiv, int tid, int nthreads) {
    for (int i = tid; i < 256; i += nthreads) {
        sdiv[i] = g_hsv_sdiv[i];
        hdiv[i] = g_hsv_hdiv[i];
    }
}

// packed HSV of one pixel: h | s << 10 | v << 20
__device__ __forceinline__ uint32_t bgr_to_hsv_packed(int b, int g, int r, const int* sdiv, const int* hdiv) {
    const int v = max(max(b, g), r);
    const int mn = min(min(b, g), r);
    const int diff = v - mn;
    const int s = (diff * sdiv[v] + (1 << 11)) >> 12;
    int hn;
    if (v == r) hn = g - b;
    else if (v == g) hn = b - r + 2 * diff;
    else hn = r - g + 4 * diff;
    int h = (hn * hdiv[diff] + (1 << 11)) >> 12;
    if (h < 0) h += 180;
    return (uint32_t)h | ((uint32_t)s << 10) | ((uint32_t)v << 20);
}

// true when the pixel at (x, y) falls in any filter (and its zone): pixel is keyed OUT
__device__ __forceinline__ bool hsv_excluded(uint32_t X, int x, int y, const HsvFilterSet& f) {
    bool ex = false;
#pragma unroll 4
    for (int k = 0; k < f.n; ++k) {
        const bool in = (((X + f.clo[k]) & (f.chi[k] - X)) & kGuard) == kGuard;
        const bool zone = f.all_zones_full || (x >= f.zx0[k] && x < f.zx1[k] && y >= f.zy0[k] && y < f.zy1[k]);
        ex |= (in && zone);
    }
    return ex;
}

}  // namespace ipp
