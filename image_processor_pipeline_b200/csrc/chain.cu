// chain.cu — the fused, batched, device-resident transform chain (include/ippgpu.h §9).
//
//   stage A   frame -> (border crop folded into addressing) -> HSV key -> 1-bit mask
//             -> largest 8-connected component (ccl.cu) -> keep-mask + bbox/area
//   stage B1  per rotated sprite: bbox of the rotated keep-mask (PIL getbbox)
//   stage B2  gather: rotate∘fit∘flip folded into one inverse map frame -> premultiplied
//             sprite; LANCZOS horizontal pass; vertical pass + un-premultiply
//   stage C   background crop + alpha blend + final square crop in one streaming pass
//
// Per frame the 6.2 MB of pixels are read from HBM exactly once at full resolution
// (stage A); everything between A and C works on bits or on the small sprite.
#include "ccl.cuh"
#include "hsv.cuh"
#include "staging.cuh"

#include <algorithm>
#include <type_traits>
#include <cstdlib>

namespace ipp {

// ===================================================================== stage A: mask
struct MaskParams {
    const uint8_t* frames;
    int64_t pitch, stride;
    int crop_l, crop_t, w, h;     // w, h = cropped size
    int wpr;
    uint32_t* bits;
    int64_t bits_stride;          // words per image
    int rows_per_cta;
    uint32_t smem_row_stride;
    HsvFilterSet f;
};

constexpr int kMaskThreads = 256;

// One CTA = a few cropped rows of one frame, staged by the bulk-copy engine.  A lane
// converts 4 pixels (12 staged bytes, conflict-free 12-byte lane stride) into a nibble;
// 8 lanes are OR-combined by shuffles into one mask word.
__global__ void __launch_bounds__(kMaskThreads) chain_mask_kernel(const MaskParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ int sdiv[256], hdiv[256];
    const int tid = threadIdx.x;
    const int img = blockIdx.y;
    const int row0 = blockIdx.x * p.rows_per_cta;
    const int nrows = min(p.rows_per_cta, p.h - row0);
    const uint32_t rowbytes = (uint32_t)p.w * 3u;
    const uint8_t* src = p.frames + (int64_t)img * p.stride + (int64_t)p.crop_t * p.pitch + (int64_t)p.crop_l * 3;

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        uint32_t total = 0;
        for (int r = 0; r < nrows; ++r) total += make_window(src + (int64_t)(row0 + r) * p.pitch, rowbytes).bytes;
        mbar_expect_tx(&bar, total);
        for (int r = 0; r < nrows; ++r) {
            const Window w = make_window(src + (int64_t)(row0 + r) * p.pitch, rowbytes);
            bulk_g2s(smem + (size_t)r * p.smem_row_stride, w.base, w.bytes, &bar);
        }
    }
    fill_hsv_tables(sdiv, hdiv, tid, kMaskThreads);
    __syncthreads();
    mbar_wait(&bar, 0);

    const int gpr = p.wpr * 8;            // groups of 4 pixels per row, padded to whole words
    const int items = nrows * gpr;
    uint32_t* out_img = p.bits + (int64_t)img * p.bits_stride;
    for (int base = 0; base < items; base += kMaskThreads) {
        const int it = base + tid;
        uint32_t nib = 0;
        int r = 0, gi = 0;
        if (it < items) {
            r = it / gpr;
            gi = it - r * gpr;
            const int x = gi * 4;
            if (x < p.w) {
                const int y = row0 + r;
                const uint8_t* s = src + (int64_t)y * p.pitch;
                const uint32_t shift = (uint32_t)((uintptr_t)s & 15);
                const uint4 raw = window_load16(smem + (size_t)r * p.smem_row_stride, shift + (uint32_t)x * 3u);
                uint32_t bgr[4];
                bgr[0] = raw.x & 0x00ffffffu;
                bgr[1] = __funnelshift_r(raw.x, raw.y, 24) & 0x00ffffffu;
                bgr[2] = __funnelshift_r(raw.y, raw.z, 16) & 0x00ffffffu;
                bgr[3] = raw.z >> 8;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (x + k < p.w) {
                        const uint32_t X = bgr_to_hsv_packed(bgr[k] & 255u, (bgr[k] >> 8) & 255u, bgr[k] >> 16, sdiv, hdiv);
                        if (!hsv_excluded(X, x + k, y, p.f)) nib |= 1u << k;     // kept by the key => alpha 255 => foreground
                    }
                }
            }
        }
        uint32_t v = nib << (4 * (gi & 7));
        v |= __shfl_xor_sync(0xffffffffu, v, 1);
        v |= __shfl_xor_sync(0xffffffffu, v, 2);
        v |= __shfl_xor_sync(0xffffffffu, v, 4);
        if (it < items && (gi & 7) == 0) out_img[(int64_t)(row0 + r) * p.wpr + (gi >> 3)] = v;
    }
}

// ===================================================================== stage A: mask, table-driven path
// Used when every filter covers the whole image (no zones) and at most 7 filters have a real hue range —
// the reference's own use (filtres_liste.py:186-199).  V and S of OpenCV's HSV depend only on
// (v = max, d = max - min), so "which filters accept this pixel's V and S" is one byte of a 64 KB table
// T[v][d] living in shared memory for the whole kernel (built once per call by hsv_lut_kernel; exact, same
// sdiv table).  A filter whose H range is [0, >=180] needs nothing else: the pixel is keyed out by one table
// hit.  Hue is computed only for pixels that pass V/S of a filter with a real H range.
struct HsvFast {
    int n_hue;                       // filters with a real hue range (<= 7), table bits 0..6
    int h_lo[7], h_hi[7];
    // "V <= T, any H, any S" filter: a group of pixels whose bytes are all <= T is keyed out at once.
    // dark_mode 0: no such filter; 1: T < 128 (byte > T <=> bit7 | carry(low7 + 127 - T));
    // 2: T >= 128 (byte > T <=> bit7 & carry(low7 + 255 - T)).  dark_add = the per-byte addend.
    int dark_mode;
    uint32_t dark_add;
    // byte s (0 = max channel is R, 1 = G, 2 = B): hue filters whose range meets the hues that sector can produce
    // (R: [150,180) u [0,30], G: [30,90], B: [90,150])
    uint32_t sector_mask;
};

// table byte: bit 7 = V,S accepted by a filter that accepts every hue (pixel keyed out at once);
// bits 0..6 = V,S accepted by hue filter j (hue test still to do)
__global__ void hsv_lut_kernel(const HsvFilterSet f, uint8_t* __restrict__ lut) {
    const int v = blockIdx.x, d = threadIdx.x;
    uint32_t m = 0;
    if (d <= v) {
        const int s = (d * g_hsv_sdiv[v] + (1 << 11)) >> 12;
        const uint32_t X = ((uint32_t)s << 10) | ((uint32_t)v << 20);
        constexpr uint32_t kGuardSV = (1u << 19) | (1u << 29);
        int j = 0;
        for (int k = 0; k < f.n; ++k) {
            const bool sv = (((X + (f.clo[k] & ~0x3ffu)) & ((f.chi[k] & ~0x3ffu) - X)) & kGuardSV) == kGuardSV;
            const int lo = 512 - (int)(f.clo[k] & 0x3ffu), hi = (int)(f.chi[k] & 0x3ffu) - 512;
            const bool any_hue = lo <= 0 && hi >= 180;
            if (any_hue) {
                if (sv) m |= 0x80u;
            } else {
                if (sv) m |= 1u << j;
                ++j;
            }
        }
    }
    lut[v * 256 + d] = (uint8_t)m;
}

__device__ __forceinline__ int hue_of(int b, int g, int r, int v, int d, const int* hdiv) {
    int hn;
    if (v == r) hn = g - b;
    else if (v == g) hn = b - r + 2 * d;
    else hn = r - g + 4 * d;
    int h = (hn * hdiv[d] + (1 << 11)) >> 12;
    if (h < 0) h += 180;
    return h;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// ===================================================================== stage A: mask, segment kernel
// Every WARP runs its own pipeline:
//
//  * work unit = a SEGMENT: up to 512 pixels of one row (1536 bytes, 16 mask words).  Warps draw
//    batches of 16 consecutive segments from a global counter (dynamic balance: segments that cross
//    the object cost several times more than backdrop segments).
//  * each warp owns a private 3-stage ring in shared memory; lane 0 keeps two bulk copies (UBLKCP)
//    in flight while the warp converts the third — 32 warps x 2 x 1.5 KB = 96 KB in flight per SM,
//    no CTA-wide barrier anywhere, no producer warp.
//  * pass 1 (whole warp, 16 pixels per lane): conservative "any byte above the dark threshold" test
//    on the four aligned 16-byte vectors covering the lane's two 8-pixel groups.  With a threshold
//    T >= 128 the first cut is just the OR of the vectors' bit 7 (0.5 op per word); only when many
//    groups are flagged is the exact SWAR compare run to thin them out.
//  * pass 2 (flagged groups only): a task = HALF a group (4 pixels, two 16x2 DPX min/max pairs), so a
//    segment with a dozen bright specks still fills most of the warp's lanes in one round.  V/S
//    classification is one byte of the 64 KB table per pixel; the hue is computed only for pixels
//    the table marks as candidates of a filter with a real hue range.
//  * mask nibbles are assembled in a per-warp 128-byte scratch and leave as 16 coalesced words.
struct MaskSegParams {
    const uint8_t* frames;
    int64_t pitch, stride;
    int crop_l, crop_t, w, h;
    int wpr;
    uint32_t* bits;
    int64_t bits_stride;             // words per image
    int segs_per_row, segs_per_frame;
    uint32_t n_segs;
    unsigned int* counter;           // zeroed before the launch
    const uint8_t* lut;
    HsvFast f;
};

constexpr int kSegPx = 512;
#ifndef IPP_SEG_THREADS
#define IPP_SEG_THREADS 1024
#endif
constexpr int kSegThreads = IPP_SEG_THREADS;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegStages = 3;
constexpr uint32_t kSegStageBytes = 1552;          // 1536 + 15 bytes of misalignment, rounded up to 16 (covers every whole-vector / whole-word read)
#ifndef IPP_SEG_BATCH
#define IPP_SEG_BATCH 16
#endif
constexpr int kSegBatch = IPP_SEG_BATCH;
#ifndef IPP_SEG_REFINE
#define IPP_SEG_REFINE 20
#endif
constexpr int kSegRefine = IPP_SEG_REFINE;          // more cheaply-flagged groups than this: run the exact dark test
constexpr size_t kSegSmem = 65536 + (size_t)kSegWarps * kSegStages * kSegStageBytes + (size_t)kSegWarps * 192 + (size_t)kSegWarps * kSegBatch * 16;
static_assert(kSegBatch >= 1 && kSegBatch <= 32, "a batch's descriptors are computed by one lane each");

__global__ void __launch_bounds__(kSegThreads, 1024 / kSegThreads) chain_mask_seg_kernel(const MaskSegParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[kSegWarps][kSegStages];
    __shared__ int hdiv[256];
    __shared__ uint8_t h2f[192];                 // hue (0..179) -> bit j set when hue filter j accepts it
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 192) {
        uint32_t mk = 0;
        for (int j = 0; j < p.f.n_hue; ++j) mk |= (uint32_t)(tid >= p.f.h_lo[j] && tid <= p.f.h_hi[j]) << j;
        h2f[tid] = (uint8_t)mk;
    }
    uint8_t* lut = smem;
    uint8_t* ring = smem + 65536 + (size_t)warp * kSegStages * kSegStageBytes;
    uint8_t* wq = smem + 65536 + (size_t)kSegWarps * kSegStages * kSegStageBytes + (size_t)warp * 192;   // 64 queue bytes
    uint8_t* nib = wq + 64;                                                                              // 128 nibble bytes

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kSegStages; ++s) mbar_init(&full[warp][s], 1);
    }
    for (int i = tid; i < 65536 / 16; i += kSegThreads)
        reinterpret_cast<uint4*>(lut)[i] = __ldg(reinterpret_cast<const uint4*>(p.lut) + i);
    for (int i = tid; i < 256; i += kSegThreads) hdiv[i] = g_hsv_hdiv[i];
    __syncthreads();

    // ---- issue side.  The descriptors of a drawn batch (source window, geometry, first mask word) are computed by its first
    // lanes in parallel, one segment each, and parked in shared memory; requesting a segment is then one broadcast load plus
    // lane 0's bulk copy (walking row / frame boundaries with warp-uniform pointer arithmetic cost ~45 instructions per segment,
    // an eighth of the kernel).
    uint4* desc = reinterpret_cast<uint4*>(smem + 65536 + (size_t)kSegWarps * kSegStages * kSegStageBytes + (size_t)kSegWarps * 192) + warp * kSegBatch;
    uint32_t is_left = 0, is_k = 0;              // segments left in the drawn batch, next descriptor
    bool is_done = false;
    // per stage: word index of the segment's first mask word, and npx | shift << 16 (npx == 0: no segment)
    uint32_t d_out[kSegStages], d_geo[kSegStages];

    auto issue = [&](int s) {
        if (!is_done && is_left == 0) {
            uint32_t b = 0;
            if (lane == 0) b = atomicAdd(p.counter, (unsigned)kSegBatch);
            b = __shfl_sync(0xffffffffu, b, 0);
            if (b >= p.n_segs) {
                is_done = true;
            } else {
                is_left = min((uint32_t)kSegBatch, p.n_segs - b);
                is_k = 0;
                __syncwarp();                    // every lane has read the last descriptor of the previous batch
                if ((uint32_t)lane < is_left) {
                    const uint32_t seg = b + (uint32_t)lane;
                    const uint32_t frame = seg / (uint32_t)p.segs_per_frame;
                    const uint32_t rem = seg - frame * (uint32_t)p.segs_per_frame;
                    const uint32_t row = rem / (uint32_t)p.segs_per_row;
                    const uint32_t sx = rem - row * (uint32_t)p.segs_per_row;
                    const int x0 = (int)sx * kSegPx;
                    const int npx = min(kSegPx, p.w - x0);
                    const uint8_t* src = p.frames + (int64_t)frame * p.stride + (int64_t)(p.crop_t + (int)row) * p.pitch + (int64_t)(p.crop_l + x0) * 3;
                    const uint32_t shift = (uint32_t)((uintptr_t)src & 15u);
                    const uint64_t base = (uint64_t)(uintptr_t)(src - shift);
                    desc[lane] = make_uint4((uint32_t)base, (uint32_t)(base >> 32), (uint32_t)npx | (shift << 16),
                                            (uint32_t)((int64_t)frame * p.bits_stride + (int64_t)row * p.wpr) + sx * (kSegPx / 32));
                }
                __syncwarp();
            }
        }
        if (is_done) {
            d_geo[s] = 0;
            return;
        }
        const uint4 d = desc[is_k];
        d_out[s] = d.w;
        d_geo[s] = d.z;
        if (lane == 0) {
            const uint32_t bytes = ((d.z >> 16) + (d.z & 0xffffu) * 3u + 15u) & ~15u;
            mbar_expect_tx(&full[warp][s], bytes);
            bulk_g2s(ring + (size_t)s * kSegStageBytes, reinterpret_cast<const void*>((uintptr_t)(((uint64_t)d.y << 32) | d.x)), bytes, &full[warp][s]);
        }
        ++is_k;
        --is_left;
    };

    const int dark_mode = p.f.dark_mode;
    const uint32_t dark_add = p.f.dark_add;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t sm_r = p.f.sector_mask & 0x7fu, sm_g = (p.f.sector_mask >> 8) & 0x7fu, sm_b = (p.f.sector_mask >> 16) & 0x7fu;

#pragma unroll
    for (int s = 0; s < kSegStages - 1; ++s) issue(s);

    for (uint32_t round = 0;; ++round) {
        bool finished = false;
#pragma unroll
        for (int s = 0; s < kSegStages; ++s) {
            if (finished) break;
            // refill the stage consumed one step ago (all lanes are past it: __syncwarp at the end of the step)
            if (s == 0) issue(kSegStages - 1); else issue(s - 1);
            const uint32_t geo = d_geo[s];
            const int npx = (int)(geo & 0xffffu);
            if (npx == 0) { finished = true; break; }
            const uint32_t shift = geo >> 16;
            // stage s is refilled once per round (stage kSegStages-1 first in the prologue's absence: see parity below)
            mbar_wait(&full[warp][s], round & 1u);
            const uint8_t* tile = ring + (size_t)s * kSegStageBytes;
            const int ngroups = (npx + 7) >> 3;

            // ---------------------------------------------------------------- pass 1
            const int g0 = 2 * lane;
            const bool inA = g0 < ngroups, inB = g0 + 1 < ngroups;
            bool needA = inA, needB = inB;
            if (dark_mode != 0) {
                const uint4* vec = reinterpret_cast<const uint4*>(tile + 48 * lane);
                const uint4 q0 = vec[0], q1 = vec[1], q2 = vec[2], q3 = vec[3];
                const uint32_t mA2 = shift > 8u ? 0xffffffffu : 0u, mB1 = shift < 8u ? 0xffffffffu : 0u, mB3 = shift > 0u ? 0xffffffffu : 0u;
                uint32_t f0, f1, f2, f3;
                if (dark_mode == 2) {                 // T >= 128: a byte above T has bit 7 set (conservative first cut)
                    f0 = q0.x | q0.y | q0.z | q0.w; f1 = q1.x | q1.y | q1.z | q1.w;
                    f2 = q2.x | q2.y | q2.z | q2.w; f3 = q3.x | q3.y | q3.z | q3.w;
                } else {                              // T < 128: exact, bit 7 or carry out of the low 7 bits
#define IPP_DARK1(q) (((q.x & 0x7f7f7f7fu) + dark_add) | q.x | ((q.y & 0x7f7f7f7fu) + dark_add) | q.y | \
                      ((q.z & 0x7f7f7f7fu) + dark_add) | q.z | ((q.w & 0x7f7f7f7fu) + dark_add) | q.w)
                    f0 = IPP_DARK1(q0); f1 = IPP_DARK1(q1); f2 = IPP_DARK1(q2); f3 = IPP_DARK1(q3);
#undef IPP_DARK1
                }
                needA = inA && (((f0 | f1 | (f2 & mA2)) & 0x80808080u) != 0u);
                needB = inB && ((((f1 & mB1) | f2 | (f3 & mB3)) & 0x80808080u) != 0u);
                if (dark_mode == 2) {
                    const int cheap = __popc(__ballot_sync(0xffffffffu, needA)) + __popc(__ballot_sync(0xffffffffu, needB));
                    if (cheap > kSegRefine) {         // busy segment: exact compare (bit 7 AND carry of low7 + 255 - T)
#define IPP_DARK2(q) ((((q.x & 0x7f7f7f7fu) + dark_add) & q.x) | (((q.y & 0x7f7f7f7fu) + dark_add) & q.y) | \
                      (((q.z & 0x7f7f7f7fu) + dark_add) & q.z) | (((q.w & 0x7f7f7f7fu) + dark_add) & q.w))
                        f0 = IPP_DARK2(q0); f1 = IPP_DARK2(q1); f2 = IPP_DARK2(q2); f3 = IPP_DARK2(q3);
#undef IPP_DARK2
                        needA = inA && (((f0 | f1 | (f2 & mA2)) & 0x80808080u) != 0u);
                        needB = inB && ((((f1 & mB1) | f2 | (f3 & mB3)) & 0x80808080u) != 0u);
                    }
                }
            }
            const uint32_t balA = __ballot_sync(0xffffffffu, needA), balB = __ballot_sync(0xffffffffu, needB);
            if ((balA | balB) == 0u) {                                      // nothing above the dark threshold: sixteen zero words
                if (lane < 16 && 32 * lane < npx) p.bits[(size_t)d_out[s] + lane] = 0u;
                continue;
            }
            reinterpret_cast<uint32_t*>(nib)[lane] = 0u;                    // 16 pixels keyed out unless pass 2 says otherwise
            const int nA = __popc(balA);
            const int ntasks = 2 * (nA + __popc(balB));
            if (needA) wq[__popc(balA & lt_mask)] = (uint8_t)g0;
            if (needB) wq[nA + __popc(balB & lt_mask)] = (uint8_t)(g0 + 1);
            __syncwarp();

            // ---------------------------------------------------------------- pass 2: half-group tasks
            for (int t0 = 0; t0 < ntasks; t0 += 32) {
                const int t = t0 + lane;
                if (t < ntasks) {
                    const int hg = 2 * (int)wq[t >> 1] + (t & 1);           // half-group index: pixels 4*hg .. 4*hg+3
                    const uint32_t boff = shift + 12u * (uint32_t)hg;
                    const uint32_t bs = (boff & 3u) * 8u;
                    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(tile + (boff & ~3u));
                    uint32_t rx = w32[0], ry = w32[1], rz = w32[2];
                    if (bs) {
                        const uint32_t r3 = w32[3];
                        rx = __funnelshift_r(rx, ry, bs);
                        ry = __funnelshift_r(ry, rz, bs);
                        rz = __funnelshift_r(rz, r3, bs);
                    }
                    // rx ry rz = b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3
                    const uint32_t B01 = __byte_perm(rx, 0u, 0x4340), G01 = __byte_perm(rx, ry, 0x0401) & 0x00ff00ffu;
                    const uint32_t R01 = __byte_perm(rx, ry, 0x0502) & 0x00ff00ffu;
                    const uint32_t B23 = __byte_perm(ry, rz, 0x0502) & 0x00ff00ffu, G23 = __byte_perm(ry, rz, 0x0603) & 0x00ff00ffu;
                    const uint32_t R23 = __byte_perm(rz, 0u, 0x4340);
                    // max with predicates (DPX): gb = (g >= b), rm = (r >= max(b, g)) per pixel — OpenCV's hue sector
                    // (v == r first, then v == g, else b) falls out of the max computation for free
                    bool gb0, gb1, gb2, gb3, rm0, rm1, rm2, rm3;
                    const uint32_t X01 = __vibmax_s16x2(G01, B01, &gb1, &gb0), X23 = __vibmax_s16x2(G23, B23, &gb3, &gb2);
                    const uint32_t V01 = __vibmax_s16x2(R01, X01, &rm1, &rm0), V23 = __vibmax_s16x2(R23, X23, &rm3, &rm2);
                    const uint32_t D01 = V01 - __vimin3_s16x2(B01, G01, R01), D23 = V23 - __vimin3_s16x2(B23, G23, R23);
                    const uint32_t I01 = (V01 << 8) | D01, I23 = (V23 << 8) | D23;
                    const uint32_t M = (uint32_t)lut[I01 & 0xffffu] | ((uint32_t)lut[I01 >> 16] << 8) |
                                       ((uint32_t)lut[I23 & 0xffffu] << 16) | ((uint32_t)lut[I23 >> 16] << 24);
                    const uint32_t hit = (M & 0x80808080u) >> 7;
                    uint32_t ex = ((hit * 0x00204081u) >> 21) & 15u;        // bit 7 of every byte -> 4 adjacent bits
                    // hue filters whose range the pixel's sector can reach at all (a bluish object never meets an orange filter)
                    const uint32_t S = (rm0 ? sm_r : gb0 ? sm_g : sm_b) | ((rm1 ? sm_r : gb1 ? sm_g : sm_b) << 8) |
                                       ((rm2 ? sm_r : gb2 ? sm_g : sm_b) << 16) | ((rm3 ? sm_r : gb3 ? sm_g : sm_b) << 24);
                    const uint32_t M2 = M & S & ~(hit * 0x7fu);             // hue-filter candidates of pixels not keyed out yet
                    // hue test of the candidates, one pixel per iteration and lane (bytes re-read from the tile): a lane
                    // with a single bright speck loops once, a lane inside the object four times
                    uint32_t m = M2;
                    while (m) {
                        const int k = (__ffs(m) - 1) >> 3;
                        const uint32_t cand = (m >> (8 * k)) & 0x7fu;
                        m &= ~(0xffu << (8 * k));
                        const uint8_t* t = tile + boff + 3 * k;
                        const int b = t[0], g = t[1], r = t[2];
                        const int v = max(max(b, g), r), d = v - min(min(b, g), r);
                        const int hh = hue_of(b, g, r, v, d, hdiv);
                        if (h2f[hh] & cand) ex |= 1u << k;
                    }
                    uint32_t nb = ~ex & 15u;
                    const int left = npx - 4 * hg;                          // partial last group of the row
                    if (left < 4) nb &= (1u << max(left, 0)) - 1u;
                    nib[hg] = (uint8_t)nb;
                }
            }
            __syncwarp();

            // ---------------------------------------------------------------- nibbles -> mask words
            uint32_t x = reinterpret_cast<const uint32_t*>(nib)[lane];      // four nibbles: pixels 16*lane .. 16*lane+15
            x = (x | (x >> 4)) & 0x00ff00ffu;
            x = (x | (x >> 8)) & 0xffffu;
            const uint32_t hi = __shfl_down_sync(0xffffffffu, x, 1);
            if ((lane & 1) == 0 && 16 * lane < npx) p.bits[(size_t)d_out[s] + (lane >> 1)] = x | (hi << 16);
            __syncwarp();                                                   // everyone is done with this stage, the queue and the nibbles
        }
        if (finished) break;
    }
}

// ===================================================================== shared: inverse map
struct SrcMap {
    int kind, a0, a1, a2, a3, a4, a5, w, h;     // w, h = source sprite (component rect) size
};
// canvas pixel (cx, cy) -> source sprite pixel; false when outside the source
__device__ __forceinline__ bool map_canvas(const SrcMap& m, int cx, int cy, int& xin, int& yin) {
    switch (m.kind) {
        case 0:
            xin = (m.a2 + cy * m.a1 + cx * m.a0) >> 16;
            yin = (m.a5 + cy * m.a4 + cx * m.a3) >> 16;
            break;
        case 1: xin = cx; yin = cy; break;
        case 2: xin = m.w - 1 - cx; yin = m.h - 1 - cy; break;
        case 3: xin = m.w - 1 - cy; yin = cx; break;          // ROTATE_90: out[y][x] = in[x][w-1-y]
        default: xin = cy; yin = m.h - 1 - cx; break;         // ROTATE_270: out[y][x] = in[h-1-x][y]
    }
    return xin >= 0 && xin < m.w && yin >= 0 && yin < m.h;
}

// flat tile index -> record index (the last record whose tile_start is <= tile).  Called by the 32 lanes of one warp: every
// round probes 32 records at once and narrows the range 33-fold, i.e. two dependent L2 round trips for up to 1089 records
// where a scalar binary search needs ten — the search sits at the head of every CTA of the tile-list kernels.
template <typename Rec, typename GetStart>
__device__ __forceinline__ int find_record(const Rec* recs, int n, int tile, GetStart start_of) {
    const int lane = threadIdx.x & 31;
    int lo = 0, hi = n;                             // the answer lies in [lo, hi)
    while (hi - lo > 1) {
        const int step = (hi - lo + 32) / 33;
        const int idx = lo + step * (lane + 1);
        const bool le = idx < hi && start_of(recs[idx]) <= tile;        // monotone over the lanes
        const int k = __popc(__ballot_sync(0xffffffffu, le));
        hi = min(hi, lo + step * (k + 1));
        lo += step * k;
    }
    return lo;
}

// ===================================================================== stage B1: rotated bbox
// PIL's getbbox of the rotated sprite WITHOUT touching the canvas pixel by pixel.
//
// A canvas pixel P is set iff its source S = map(P) (16.16 fixed point, nearest) is a kept pixel.
// If P is extreme (say the left-most set pixel), its neighbour P' = P - (1, 0) is not set; the
// fixed-point steps |a0|, |a3| (and |a1|, |a4| for vertical moves) are <= 1.0, so map(P') lies in the
// 8-neighbourhood of S (or is S itself) — hence S has an unset 8-neighbour (or lies on the rect's
// border): S is a BOUNDARY pixel of the kept component.  Conversely every canvas pixel that maps
// to a given S lies within +-1 of round(forward(S)): the pre-image of a unit square under a rotation
// is a unit square around forward(centre of S), per-axis half extent <= 0.7072, plus <= 0.5 of rounding.
// So: scan the component rect's mask words (bits only), peel the 8-boundary with shifts and ANDs,
// forward-map each boundary pixel in float, and test the 3x3 canvas pixels around it with the EXACT
// integer inverse map.  Work ~ perimeter instead of canvas area; the result is identical.
// One thread = four mask words; a CTA = 256 consecutive word quads of one sprite's rect (flat tile list).
__global__ void __launch_bounds__(256) chain_rot_bbox_kernel(const ipp_chain_rot* __restrict__ rots, int n_rot,
                                                             const uint32_t* __restrict__ bits, int64_t bits_stride, int wpr,
                                                             int frame_h, int32_t* bbox_out) {
    __shared__ int s_rec;
    if (threadIdx.x < 32) {
        const int found = find_record(rots, n_rot, (int)blockIdx.x, [](const ipp_chain_rot& r) { return r.tile_start; });
        if (threadIdx.x == 0) s_rec = found;
    }
    __syncthreads();
    const ipp_chain_rot rec = rots[s_rec];
    const int chunk = (int)blockIdx.x - rec.tile_start;
    int32_t* out = bbox_out + 4 * s_rec;
    if (rec.kind != 0) {
        // copy / rot90 / rot180 / rot270 are exact permutations and the rect is the component's tight box
        if (chunk == 0 && threadIdx.x == 0) { out[0] = 0; out[1] = 0; out[2] = rec.canvas_w; out[3] = rec.canvas_h; }
        return;
    }
    // one thread = FOUR consecutive mask words of a row (a 16-byte load when rows are 16-byte aligned):
    // 9 loads per 128 pixels for the 3 x 3 neighbourhood instead of 9 per 32
    const int q0 = (rec.cc_x >> 5) >> 2;
    const int nq = ((((rec.cc_x + rec.cc_w + 31) >> 5) + 3) >> 2) - q0;
    const int q = chunk * 256 + (int)threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t* fb = bits + (int64_t)rec.frame * bits_stride;
    int minx = INT_MAX, miny = INT_MAX, maxx1 = 0, maxy1 = 0;
    // ---- phase 1: boundary pixels of this thread's four words
    uint32_t bnd[4] = {0u, 0u, 0u, 0u};
    int row = 0, wi = 0;
    if (q < nq * rec.cc_h) {
        row = q / nq;
        wi = 4 * (q0 + (q - row * nq));
        const int fy = rec.cc_y + row;
        const bool vec = ((wpr & 3) == 0) && ((((uintptr_t)fb) & 15) == 0);
        // words wi-1 .. wi+4 of a mask row (zero outside the row / when the row does not exist)
        auto load6 = [&](int y, bool exists, uint32_t (&v)[6]) {
#pragma unroll
            for (int k = 0; k < 6; ++k) v[k] = 0u;
            if (!exists) return;
            const uint32_t* r = fb + (int64_t)y * wpr;
            if (vec) {
                const uint4 t = *reinterpret_cast<const uint4*>(r + wi);
                v[1] = t.x; v[2] = t.y; v[3] = t.z; v[4] = t.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) if (wi + k < wpr) v[1 + k] = r[wi + k];
            }
            if (wi > 0) v[0] = r[wi - 1];
            if (wi + 4 < wpr) v[5] = r[wi + 4];
        };
        uint32_t mid[6];
        load6(fy, true, mid);
        if (mid[1] | mid[2] | mid[3] | mid[4]) {
            uint32_t up[6], dn[6];
            load6(fy - 1, fy > 0, up);
            load6(fy + 1, fy + 1 < frame_h, dn);
            // pixels whose left, own and right neighbours are all set, for one row
            auto hcore = [](const uint32_t (&v)[6], int k) { return v[k] & ((v[k] << 1) | (v[k - 1] >> 31)) & ((v[k] >> 1) | (v[k + 1] << 31)); };
#pragma unroll
            for (int k = 0; k < 4; ++k) bnd[k] = mid[1 + k] & ~(hcore(mid, 1 + k) & hcore(up, 1 + k) & hcore(dn, 1 + k));
        }
    }
    // ---- phase 2: the warp's boundary pixels, compacted (a flat edge puts 32 of them in one word), word slot by word slot
    __shared__ uint32_t s_q[8][1024];
    // forward map (float): canvas = A^-1 * (65536 * (S + 0.5) - (a2, a5))
    const double det = (double)rec.a0 * (double)rec.a4 - (double)rec.a1 * (double)rec.a3;
    const float i00 = (float)((double)rec.a4 / det), i01 = (float)(-(double)rec.a1 / det);
    const float i10 = (float)(-(double)rec.a3 / det), i11 = (float)((double)rec.a0 / det);
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        uint32_t bk = bnd[k];
        const int cnt = __popc(bk);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;                                  // warp-uniform
        int off = incl - cnt;
        const uint32_t hi = (uint32_t)row << 16;
        const int xbase = (wi + k) * 32 - rec.cc_x;
        while (bk) {
            const int bit = __ffs(bk) - 1;
            bk &= bk - 1;
            s_q[warp][off++] = (uint32_t)(xbase + bit) | hi;
        }
        __syncwarp();
        for (int i = lane; i < total; i += 32) {
            const uint32_t e = s_q[warp][i];
            const int xin = (int)(e & 0xffffu), yin = (int)(e >> 16);
            const float u = (float)((xin << 16) + 32768 - rec.a2), v = (float)((yin << 16) + 32768 - rec.a5);
            const int cx0 = (int)rintf(i00 * u + i01 * v) - 1, cy0 = (int)rintf(i10 * u + i11 * v) - 1;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int cy = cy0 + dy;
                if (cy < 0 || cy >= rec.canvas_h) continue;
                int xx = rec.a2 + cy * rec.a1 + cx0 * rec.a0, yy = rec.a5 + cy * rec.a4 + cx0 * rec.a3;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int cx = cx0 + dx;
                    const int xi = xx >> 16, yi = yy >> 16;
                    xx += rec.a0;
                    yy += rec.a3;
                    if (cx < 0 || cx >= rec.canvas_w) continue;
                    if (cx >= minx && cx < maxx1 && cy >= miny && cy < maxy1) continue;      // cannot extend this lane's box
                    if ((unsigned)xi >= (unsigned)rec.cc_w || (unsigned)yi >= (unsigned)rec.cc_h) continue;
                    const int fx = rec.cc_x + xi;
                    if ((fb[(int64_t)(rec.cc_y + yi) * wpr + (fx >> 5)] >> (fx & 31)) & 1u) {
                        minx = min(minx, cx); maxx1 = max(maxx1, cx + 1);
                        miny = min(miny, cy); maxy1 = max(maxy1, cy + 1);
                    }
                }
            }
        }
        __syncwarp();                                              // the queue is refilled by the next word slot
    }
    minx = warp_min(minx); miny = warp_min(miny); maxx1 = warp_max(maxx1); maxy1 = warp_max(maxy1);
    if ((threadIdx.x & 31) == 0 && maxx1 > 0) {
        if (minx < __ldcg(&out[0])) atomicMin(&out[0], minx);
        if (miny < __ldcg(&out[1])) atomicMin(&out[1], miny);
        if (maxx1 > __ldcg(&out[2])) atomicMax(&out[2], maxx1);
        if (maxy1 > __ldcg(&out[3])) atomicMax(&out[3], maxy1);
    }
}

// ===================================================================== stage B2
struct B2Common {
    const uint8_t* frames;
    int64_t pitch, stride;
    int crop_l, crop_t;
    const uint32_t* bits;
    int64_t bits_stride;
    int wpr;
    const ipp_chain_item* items;
    int n_items;
    const int32_t* tables;
    uint2* band;                 // banded int8 fragment tables (chain_band_kernel)
    uint8_t* sprite_ws;
    int h_blocks32, v_blocks32;  // fragment tables of that axis in 32-position blocks (shared-memory passes) instead of 64-position super-blocks
};

// gather: sprite pixel (X, Y) after flip <- canvas (X' + fit_x0, Y' + fit_y0) <- frame pixel.
// Stored premultiplied (alpha is 0 or 255 here) as FOUR BYTE PLANES [R, G, B, A][spr_h][Wp] (Wp = spr_w
// rounded up to 16): the layout the tensor-core LANCZOS pass reads its A fragments from.
constexpr int kGatherRows = 8;          // rows per lane: tile = 32 columns x 64 rows
__global__ void __launch_bounds__(256) chain_gather_kernel(const B2Common c) {
    __shared__ int s_rec;
    if (threadIdx.x < 32) {
        const int found = find_record(c.items, c.n_items, (int)blockIdx.x, [](const ipp_chain_item& r) { return r.gather_tile_start; });
        if (threadIdx.x == 0) s_rec = found;
    }
    __syncthreads();
    const ipp_chain_item& it = c.items[s_rec];
    const int spr_w = it.spr_w, spr_h = it.spr_h;
    const int tile = (int)blockIdx.x - it.gather_tile_start;
    const int tiles_x = (spr_w + 31) >> 5;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int X = tx * 32 + lane;
    const int Y0 = ty * (8 * kGatherRows) + warp * kGatherRows;
    if (X >= spr_w || Y0 >= spr_h) return;
    const SrcMap m{it.kind, it.a0, it.a1, it.a2, it.a3, it.a4, it.a5, it.cc_w, it.cc_h};
    const uint32_t* fb = c.bits + (int64_t)it.frame * c.bits_stride + (int64_t)it.cc_y * c.wpr;
    const uint8_t* frame = c.frames + (int64_t)it.frame * c.stride + (int64_t)(c.crop_t + it.cc_y) * c.pitch +
                           (int64_t)(c.crop_l + it.cc_x) * 3;
    const int64_t Wp = (spr_w + 15) & ~15;
    const int64_t plane = (int64_t)spr_h * Wp;
    uint8_t* dst = c.sprite_ws + it.rot_off + X + (int64_t)Y0 * Wp;
    const int cx = (it.flip_h ? (spr_w - 1 - X) : X) + it.fit_x0;
    // affine case: canvas row cy = fit_y0 + (flip_v ? spr_h-1-Y : Y) moves by +-1 per sprite row
    const int cy0 = (it.flip_v ? (spr_h - 1 - Y0) : Y0) + it.fit_y0;
    const int dy = it.flip_v ? -1 : 1;
    int xx = it.a2 + cy0 * it.a1 + cx * it.a0, yy = it.a5 + cy0 * it.a4 + cx * it.a3;
    const int sx = dy * it.a1, sy = dy * it.a4;
    // phase 1: all mask words and all pixel bytes of the lane's rows are requested before anything is used
    // (the bit test and the three byte loads of a pixel do not wait for each other; pixels outside the
    // component rect are fetched from a clamped in-rect address and discarded)
    uint32_t word[kGatherRows], w0[kGatherRows], w1[kGatherRows];
    int sh[kGatherRows], bsh[kGatherRows];
#pragma unroll
    for (int j = 0; j < kGatherRows; ++j) {
        int xin, yin;
        bool inside;
        if (it.kind == 0) {
            xin = xx >> 16;
            yin = yy >> 16;
            inside = (unsigned)xin < (unsigned)it.cc_w && (unsigned)yin < (unsigned)it.cc_h;
        } else {
            inside = map_canvas(m, cx, cy0 + dy * j, xin, yin);
        }
        xx += sx;
        yy += sy;
        const bool use = inside && Y0 + j < spr_h;
        if (!use) { xin = 0; yin = 0; }
        const int fx = it.cc_x + xin;
        sh[j] = use ? (fx & 31) : -1;
        word[j] = fb[yin * c.wpr + (fx >> 5)];
        // the pixel's three bytes come from the aligned word that holds its first byte and, only when they
        // spill over (half of the pixels), the next one: 1.5 loads per pixel instead of three byte loads
        const uintptr_t a = (uintptr_t)(frame + (int64_t)yin * c.pitch + xin * 3);
        const uint32_t* aw = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        bsh[j] = (int)(a & 3) * 8;
        w0[j] = aw[0];
        w1[j] = bsh[j] > 8 ? aw[1] : 0u;
    }
    // phase 2: select, premultiply (alpha is 0 or 255), store the four planes
#pragma unroll
    for (int j = 0; j < kGatherRows; ++j) {
        if (Y0 + j >= spr_h) break;
        const bool on = sh[j] >= 0 && ((word[j] >> sh[j]) & 1u);
        uint8_t* d = dst + (int64_t)j * Wp;
        const uint32_t bgr = on ? __funnelshift_r(w0[j], w1[j], bsh[j]) : 0u;
        d[0] = (uint8_t)(bgr >> 16);                         // BGR -> R, G, B, A planes
        d[plane] = (uint8_t)(bgr >> 8);
        d[2 * plane] = (uint8_t)bgr;
        d[3 * plane] = on ? (uint8_t)255 : (uint8_t)0;
    }
}

// --------------------------------------------------------------------- LANCZOS on the tensor cores
// Pillow's 8bpc resample is an exact integer convolution: out = clip8((2^21 + sum px * k) >> 22), k a
// signed 23-bit coefficient, ~15-40 taps per output and pass.  On the CUDA cores every tap costs four
// multiply-adds plus unpacking (the scalar kernels ran at ~70 % issue utilisation and 0.9 TB/s); as a
// banded integer GEMM the same sums are a handful of IMMA instructions:
//     coefficient k = k2 * 65536 + k1 * 256 + k0   (k0, k1 unsigned bytes, k2 a signed byte)
//     sum px * k   = (sum px * k2) << 16  +  (sum px * k1) << 8  +  sum px * k0
// with three u8 x {u8, u8, s8} -> s32 accumulators (mma.sync.m16n8k32), exact in two's complement because
// the full sum fits 32 bits (Pillow accumulates in int32 as well).
//
// The reduction index of an MMA may be permuted freely as long as A and B agree, so the window a tile of
// 8 outputs reads is cut into SUPER-BLOCKS of 64 source positions = two k-blocks, ordered such that one
// aligned 16-byte load per lane and row IS the lane's share of both A fragments: lane (g, t) reads bytes
// [16t, 16t+16) of row g; its four words are {k-block 0: a0, a2; k-block 1: a0, a2}.  (Fetching fragments in
// their natural order costs four 4-byte loads per row group with 16 useful bytes per 128-byte line — the L1
// tag stage then bounds the kernel.)  chain_band_kernel lays the coefficient bands out in the same order.
// A tile touches ceil(7 * in / out) + ksize + 16 source positions (window drift over 7 outputs + one window
// + alignment of the first position to 16): one super-block for scale factors up to ~3.5.
// A skipped pass (size unchanged; Pillow skips it) uses the identity band k = 2^22, which reproduces
// the input bytes exactly, so both kernels have a single code path.
constexpr int kPrec = 32 - 8 - 2;
__device__ __forceinline__ uint32_t c_clip8(int v) {
    uint32_t r;
    asm("cvt.sat.u8.s32 %0, %1;" : "=r"(r) : "r"(v >> kPrec));      // clamp to [0, 255] in one instruction
    return r;
}
// libImaging RGBa -> RGBA: c = min(255, 255 * c / a).  One division per pixel instead of three: for n = 255 c < 2^16 and
// a <= 255, n / a == umulhi(n, floor((2^32 - 1) / a) + 1) exactly (the multiplier's error times n stays below 2^32).
__device__ __forceinline__ uint32_t c_unpremultiply(uint32_t v) {
    const uint32_t a = v >> 24;
    if (a == 0u || a == 255u) return v;
    if (a == 1u) return (v & 0xff000000u) | (((v & 255u) ? 255u : 0u)) | (((v >> 8) & 255u) ? 0xff00u : 0u) | (((v >> 16) & 255u) ? 0xff0000u : 0u);
    const uint32_t M = 0xffffffffu / a + 1u;
    const uint32_t c0 = min(255u, __umulhi(255u * (v & 255u), M));
    const uint32_t c1 = min(255u, __umulhi(255u * ((v >> 8) & 255u), M));
    const uint32_t c2 = min(255u, __umulhi(255u * ((v >> 16) & 255u), M));
    return c0 | (c1 << 8) | (c2 << 16) | (a << 24);
}

__device__ __forceinline__ void mma_u8u8(int (&d)[4], const uint32_t (&a)[4], const uint2 b) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void mma_u8s8(int (&d)[4], const uint32_t (&a)[4], const uint2 b) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

// layout of one axis' band table at `base` (uint2 units): [ceil(ntiles / 2)] header words holding the
// first source position xb of every tile (int32 each, multiple of 16), then
// [ntiles][nsb super-blocks][2 k-blocks][3 byte planes][32 lanes] uint2.
struct BandAxis {
    int in_size, out_size, ksize, nsb, ntiles;
    const int32_t* bounds;       // [out][2] (xmin, n)   (unused when ksize == 0)
    const int32_t* kk;           // [out][ksize]
};
__device__ __forceinline__ BandAxis band_axis(const ipp_chain_item& it, const int32_t* tables, int axis) {
    BandAxis a;
    a.in_size = axis ? it.spr_h : it.spr_w;
    a.out_size = axis ? it.new_h : it.new_w;
    a.ksize = axis ? it.v_ksize : it.h_ksize;
    a.nsb = axis ? (it.nkb >> 16) : (it.nkb & 0xffff);
    a.ntiles = (a.out_size + 7) >> 3;
    a.bounds = tables + (axis ? it.vb_off : it.hb_off);
    a.kk = tables + (axis ? it.vk_off : it.hk_off);
    return a;
}

// one warp = one (item, axis, tile of 8 outputs, super-block): the 64 x 8 coefficient band in B-fragment order,
// split into three byte planes.  B fragment of m16n8k32: lane (g = lane / 4, t = lane % 4) holds, for output
// column g, reduction indices 4t..4t+3 (register 0) and 16+4t.. (register 1) of a k-block; with the permuted
// order of the A side, index (k-block kbl, register h, byte j) is source position 16t + 4(2 kbl + h) + j.
__global__ void __launch_bounds__(256) chain_band_kernel(const B2Common c) {
    const ipp_chain_item& it = c.items[blockIdx.y];
    const int lane = threadIdx.x & 31;
    int w = (int)blockIdx.x * 8 + (int)(threadIdx.x >> 5);
    const BandAxis ah = band_axis(it, c.tables, 0), av = band_axis(it, c.tables, 1);
    const int nh_warps = ah.ntiles * ah.nsb;
    const int axis = w >= nh_warps;
    if (axis) w -= nh_warps;
    const BandAxis a = axis ? av : ah;
    if (w >= a.ntiles * a.nsb) return;
    const int tile = w / a.nsb, sb = w - tile * a.nsb;
    uint2* base = c.band + (axis ? it.vband_off : it.hband_off);
    const bool blocks32 = axis ? c.v_blocks32 : c.h_blocks32;
    const int X0 = tile * 8;
    const int xfirst = a.ksize ? a.bounds[2 * X0] : X0;
    const int xb = blocks32 ? (xfirst & ~7) : (xfirst & ~15);
    const int g = lane >> 2, t = lane & 3;
    const int X = X0 + g;
    int xmin = 0, n = 0;
    const int32_t* k = nullptr;
    if (X < a.out_size) {
        if (a.ksize) {
            xmin = a.bounds[2 * X];
            n = a.bounds[2 * X + 1];
            k = a.kk + (int64_t)X * a.ksize;
        } else {
            xmin = X;               // identity: one tap of weight 1.0
            n = 1;
        }
    }
    auto weight = [&](int pos) {
        const int idx = pos - xmin;
        return (idx >= 0 && idx < n) ? (k ? k[idx] : (1 << kPrec)) : 0;
    };
    uint2* frag = base + ((a.ntiles + 1) >> 1) + ((int64_t)(tile * a.nsb + sb) * 6) * 32 + lane;
    if (blocks32) {
        // 32-position blocks 2 sb, 2 sb + 1 of the tile; the tile's entry carries how many blocks hold a non-zero tap
        const int last = __reduce_max_sync(0xffffffffu, xmin + n);                  // one past the last source position of the tile
        if (sb == 0 && lane == 0) reinterpret_cast<int32_t*>(base)[tile] = xb | (max(1, (last - xb + 31) >> 5) << 24);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            uint32_t r[3][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}};
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int wgt = weight(xb + 32 * (2 * sb + blk) + 8 * t + 4 * half + j);
                    r[0][half] |= ((uint32_t)wgt & 255u) << (8 * j);
                    r[1][half] |= (((uint32_t)wgt >> 8) & 255u) << (8 * j);
                    r[2][half] |= (((uint32_t)wgt >> 16) & 255u) << (8 * j);
                }
            }
            frag[(3 * blk + 0) * 32] = make_uint2(r[0][0], r[0][1]);
            frag[(3 * blk + 1) * 32] = make_uint2(r[1][0], r[1][1]);
            frag[(3 * blk + 2) * 32] = make_uint2(r[2][0], r[2][1]);
        }
        return;
    }
    if (sb == 0 && lane == 0) reinterpret_cast<int32_t*>(base)[tile] = xb;
#pragma unroll
    for (int kbl = 0; kbl < 2; ++kbl) {
        uint32_t r[3][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}};
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int wgt = weight(xb + 64 * sb + 16 * t + 4 * (2 * kbl + half) + j);
                r[0][half] |= ((uint32_t)wgt & 255u) << (8 * j);
                r[1][half] |= (((uint32_t)wgt >> 8) & 255u) << (8 * j);
                r[2][half] |= (((uint32_t)wgt >> 16) & 255u) << (8 * j);      // signed byte: |wgt| < 2^23
            }
        }
        frag[(3 * kbl + 0) * 32] = make_uint2(r[0][0], r[0][1]);
        frag[(3 * kbl + 1) * 32] = make_uint2(r[1][0], r[1][1]);
        frag[(3 * kbl + 2) * 32] = make_uint2(r[2][0], r[2][1]);
    }
}

// the six MMAs of one super-block: data words lo (row g) / hi (row g + 8), coefficient fragments b[0..5]
__device__ __forceinline__ void mma_superblock(int (&acc)[3][4], const uint4& lo, const uint4& hi, const uint2 (&b)[6]) {
    const uint32_t a0[4] = {lo.x, hi.x, lo.y, hi.y}, a1[4] = {lo.z, hi.z, lo.w, hi.w};
    mma_u8u8(acc[0], a0, b[0]); mma_u8u8(acc[1], a0, b[1]); mma_u8s8(acc[2], a0, b[2]);
    mma_u8u8(acc[0], a1, b[3]); mma_u8u8(acc[1], a1, b[4]); mma_u8s8(acc[2], a1, b[5]);
}
__device__ __forceinline__ void load_band(uint2 (&b)[6], const uint2* frag) {
#pragma unroll
    for (int q = 0; q < 6; ++q) b[q] = __ldg(frag + q * 32);
}
// 16 rows of one byte plane -> acc.  lo / hi: the staged first super-block words of rows g, g+8; further
// super-blocks (rare: scale factors above ~3.5) are fetched here through r_lo (row g; row g + 8 is 8 * pitch further).
// acc[0] starts at the rounding constant 2^21 of clip8((2^21 + sum) >> 22).
__device__ __forceinline__ void mma_rows(int (&acc)[3][4], const uint4& lo, const uint4& hi, const uint8_t* r_lo, int64_t pitch,
                                         const uint2 (&b0)[6], const uint2* frag, int nsb) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[0][i] = 1 << (kPrec - 1); acc[1][i] = 0; acc[2][i] = 0; }
    mma_superblock(acc, lo, hi, b0);
    for (int sb = 1; sb < nsb; ++sb) {
        uint2 b[6];
        load_band(b, frag + sb * 192);
        const uint4 lo2 = *reinterpret_cast<const uint4*>(r_lo + 64 * sb);
        const uint4 hi2 = *reinterpret_cast<const uint4*>(r_lo + 8 * pitch + 64 * sb);
        mma_superblock(acc, lo2, hi2, b);
    }
}

// Pull the first two 32-position blocks of a tile's fragment table (1536 bytes = 12 lines) into L1 ahead of use: the
// table reads are the only global loads of a tile and would otherwise cost one L2 round trip per tile and warp.
__device__ __forceinline__ void band_prefetch(const uint2* tile_frags, int lane) {
    if (lane < 12) asm volatile("prefetch.global.L1 [%0];" :: "l"(tile_frags + lane * 16));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// Shared-memory passes: 32-position blocks.  Lane (g, t) reads 8 bytes of rows g and g + 8 per block (positions 8t .. 8t+7
// of the block = both k-halves of its A fragment); the tile's entry says how many blocks carry non-zero taps, usually one
// for down-scales up to ~2.5x — half the MMAs of a 64-position super-block.
__device__ __forceinline__ void mma_rows32(int (&acc)[3][4], const uint8_t* r_lo, int64_t pitch, const uint2 (&b0)[3],
                                           const uint2* frag, int nblk) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[0][i] = 1 << (kPrec - 1); acc[1][i] = 0; acc[2][i] = 0; }
    // four 32-bit loads straight into the fragment's register order (64-bit loads would need the halves re-shuffled)
    const uint32_t s_lo = smem_u32(r_lo), s_hi = smem_u32(r_lo + 8 * pitch);
    {
        const uint32_t a[4] = {lds32(s_lo), lds32(s_hi), lds32(s_lo + 4), lds32(s_hi + 4)};
        mma_u8u8(acc[0], a, b0[0]); mma_u8u8(acc[1], a, b0[1]); mma_u8s8(acc[2], a, b0[2]);
    }
#pragma unroll 1
    for (int blk = 1; blk < nblk; ++blk) {
        const uint2 b[3] = {__ldg(frag + blk * 96), __ldg(frag + blk * 96 + 32), __ldg(frag + blk * 96 + 64)};
        const uint32_t a[4] = {lds32(s_lo + 32 * blk), lds32(s_hi + 32 * blk), lds32(s_lo + 32 * blk + 4), lds32(s_hi + 32 * blk + 4)};
        mma_u8u8(acc[0], a, b[0]); mma_u8u8(acc[1], a, b[1]); mma_u8s8(acc[2], a, b[2]);
    }
}

// Same, for rows stored pairwise interleaved (the fused gather's layout): words of rows g and g + 8 alternate every four
// positions, so ONE 16-byte load per lane and block is the whole A fragment {a0, a1, a2, a3} — no register shuffling.
__device__ __forceinline__ void mma_rows32i(int (&acc)[3][4], const uint8_t* pair, const uint2 (&b0)[3], const uint2* frag, int nblk) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[0][i] = 1 << (kPrec - 1); acc[1][i] = 0; acc[2][i] = 0; }
    {
        const uint4 q = *reinterpret_cast<const uint4*>(pair);
        const uint32_t a[4] = {q.x, q.y, q.z, q.w};
        mma_u8u8(acc[0], a, b0[0]); mma_u8u8(acc[1], a, b0[1]); mma_u8s8(acc[2], a, b0[2]);
    }
#pragma unroll 1
    for (int blk = 1; blk < nblk; ++blk) {
        const uint2 b[3] = {__ldg(frag + blk * 96), __ldg(frag + blk * 96 + 32), __ldg(frag + blk * 96 + 64)};
        const uint4 q = *reinterpret_cast<const uint4*>(pair + 64 * blk);
        const uint32_t a[4] = {q.x, q.y, q.z, q.w};
        mma_u8u8(acc[0], a, b[0]); mma_u8u8(acc[1], a, b[1]); mma_u8s8(acc[2], a, b[2]);
    }
}

// ---- data pipeline of the two passes.  A miss to HBM costs > 1000 cycles here while a tile's MMAs + epilogue are
// ~100 instructions, so every warp keeps kRing - 1 tiles (16 rows x 64 bytes = 1 KB each) in flight with 16-byte
// cp.async (LDGSTS: no registers held while the data travels).  Chunk id = lane + 32 j sits at byte 16 * id of the
// stage: lane (g, t) = 4g + t stages exactly the two 16-byte pieces (row g, row g + 8; column 16 t) it later reads
// back as its A-fragment words — no cross-lane traffic, conflict-free 128-bit shared accesses.
#ifndef IPP_RING
#define IPP_RING 4
#endif
constexpr int kRing = IPP_RING;
// g0: this lane's piece of rows 0..7 of the tile (row lane / 4, column 16 * (lane % 4)); rows 8..15 are 8 * pitch further.
// Rows past the end of a plane are NOT clamped: the planes of a sprite are adjacent and its workspace is padded by
// 16 rows (ipp_chain_layout_host), so the reads stay inside the buffer; their results are never stored.
__device__ __forceinline__ void ring_issue(uint8_t* stage, const uint8_t* g0, int64_t pitch, bool valid, int lane) {
    if (valid) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(stage + 16 * lane)), "l"(g0) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(stage + 16 * lane + 512)), "l"(g0 + 8 * pitch) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");          // (possibly empty) group: keeps the wait count uniform
}
__device__ __forceinline__ void ring_wait() {
    asm volatile("cp.async.wait_group %0;" :: "n"(kRing - 2) : "memory");
}

// horizontal pass.  CTA = (item, byte plane, group of 64 outputs); warp = 8 outputs, whose first super-block of
// coefficients stays in registers; loop over the rows, 32 at a time.  Result is written TRANSPOSED,
// [plane][X][y] with pitch Hp, so that the vertical pass finds its reduction index contiguous: a warp
// transposes its 32 x 8 byte tile through shared memory and stores 32-byte runs.
__global__ void __launch_bounds__(256) chain_hpass_kernel(const B2Common c) {
    __shared__ int s_rec;
    __shared__ __align__(16) uint8_t s_t[8][8 * 36];
    __shared__ __align__(128) uint8_t s_ring[8][kRing][1024];
    if (threadIdx.x < 32) {
        const int found = find_record(c.items, c.n_items, (int)blockIdx.x, [](const ipp_chain_item& r) { return r.h_tile_start; });
        if (threadIdx.x == 0) s_rec = found;
    }
    __syncthreads();
    const ipp_chain_item& it = c.items[s_rec];
    const int lt = (int)blockIdx.x - it.h_tile_start;
    const int pl = lt & 3, xgroup = lt >> 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int new_w = it.new_w, spr_h = it.spr_h;
    const int ntiles = (new_w + 7) >> 3, nsb = it.nkb & 0xffff;
    const int tile = xgroup * 8 + warp;
    if (tile >= ntiles) return;
    const uint2* base = c.band + it.hband_off;
    const int xb = reinterpret_cast<const int32_t*>(base)[tile];
    const uint2* frag = base + ((ntiles + 1) >> 1) + (int64_t)tile * nsb * 192 + lane;
    const int64_t Wp = (it.spr_w + 15) & ~15, Hp = (spr_h + 15) & ~15;
    const uint8_t* src = c.sprite_ws + it.rot_off + (int64_t)pl * spr_h * Wp + xb + 16 * t;
    uint8_t* dst = c.sprite_ws + it.tmp_off + (int64_t)pl * new_w * Hp;
    uint8_t* st = s_t[warp];
    uint2 b0[6];
    load_band(b0, frag);
    const int n_rt = (spr_h + 15) >> 4;
    uint8_t* ring = s_ring[warp][0];
    const int64_t tile_step = 16 * Wp;
    const uint8_t* gnext = src - 16 * t + (int64_t)(lane >> 2) * Wp + 16 * (lane & 3);     // this lane's staging piece of tile 0
    const uint8_t* rcur = src + (int64_t)g * Wp;                                            // row g of the current tile (extra super-blocks)
#pragma unroll
    for (int d = 0; d < kRing - 1; ++d) {
        ring_issue(ring + 1024 * d, gnext, Wp, d < n_rt, lane);
        gnext += tile_step;
    }
    for (int ti = 0; ti < n_rt; ++ti) {
        ring_wait();                                               // tile ti has landed (this lane's own two pieces)
        const uint8_t* stg = ring + 1024 * (ti % kRing);
        const uint4 lo = *reinterpret_cast<const uint4*>(stg + 16 * lane), hi = *reinterpret_cast<const uint4*>(stg + 16 * lane + 512);
        ring_issue(ring + 1024 * ((ti + kRing - 1) % kRing), gnext, Wp, ti + kRing - 1 < n_rt, lane);
        gnext += tile_step;
        int acc[3][4];
        mma_rows(acc, lo, hi, rcur, Wp, b0, frag, nsb);
        rcur += tile_step;
        const int m = ti & 1, y0 = (ti & ~1) << 4;
        // C fragment: {row g, col 2t}, {row g, col 2t+1}, {row g+8, col 2t}, {row g+8, col 2t+1}
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int v = acc[0][i] + (acc[1][i] << 8) + (acc[2][i] << 16);
            const int yl = 16 * m + g + 8 * (i >> 1), xl = 2 * t + (i & 1);
            st[xl * 36 + yl] = (uint8_t)c_clip8(v);
        }
        if (m == 1 || ti + 1 == n_rt) {
            __syncwarp();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int wi = lane + 32 * h;                       // 8 columns x 8 words
                const int xl = wi >> 3, q = wi & 7;
                const int X = tile * 8 + xl, y = y0 + 4 * q;
                if (X < new_w && y < spr_h)
                    *reinterpret_cast<uint32_t*>(dst + (int64_t)X * Hp + y) = *reinterpret_cast<const uint32_t*>(st + xl * 36 + 4 * q);
            }
            __syncwarp();
        }
    }
}

// ---- fused gather + horizontal pass.  The two kernels above exchange the gathered sprite through global memory
// (1.7 MB written and re-read ~3.5x per 1080p sprite).  Here a CTA gathers SIXTEEN SPRITE ROWS x all columns x four byte
// planes straight into shared memory (rows padded to 64 mod 128 bytes: the 128-bit A-fragment reads of a quarter warp
// then hit all 32 banks) and its eight warps run the MMA tiles of those rows from there: no planar sprite in global
// memory, no load latency in the MMA loop.  Used whenever the widest sprite of the batch fits (ipp_chain_layout_host).
__host__ __device__ inline int gh_row_pitch(int spr_w) { return ((spr_w + 64 + 127) / 128) * 128 + 64; }   // >= spr_w + 128, == 64 mod 128
// rows of the staged vertical pass: Hp bytes (multiple of 16) + room for the 32-byte skew, == 64 mod 128.  Windows that run
// past a row's end read the next row's bytes against zero coefficients.
__host__ __device__ inline int vs_row_pitch(int spr_h) { return ((((spr_h + 15) & ~15) + 32 - 64 + 127) / 128) * 128 + 64; }
__host__ __device__ inline int gh_row_skew(int r) { return (r & 2) * 16; }      // rows 2, 3 (mod 4): +32 bytes (same for r and r + 8)
#ifndef IPP_GH_FLY
#define IPP_GH_FLY 6
#endif
constexpr int kGhFly = IPP_GH_FLY;      // patches per lane in flight
constexpr int kGhEntries = 256;      // tile entries kept in shared memory (sprites up to 2048 columns; wider ones read the rest from global)
constexpr int kGhTail = 1024;          // slack behind the planes: windows of multi-super-block tiles may run past the last row

__global__ void __launch_bounds__(256, 3) chain_gh_kernel(const B2Common c) {
    extern __shared__ __align__(128) uint8_t gh_smem[];
    __shared__ int s_rec;
    __shared__ __align__(16) uint8_t s_t[8][8 * 20];
    __shared__ int s_ent[kGhEntries];
    if (threadIdx.x < 32) {
        const int found = find_record(c.items, c.n_items, (int)blockIdx.x, [](const ipp_chain_item& r) { return r.gather_tile_start; });
        if (threadIdx.x == 0) s_rec = found;
    }
    __syncthreads();
    const ipp_chain_item& it = c.items[s_rec];
    const int spr_w = it.spr_w, spr_h = it.spr_h, new_w = it.new_w;
    const int y0 = ((int)blockIdx.x - it.gather_tile_start) * 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // shared-memory layout of a plane: eight row pairs (r, r + 8) of pitch P2 = 64 mod 128, the two rows of a pair
    // interleaved every four columns: byte (r, x) at (r & 7) * P2 + skew + 8 * (x / 4) + 4 * (r / 8) + x % 4
    const int P2 = 2 * gh_row_pitch(spr_w) - 64;
    const int plane_bytes = 8 * P2;
    // tile entries and the first tile's fragments are requested before the gather so that phase 2 finds them on chip
    const int ntiles = (new_w + 7) >> 3, nsb = it.nkb & 0xffff;
    const uint2* base = c.band + it.hband_off;
    const uint2* frags = base + ((ntiles + 1) >> 1);
    for (int i = threadIdx.x; i < min(ntiles, kGhEntries); i += 256) s_ent[i] = reinterpret_cast<const int32_t*>(base)[i];
    if (warp < ntiles) band_prefetch(frags + (int64_t)warp * nsb * 192, lane);

    // ---------------------------------------------------------------- phase 1: gather rows y0 .. y0+15 into shared memory
    {
        const SrcMap m{it.kind, it.a0, it.a1, it.a2, it.a3, it.a4, it.a5, it.cc_w, it.cc_h};
        const uint32_t* fb = c.bits + (int64_t)it.frame * c.bits_stride + (int64_t)it.cc_y * c.wpr;
        const uint8_t* frame = c.frames + (int64_t)it.frame * c.stride + (int64_t)(c.crop_t + it.cc_y) * c.pitch +
                               (int64_t)(c.crop_l + it.cc_x) * 3;
        const uint32_t frame_lsb = (uint32_t)((uintptr_t)frame & 3);     // rect origin = 4-byte aligned base + 0..3
        const uint8_t* frame_al = frame - frame_lsb;
        const uint32_t pitch32 = (uint32_t)c.pitch;
        // column X of the sprite is canvas column cxb + cxs * X (flip folded in); in the affine case the fixed-point source
        // coordinates are then linear in X: (ax + X * dax) >> 16, (ay + X * day) >> 16
        const int cxs = it.flip_h ? -1 : 1, cxb = (it.flip_h ? spr_w - 1 : 0) + it.fit_x0;
        const int dax = cxs * it.a0, day = cxs * it.a3;
        const int cc_w = it.cc_w, cc_h = it.cc_h, cc_x = it.cc_x;       // registers: the record itself lives in global memory
        const uint8_t* fb_bytes = reinterpret_cast<const uint8_t*>(fb);
        const int wpr = c.wpr;
        // A warp gathers 8 x 4 patches (lane = column lx of row ly of the patch): the source pixels of a patch are a
        // rotated 8 x 4 block of the frame, i.e. ~10 sectors per request where 32 pixels along one rotated line touch up to
        // 32.  Warp w owns rows 4*(w&3) .. +3 and every other group of 8 * kGhFly columns; kGhFly patches per lane are in flight.
        // Row pairs 2, 3 (mod 4) sit 32 bytes further right in shared memory (gh_row_skew), which puts the four rows of a
        // patch into different banks and keeps the fragment loads of phase 2 conflict-free.
        auto gather_rows = [&](auto affine) {
            const int lx = lane & 7, r = 4 * (warp & 3) + (lane >> 3), Y = y0 + r;
            const bool row_ok = Y < spr_h;
            const int cy = (it.flip_v ? (spr_h - 1 - Y) : Y) + it.fit_y0;
            const int ax = it.a2 + cy * it.a1 + cxb * it.a0, ay = it.a5 + cy * it.a4 + cxb * it.a3;
            uint8_t* drow = gh_smem + (r & 7) * P2 + gh_row_skew(r) + 4 * (r >> 3) + lx + (lx & 4);
            for (int X0 = (warp >> 2) * (8 * kGhFly); X0 < spr_w; X0 += 16 * kGhFly) {
                uint32_t word[kGhFly], w0[kGhFly], w1[kGhFly];
                int sh[kGhFly], bsh[kGhFly];
#pragma unroll
                for (int j = 0; j < kGhFly; ++j) {
                    const int X = X0 + 8 * j + lx;
                    int xin, yin;
                    bool inside;
                    if constexpr (decltype(affine)::value) {
                        xin = (ax + X * dax) >> 16;
                        yin = (ay + X * day) >> 16;
                        inside = ((unsigned)xin < (unsigned)cc_w) & ((unsigned)yin < (unsigned)cc_h);
                    } else {
                        inside = map_canvas(m, cxb + cxs * X, cy, xin, yin);
                    }
                    const bool use = inside && X < spr_w && row_ok;
                    if (!use) { xin = 0; yin = 0; }
                    const int fx = cc_x + xin;
                    sh[j] = use ? (fx & 31) : -1;
                    word[j] = *reinterpret_cast<const uint32_t*>(fb_bytes + ((uint32_t)(yin * wpr + (fx >> 5)) << 2));   // 32-bit byte offset
                    // 32-bit byte offset inside the component rect (a frame is far smaller than 4 GB)
                    const uint32_t off = (uint32_t)yin * pitch32 + (uint32_t)xin * 3u + frame_lsb;
                    const uint32_t* aw = reinterpret_cast<const uint32_t*>(frame_al + (off & ~3u));
                    bsh[j] = (int)(off & 3u) * 8;
                    w0[j] = aw[0];
                    w1[j] = bsh[j] > 8 ? aw[1] : 0u;
                }
                uint8_t* q0 = drow + 2 * X0;
                uint8_t* q1 = q0 + plane_bytes;
                uint8_t* q2 = q1 + plane_bytes;
                uint8_t* q3 = q2 + plane_bytes;
#pragma unroll
                for (int j = 0; j < kGhFly; ++j) {
                    {   // columns past the sprite / rows past its end are written too (zeros, into the padding of the row pair)
                        const bool on = sh[j] >= 0 && ((word[j] >> sh[j]) & 1u);
                        const uint32_t bgr = on ? __funnelshift_r(w0[j], w1[j], bsh[j]) : 0u;
                        // X0 is a multiple of 8: 8 * (X / 4) + X % 4 = 2 * X0 + 16 j + lx + (lx & 4)
                        q0[16 * j] = (uint8_t)(bgr >> 16);              // BGR -> R, G, B, A planes (premultiplied, alpha 0 / 255)
                        q1[16 * j] = (uint8_t)(bgr >> 8);
                        q2[16 * j] = (uint8_t)bgr;
                        q3[16 * j] = on ? (uint8_t)255 : (uint8_t)0;
                    }
                }
            }
        };
#ifndef IPP_GH_SKIP1
        if (it.kind == 0) gather_rows(std::true_type{});
        else gather_rows(std::false_type{});
#endif
    }
    __syncthreads();

    // ---------------------------------------------------------------- phase 2: MMA tiles of these 16 rows, all planes
    const int g = lane >> 2, t = lane & 3;
    const int64_t Hp = (spr_h + 15) & ~15;
    const int64_t plane_out = (int64_t)new_w * Hp;
    uint8_t* st = s_t[warp];
#ifdef IPP_GH_SKIP2
    if (c.n_items >= 0) return;
#endif
    for (int tile = warp; tile < ntiles; tile += 8) {
        const int entry = tile < kGhEntries ? s_ent[tile] : reinterpret_cast<const int32_t*>(base)[tile];
        const int xb = entry & 0xffffff, nblk = entry >> 24;
        const uint2* frag = frags + (int64_t)tile * nsb * 192 + lane;
        const uint2 b0[3] = {__ldg(frag), __ldg(frag + 32), __ldg(frag + 64)};
        if (tile + 8 < ntiles) band_prefetch(frags + (int64_t)(tile + 8) * nsb * 192, lane);
        const uint8_t* r_lo = gh_smem + g * P2 + gh_row_skew(g) + 2 * xb + 16 * t;  // row pair (g, g + 8) of plane 0, this lane's 8 columns
        const int xl = lane >> 2, q = lane & 3;                         // transposed store: 8 columns x 4 words of 4 rows
        const int X = tile * 8 + xl, y = y0 + 4 * q;
        uint8_t* outp = c.sprite_ws + it.tmp_off + (int64_t)X * Hp + y;
        const bool out_ok = X < new_w && y < spr_h;
#pragma unroll 1
        for (int pl = 0; pl < 4; ++pl) {
            int acc[3][4];
            mma_rows32i(acc, r_lo + pl * plane_bytes, b0, frag, nblk);
            // C fragment: {row g, col 2t}, {row g, col 2t+1}, {row g+8, col 2t}, {row g+8, col 2t+1} -> transposed [X][y]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int v = acc[0][i] + (acc[1][i] << 8) + (acc[2][i] << 16);
                st[(2 * t + (i & 1)) * 20 + g + 8 * (i >> 1)] = (uint8_t)c_clip8(v);
            }
            __syncwarp();
            if (out_ok) *reinterpret_cast<uint32_t*>(outp) = *reinterpret_cast<const uint32_t*>(st + xl * 20 + 4 * q);
            outp += plane_out;
            __syncwarp();
        }
    }
}

// vertical pass + un-premultiply.  CTA = (item, group of 64 output rows, chunk of 128 columns); warp = 8 output
// rows; the reduction runs over y, contiguous in the transposed planar intermediate.  All four byte planes of a
// tile are accumulated by the same warp, so un-premultiply and the interleaved RGBA store are lane-local:
// C fragment element (row = column X of the sprite, col = output row Y).
__global__ void __launch_bounds__(256) chain_vpass_kernel(const B2Common c) {
    __shared__ int s_rec;
    __shared__ __align__(128) uint8_t s_ring[8][kRing][1024];
    if (threadIdx.x < 32) {
        const int found = find_record(c.items, c.n_items, (int)blockIdx.x, [](const ipp_chain_item& r) { return r.v_tile_start; });
        if (threadIdx.x == 0) s_rec = found;
    }
    __syncthreads();
    const ipp_chain_item& it = c.items[s_rec];
    const int lt = (int)blockIdx.x - it.v_tile_start;
    const int new_w = it.new_w, new_h = it.new_h, spr_h = it.spr_h;
    const int xchunks = (new_w + 127) >> 7;
    const int ygroup = lt / xchunks, xchunk = lt - ygroup * xchunks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ntiles = (new_h + 7) >> 3, nsb = it.nkb >> 16;
    const int tile = ygroup * 8 + warp;
    if (tile >= ntiles) return;
    const uint2* base = c.band + it.vband_off;
    const int yb = reinterpret_cast<const int32_t*>(base)[tile];
    const uint2* frag = base + ((ntiles + 1) >> 1) + (int64_t)tile * nsb * 192 + lane;
    const int64_t Hp = (spr_h + 15) & ~15;
    const int64_t plane = (int64_t)new_w * Hp;
    const uint8_t* src = c.sprite_ws + it.tmp_off + yb + 16 * t;
    uint32_t* dst = reinterpret_cast<uint32_t*>(c.sprite_ws + it.spr_off);
    const int x_end = min(new_w, (xchunk + 1) * 128);
    uint2 b0[6];
    load_band(b0, frag);
    // steps = (16 columns, byte plane), planes innermost; kRing - 1 steps of data in flight (see ring_issue)
    const int x_begin = xchunk * 128;
    const int n_steps = ((x_end - x_begin + 15) >> 4) * 4;
    uint8_t* ring = s_ring[warp][0];
    // pointers advance plane by plane (4 steps), then to the next 16 columns
    const int64_t back = 16 * Hp - 3 * plane;
    const uint8_t* gnext = src - 16 * t + (int64_t)(x_begin + (lane >> 2)) * Hp + 16 * (lane & 3);
    const uint8_t* rcur = src + (int64_t)(x_begin + g) * Hp;
    int issued = 0;
    auto issue = [&]() {
        ring_issue(ring + 1024 * (issued % kRing), gnext, Hp, issued < n_steps, lane);
        gnext += ((issued & 3) == 3) ? back : plane;
        ++issued;
    };
#pragma unroll
    for (int d = 0; d < kRing - 1; ++d) issue();
    uint32_t px[4] = {0u, 0u, 0u, 0u};
    for (int step = 0; step < n_steps; ++step) {
        ring_wait();
        const uint8_t* stg = ring + 1024 * (step % kRing);
        const uint4 lo = *reinterpret_cast<const uint4*>(stg + 16 * lane), hi = *reinterpret_cast<const uint4*>(stg + 16 * lane + 512);
        issue();
        int acc[3][4];
        mma_rows(acc, lo, hi, rcur, Hp, b0, frag, nsb);
        rcur += ((step & 3) == 3) ? back : plane;
        const int pl = step & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int v = acc[0][i] + (acc[1][i] << 8) + (acc[2][i] << 16);
            px[i] |= c_clip8(v) << (8 * pl);
        }
        if (pl == 3) {
            const int X0 = x_begin + 16 * (step >> 2);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int X = X0 + g + 8 * (i >> 1), Y = tile * 8 + 2 * t + (i & 1);
                if (X < x_end && Y < new_h) dst[(int64_t)Y * new_w + X] = c_unpremultiply(px[i]);
                px[i] = 0u;
            }
        }
    }
}

// ---- vertical pass with the intermediate staged per CTA.  chain_vpass_kernel lets every warp re-read its 64-byte window of
// every column through a private cp.async ring (3.4x read amplification, ring bookkeeping in the MMA loop).  Here a CTA owns
// 16 columns of the sprite: the 64 rows [4 planes][16 columns] of the transposed intermediate are contiguous in global memory and
// are bulk-copied once into shared memory (rows padded like chain_gh_kernel's); the warps then run the output tiles (8 rows of
// the final sprite each, all four planes, un-premultiply lane-local) from there.  Used when the tallest sprite of the batch fits.
__global__ void __launch_bounds__(256) chain_vs_kernel(const B2Common c) {
    extern __shared__ __align__(128) uint8_t vs_smem[];
    __shared__ int s_rec;
    __shared__ uint64_t bar;
    if (threadIdx.x < 32) {
        const int found = find_record(c.items, c.n_items, (int)blockIdx.x, [](const ipp_chain_item& r) { return r.v_tile_start; });
        if (threadIdx.x == 0) s_rec = found;
    }
    if (threadIdx.x == 32) mbar_init(&bar, 1);
    __syncthreads();
    const ipp_chain_item& it = c.items[s_rec];
    const int new_w = it.new_w, new_h = it.new_h, spr_h = it.spr_h;
    const int X0 = ((int)blockIdx.x - it.v_tile_start) * 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t Hp = (spr_h + 15) & ~15;
    const int Hps = vs_row_pitch(spr_h);
    const int plane_bytes = 16 * Hps;
    // ---- phase 1: 64 rows (plane, column) of Hp bytes each, one bulk copy per row (columns past the edge repeat the last one)
    // (the phase cannot complete before thread 0's arrive.expect_tx: its single arrival is still pending, so copies that land
    //  earlier only run the transaction count negative for a moment)
    if (threadIdx.x == 0) mbar_expect_tx(&bar, 64u * (uint32_t)Hp);
    if (threadIdx.x < 64) {
        const int pl = threadIdx.x >> 4, r = threadIdx.x & 15;
        const uint8_t* gsrc = c.sprite_ws + it.tmp_off + ((int64_t)pl * new_w + min(X0 + r, new_w - 1)) * Hp;
        bulk_g2s(vs_smem + pl * plane_bytes + r * Hps + gh_row_skew(r), gsrc, (uint32_t)Hp, &bar);
    }
    const int ntiles = (new_h + 7) >> 3, nsb = it.nkb >> 16;
    const uint2* base = c.band + it.vband_off;
    const uint2* frags = base + ((ntiles + 1) >> 1);
    if (warp < ntiles) band_prefetch(frags + (int64_t)warp * nsb * 192, lane);
    int entry = warp < ntiles ? reinterpret_cast<const int32_t*>(base)[warp] : 0;
    mbar_wait(&bar, 0);

    // ---- phase 2: output tiles of 8 rows, all planes per warp
    const int g = lane >> 2, t = lane & 3;
    uint32_t* dst = reinterpret_cast<uint32_t*>(c.sprite_ws + it.spr_off);
    for (int tile = warp; tile < ntiles; tile += 8) {
        const int yb = entry & 0xffffff, nblk = entry >> 24;
        const uint2* frag = frags + (int64_t)tile * nsb * 192 + lane;
        const uint2 b0[3] = {__ldg(frag), __ldg(frag + 32), __ldg(frag + 64)};
        if (tile + 8 < ntiles) {                                         // next tile: entry in flight, fragments towards L1
            entry = reinterpret_cast<const int32_t*>(base)[tile + 8];
            band_prefetch(frags + (int64_t)(tile + 8) * nsb * 192, lane);
        }
        const uint8_t* r_lo = vs_smem + g * Hps + gh_row_skew(g) + yb + 8 * t;
        uint32_t px[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) {
            int acc[3][4];
            mma_rows32(acc, r_lo + pl * plane_bytes, Hps, b0, frag, nblk);
#pragma unroll
            for (int i = 0; i < 4; ++i) px[i] |= c_clip8(acc[0][i] + (acc[1][i] << 8) + (acc[2][i] << 16)) << (8 * pl);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int X = X0 + g + 8 * (i >> 1), Y = tile * 8 + 2 * t + (i & 1);
            if (X < new_w && Y < new_h) dst[(int64_t)Y * new_w + X] = c_unpremultiply(px[i]);
        }
    }
}

// ===================================================================== stage C
struct CParams {
    const uint8_t* bgs;
    int64_t bg_pitch, bg_stride;
    const ipp_chain_item* items;
    const uint8_t* sprite_ws;
    uint8_t* out;
    int64_t out_pitch, out_stride;
    int out_w, out_h;                 // window size (the chain: sq_size x sq_size)
    int rows_per_cta, chunks_per_cta;
    uint32_t smem_row_stride;
};

__global__ void __launch_bounds__(kPasteThreads) chain_paste_kernel(const CParams c) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[2];
    const int tid = threadIdx.x;
    const ipp_chain_item& it = c.items[blockIdx.y];
    PasteParams p;
    p.bg = c.bgs + (int64_t)it.bg * c.bg_stride;
    p.bg_pitch = c.bg_pitch;
    p.sprite = c.sprite_ws + it.spr_off;
    p.sprite_pitch = (int64_t)it.new_w * 4;
    p.sprite_w = it.new_w; p.sprite_h = it.new_h; p.pos_x = it.pos_x; p.pos_y = it.pos_y;
    p.crop_x0 = it.sq_x0; p.crop_y0 = it.sq_y0; p.crop_w = c.out_w; p.crop_h = c.out_h;
    p.dst = c.out + (int64_t)it.out_slot * c.out_stride;
    p.dst_pitch = c.out_pitch;
    // A CTA walks chunks_per_cta (1-2) consecutive chunks of rows_per_cta rows through two shared-memory buffers: the bulk copies of
    // chunk j + 1 are in flight while chunk j is blended and written, and the item record is fetched once per CTA.
    const int first = blockIdx.x * c.rows_per_cta * c.chunks_per_cta;
    if (first >= p.crop_h) return;
    const int nchunks = min(c.chunks_per_cta, (p.crop_h - first + c.rows_per_cta - 1) / c.rows_per_cta);
    const uint32_t rowbytes = (uint32_t)p.crop_w * 3u;
    const size_t buf_bytes = (size_t)c.rows_per_cta * c.smem_row_stride;

    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    __syncthreads();
    auto issue = [&](int j) {                           // warp 0: lane r issues the bulk copy of row r of chunk j (rows_per_cta <= 8)
        const int row0 = first + j * c.rows_per_cta;
        const int nrows = min(c.rows_per_cta, p.crop_h - row0);
        Window w;
        w.base = nullptr; w.bytes = 0; w.shift = 0;
        if (tid < nrows) w = make_window(p.bg + (int64_t)(p.crop_y0 + row0 + tid) * p.bg_pitch + (int64_t)p.crop_x0 * 3, rowbytes);
        const uint32_t total = __reduce_add_sync(0xffffffffu, w.bytes);
        if (tid == 0) mbar_expect_tx(&bar[j & 1], total);
        __syncwarp();                                   // the barrier is armed before any copy can complete on it
        if (w.bytes) bulk_g2s(smem + (j & 1) * buf_bytes + (size_t)tid * c.smem_row_stride, w.base, w.bytes, &bar[j & 1]);
    };
    if (tid < 32) issue(0);
    for (int j = 0; j < nchunks; ++j) {
        if (j + 1 < nchunks) {
            __syncthreads();                            // every warp is done reading the buffer chunk j + 1 lands in (chunk j - 1)
            if (tid < 32) issue(j + 1);
        }
        mbar_wait(&bar[j & 1], (uint32_t)(j >> 1) & 1u);
        const int row0 = first + j * c.rows_per_cta;
        paste_rows_from_smem(p, smem + (j & 1) * buf_bytes, c.smem_row_stride, row0, min(c.rows_per_cta, p.crop_h - row0), tid);
    }
}

// ---- crop windows straight out of pinned host memory.  The copy engines move the 2-D windows (3240 of 5760 bytes per row) at
// ~36 GB/s next to the contiguous frame uploads; here the SMs read the rows themselves (zero-copy loads over PCIe, every aligned
// 16-byte vector of a row exactly once, staged per warp in shared memory and re-cut to the destination's alignment).  A small
// persistent grid: the bytes in flight (8 warps x ~3 KB per CTA) are what the link needs, not the SMs.
constexpr int kWinMaxItems = 256;
struct WinGatherParams {
    const uint8_t* src[kWinMaxItems];      // first byte of every item's window (host, device-accessible)
    int64_t bg_pitch;
    uint8_t* dst;
    int out_size, n;
    uint32_t smem_row_stride;
};
__global__ void __launch_bounds__(256) window_gather_kernel(const __grid_constant__ WinGatherParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* buf = smem + (size_t)warp * p.smem_row_stride;
    const uint32_t rowbytes = (uint32_t)p.out_size * 3u;
    const int total_rows = p.n * p.out_size;
    for (int r = blockIdx.x * 8 + warp; r < total_rows; r += gridDim.x * 8) {
        const int i = r / p.out_size, y = r - i * p.out_size;
        const uint8_t* s = p.src[i] + (int64_t)y * p.bg_pitch;
        const Window w = make_window(s, rowbytes);
        // eight vectors per lane requested before the first is parked: a PCIe round trip is microseconds, the loads of a row must
        // all be in flight together
        const uint32_t nv = w.bytes >> 4;
        for (uint32_t c0 = lane; c0 < nv; c0 += 32 * 8) {
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (c0 + 32 * k < nv) v[k] = ldg_stream16(w.base + ((c0 + 32 * k) << 4));
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (c0 + 32 * k < nv) *reinterpret_cast<uint4*>(buf + ((c0 + 32 * k) << 4)) = v[k];
        }
        __syncwarp();
        uint8_t* d = p.dst + ((int64_t)i * p.out_size + y) * rowbytes;
        const RowSplit sp = split_row(d, rowbytes);
        const uint32_t off0 = w.shift + sp.head;
        switch ((off0 & 15u) >> 2) {
            case 0: copy_row_body<0>(buf, off0, d, sp.head, sp.nb, lane); break;
            case 1: copy_row_body<1>(buf, off0, d, sp.head, sp.nb, lane); break;
            case 2: copy_row_body<2>(buf, off0, d, sp.head, sp.nb, lane); break;
            default: copy_row_body<3>(buf, off0, d, sp.head, sp.nb, lane); break;
        }
        if ((uint32_t)lane < sp.head) d[lane] = buf[w.shift + lane];
        if (lane >= 16 && (uint32_t)(lane - 16) < sp.tail) {
            const uint32_t b = sp.head + (sp.nb << 4) + (uint32_t)(lane - 16);
            d[b] = buf[w.shift + b];
        }
        __syncwarp();
    }
}

// ===================================================================== device LANCZOS tables
struct TabParams {
    const ipp_chain_item* items;
    int32_t* tables;
    int32_t* amb;      // [0] = count; record r at amb + 64 * (r + 1): {item, axis, row, xmin, xmax, ksize, k[0 .. 57]}
    int amb_cap;
    double eps;
};

__device__ __forceinline__ double d_sinc(double x) {
    if (x == 0.0) return 1.0;
    x = __dmul_rn(x, 3.14159265358979323846);
    return __ddiv_rn(sin(x), x);
}
__device__ __forceinline__ double d_lanczos(double x) {
    if (-3.0 <= x && x < 3.0) return __dmul_rn(d_sinc(x), d_sinc(__ddiv_rn(x, 3.0)));
    return 0.0;
}

// one thread = one output row (its bounds pair and ksize coefficients)
__global__ void __launch_bounds__(128) chain_tables_kernel(const TabParams p) {
    const ipp_chain_item& it = p.items[blockIdx.y];
    const int hrows = it.h_ksize ? it.new_w : 0, vrows = it.v_ksize ? it.new_h : 0;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= hrows + vrows) return;
    const int axis = idx >= hrows;
    const int xx = axis ? idx - hrows : idx;
    const int in_size = axis ? it.spr_h : it.spr_w, out_size = axis ? it.new_h : it.new_w;
    const int ksize = axis ? it.v_ksize : it.h_ksize;
    int32_t* bounds = p.tables + (axis ? it.vb_off : it.hb_off) + 2 * xx;
    int32_t* k = p.tables + (axis ? it.vk_off : it.hk_off) + (int64_t)xx * ksize;
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = __dmul_rn(3.0, filterscale);
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn((double)xx + 0.5, scale);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    // every weight is evaluated ONCE (two double-precision sin() each): windows of up to kTabKeep taps stay in registers /
    // local memory between the normalisation sum and the quantisation; longer windows are re-evaluated
    constexpr int kTabKeep = 24;
    double wv[kTabKeep];
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        const double w = d_lanczos(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
        if (x < kTabKeep) wv[x] = w;
        ww = __dadd_rn(ww, w);
    }
    bool amb = false;
    for (int x = 0; x < ksize; ++x) {
        int c = 0;
        if (x < xmax) {
            double w = x < kTabKeep ? wv[x] : d_lanczos(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            const double v = __dmul_rn(w, 4194304.0);
            const double t = (w < 0) ? __dadd_rn(-0.5, v) : __dadd_rn(0.5, v);
            c = (int)t;
            if (fabs(t - rint(t)) < p.eps) amb = true;
        }
        k[x] = c;
    }
    bounds[0] = xmin;
    bounds[1] = xmax;
    if (amb) {
        const int slot = atomicAdd(&p.amb[0], 1);
        if (slot < p.amb_cap) {
            // the flagged row travels with its flag: the host verifies it from ONE read-back, without touching the tables
            int32_t* rec = p.amb + 64 * (slot + 1);
            rec[0] = (int)blockIdx.y; rec[1] = axis; rec[2] = xx; rec[3] = xmin; rec[4] = xmax; rec[5] = ksize;
            for (int x = 0; x < ksize && x < 58; ++x) rec[6 + x] = k[x];
        }
    }
}

}  // namespace ipp

// ===================================================================== C ABI
extern "C" {

int64_t ipp_chain_mask_workspace_bytes(const ipp_chain_geom* g, int n_frames) {
    if (!g || n_frames <= 0 || g->crop_w <= 0 || g->crop_h <= 0) return 0;
    return ipp::ccl_workspace_bytes(ipp::make_ccl_geom(g->crop_w, g->crop_h), n_frames);
}

static int chain_check_geom(const ipp_chain_geom* g) {
    using namespace ipp;
    IPP_REQUIRE(g != nullptr, IPP_ERR_BAD_ARG, "chain: geometry is NULL");
    IPP_REQUIRE(g->frame_w > 0 && g->frame_h > 0 && g->crop_l >= 0 && g->crop_t >= 0 && g->crop_w > 0 && g->crop_h > 0 &&
                    g->crop_l + g->crop_w <= g->frame_w && g->crop_t + g->crop_h <= g->frame_h,
                IPP_ERR_BAD_SHAPE, "chain: border crop outside the frame");
    IPP_REQUIRE(g->n_ranges >= 1 && g->n_ranges <= IPP_MAX_HSV_RANGES, IPP_ERR_BAD_ARG, "chain: n_ranges=%d", g->n_ranges);
    return IPP_OK;
}

int ipp_chain_stage_a(const uint8_t* frames, int64_t frame_pitch, int64_t frame_stride,
                      const ipp_chain_geom* g, int32_t* info_out,
                      void* workspace, int64_t workspace_bytes, int n_frames, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = chain_check_geom(g);
    if (rc != IPP_OK) return rc;
    IPP_REQUIRE(frames && info_out && n_frames >= 1, IPP_ERR_BAD_ARG, "ipp_chain_stage_a: bad arguments");
    IPP_REQUIRE(frame_pitch >= (int64_t)g->frame_w * 3, IPP_ERR_BAD_SHAPE, "ipp_chain_stage_a: pitch smaller than a row");
    const CclGeom cg = make_ccl_geom(g->crop_w, g->crop_h);
    IPP_REQUIRE(workspace && workspace_bytes >= ccl_workspace_bytes(cg, n_frames) && (((uintptr_t)workspace) & 255) == 0,
                IPP_ERR_WORKSPACE, "ipp_chain_stage_a: workspace too small or misaligned (need %lld bytes)",
                (long long)ccl_workspace_bytes(cg, n_frames));
    const int te = ensure_hsv_tables();
    IPP_REQUIRE(te == 0, te, "ipp_chain_stage_a: table upload failed (%d)", te);
    const CclBuffers b = carve_ccl(workspace, cg, n_frames);
    const HsvFilterSet fs = pack_filters(g->ranges, g->n_ranges, g->crop_w, g->crop_h);
    const uint32_t rowbytes = (uint32_t)g->crop_w * 3u;
    const uint32_t row_stride = staged_row_stride(rowbytes);
    IPP_REQUIRE(row_stride <= 200 * 1024, IPP_ERR_UNSUPPORTED, "ipp_chain_stage_a: row too long");
    // the 64 KB V/S table borrows the head of the label/area scratch: ccl_init (which runs after
    // the mask kernel) writes the run-start entries there before anything reads them
    const bool lut_fits = (int64_t)cg.lab_stride * 4 * n_frames >= 65536;
    int n_hue_filters = 0;
    for (int k = 0; k < fs.n; ++k) {
        const int lo = 512 - (int)(fs.clo[k] & 0x3ffu), hi = (int)(fs.chi[k] & 0x3ffu) - 512;
        n_hue_filters += !(lo <= 0 && hi >= 180);
    }
    const int64_t n_segs = (int64_t)ceil_div(g->crop_w, kSegPx) * g->crop_h * n_frames;
    const bool fast = fs.all_zones_full && n_hue_filters <= 7 && lut_fits && n_segs < (1LL << 31) &&
                      cg.bits_stride * n_frames < (1LL << 32) && getenv("IPP_MASK_GENERIC") == nullptr;
    if (fast) {
        HsvFast hf{};
        int dark_t = -1;
        for (int k = 0; k < fs.n; ++k) {
            const int lo = 512 - (int)(fs.clo[k] & 0x3ffu), hi = (int)(fs.chi[k] & 0x3ffu) - 512;
            if (!(lo <= 0 && hi >= 180)) {
                hf.h_lo[hf.n_hue] = lo <= hi ? lo : (1 << 20);       // empty range: nothing passes the unsigned window test
                hf.h_hi[hf.n_hue] = lo <= hi ? hi : (1 << 20);
                ++hf.n_hue;
            } else {
                // any hue: is it also "any S, V in [0, T]" ?
                const int s_lo = 512 - (int)((fs.clo[k] >> 10) & 0x3ffu), s_hi = (int)((fs.chi[k] >> 10) & 0x3ffu) - 512;
                const int v_lo = 512 - (int)((fs.clo[k] >> 20) & 0x3ffu), v_hi = (int)((fs.chi[k] >> 20) & 0x3ffu) - 512;
                if (s_lo <= 0 && s_hi >= 255 && v_lo <= 0 && v_hi >= 0) dark_t = std::max(dark_t, std::min(v_hi, 255));
            }
        }
        if (dark_t >= 0 && dark_t < 255) {
            hf.dark_mode = dark_t >= 128 ? 2 : 1;
            hf.dark_add = 0x01010101u * (uint32_t)(dark_t >= 128 ? 255 - dark_t : 127 - dark_t);
        }
        for (int j = 0; j < hf.n_hue; ++j) {
            const int lo = hf.h_lo[j], hi = hf.h_hi[j];
            auto meets = [&](int a, int z) { return lo <= z && hi >= a; };
            if (meets(0, 30) || meets(150, 179)) hf.sector_mask |= 1u << j;
            if (meets(30, 90)) hf.sector_mask |= 1u << (8 + j);
            if (meets(90, 150)) hf.sector_mask |= 1u << (16 + j);
        }
        uint8_t* lut = (uint8_t*)b.la;
        { ipp::LaunchTimer _lt("hsv_lut_kernel", st); hsv_lut_kernel<<<256, 256, 0, st>>>(fs, lut); }
        IPP_LAUNCH_CHECK();
        {
            MaskSegParams p{};
            p.frames = frames; p.pitch = frame_pitch; p.stride = frame_stride;
            p.crop_l = g->crop_l; p.crop_t = g->crop_t; p.w = g->crop_w; p.h = g->crop_h;
            p.wpr = cg.wpr; p.bits = b.bits; p.bits_stride = cg.bits_stride;
            p.segs_per_row = ceil_div(g->crop_w, kSegPx);
            p.segs_per_frame = p.segs_per_row * g->crop_h;
            p.n_segs = (uint32_t)n_segs;
            p.counter = (unsigned int*)b.acc;          // ccl_init re-initialises acc afterwards
            p.lut = lut;
            p.f = hf;
            IPP_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), st));
            static bool attr_set[64] = {};              // per device: the attribute belongs to the device's context
            int dev_id = 0;
            IPP_CUDA(cudaGetDevice(&dev_id));
            if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
                IPP_CUDA(cudaFuncSetAttribute(chain_mask_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSegSmem));
                if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
            }
            const int grid = (int)std::min<int64_t>(sm_count(), ceil_div(n_segs, kSegWarps));
            { ipp::LaunchTimer _lt("chain_mask_seg_kernel", st); chain_mask_seg_kernel<<<grid, kSegThreads, kSegSmem, st>>>(p); }
            IPP_LAUNCH_CHECK();
        }
    }
    if (!fast) {
    MaskParams p{};
    p.frames = frames; p.pitch = frame_pitch; p.stride = frame_stride;
    p.crop_l = g->crop_l; p.crop_t = g->crop_t; p.w = g->crop_w; p.h = g->crop_h;
    p.wpr = cg.wpr; p.bits = b.bits; p.bits_stride = cg.bits_stride;
    p.f = fs;
    p.smem_row_stride = row_stride;
    int rows = (int)(24576 / p.smem_row_stride);
    if (rows < 1) rows = 1;
    if (rows > g->crop_h) rows = g->crop_h;
    p.rows_per_cta = rows;
    const size_t smem = (size_t)rows * p.smem_row_stride;
    if (smem > 40 * 1024)
        IPP_CUDA(cudaFuncSetAttribute(chain_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(g->crop_h, rows), n_frames);
    { ipp::LaunchTimer _lt("chain_mask_kernel", st); chain_mask_kernel<<<grid, kMaskThreads, smem, st>>>(p); }
    IPP_LAUNCH_CHECK();
    }
    rc = ccl_run(cg, b, n_frames, st);
    if (rc != IPP_OK) return rc;
    { ipp::LaunchTimer _lt("ccl_info_kernel", st); ccl_info_kernel<<<ceil_div(n_frames, 128), 128, 0, st>>>(b, info_out, n_frames); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

int ipp_chain_stage_b1(const ipp_chain_geom* g, const ipp_chain_rot* rots, int n_rot, int total_tiles,
                       const void* workspace, int32_t* bbox_out, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = chain_check_geom(g);
    if (rc != IPP_OK) return rc;
    IPP_REQUIRE(rots && workspace && bbox_out && n_rot >= 0 && total_tiles >= 0, IPP_ERR_BAD_ARG, "ipp_chain_stage_b1: bad arguments");
    if (n_rot == 0) return IPP_OK;
    const CclGeom cg = make_ccl_geom(g->crop_w, g->crop_h);
    { ipp::LaunchTimer _lt("bbox_reset_kernel", st); bbox_reset_kernel<<<ceil_div(n_rot, 256), 256, 0, st>>>(bbox_out, n_rot); }
    IPP_LAUNCH_CHECK();
    if (total_tiles == 0) return IPP_OK;
    { ipp::LaunchTimer _lt("chain_rot_bbox_kernel", st); chain_rot_bbox_kernel<<<total_tiles, 256, 0, st>>>(rots, n_rot, (const uint32_t*)workspace, cg.bits_stride, cg.wpr, cg.h, bbox_out); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

int ipp_chain_stage_b2(const uint8_t* frames, int64_t frame_pitch, int64_t frame_stride,
                       const ipp_chain_geom* g, const ipp_chain_item* items, int n_items,
                       int gather_tiles, int h_tiles, int v_tiles,
                       const int32_t* tables, void* band, int max_band_warps, int fused_smem_bytes, int vstaged_smem_bytes,
                       const void* workspace, uint8_t* sprite_ws, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = chain_check_geom(g);
    if (rc != IPP_OK) return rc;
    IPP_REQUIRE(frames && items && workspace && sprite_ws && tables && band && n_items >= 0 && max_band_warps >= 0 && fused_smem_bytes >= 0 &&
                    vstaged_smem_bytes >= 0 && vstaged_smem_bytes <= 200 * 1024,
                IPP_ERR_BAD_ARG, "ipp_chain_stage_b2: bad arguments");
    IPP_REQUIRE((((uintptr_t)sprite_ws) & 255) == 0 && (((uintptr_t)band) & 15) == 0, IPP_ERR_BAD_ARG,
                "ipp_chain_stage_b2: sprite workspace must be 256-byte aligned, fragment tables 16-byte aligned");
    IPP_REQUIRE(fused_smem_bytes <= 200 * 1024 && (fused_smem_bytes == 0 || h_tiles == 0), IPP_ERR_BAD_ARG,
                "ipp_chain_stage_b2: fused mode takes its tiles in gather_tiles (h_tiles = 0) and at most 200 KB of shared memory");
    if (n_items == 0) return IPP_OK;
    const CclGeom cg = make_ccl_geom(g->crop_w, g->crop_h);
    B2Common c{};
    c.frames = frames; c.pitch = frame_pitch; c.stride = frame_stride;
    c.crop_l = g->crop_l; c.crop_t = g->crop_t;
    c.bits = (const uint32_t*)workspace; c.bits_stride = cg.bits_stride; c.wpr = cg.wpr;
    c.items = items; c.n_items = n_items; c.tables = tables; c.band = (uint2*)band; c.sprite_ws = sprite_ws;
    c.h_blocks32 = fused_smem_bytes > 0; c.v_blocks32 = vstaged_smem_bytes > 0;
    if (max_band_warps > 0 && (gather_tiles > 0 || h_tiles > 0 || v_tiles > 0)) {
        dim3 grid(ceil_div(max_band_warps, 8), n_items);
        { ipp::LaunchTimer _lt("chain_band_kernel", st); chain_band_kernel<<<grid, 256, 0, st>>>(c); }
        IPP_LAUNCH_CHECK();
    }
    if (fused_smem_bytes > 0) {
        if (gather_tiles > 0) {
            static int smem_set[64] = {};     // per device; the attribute only ever grows
            int dev_id = 0;
            IPP_CUDA(cudaGetDevice(&dev_id));
            if (dev_id < 0 || dev_id >= 64 || fused_smem_bytes > smem_set[dev_id]) {
                IPP_CUDA(cudaFuncSetAttribute(chain_gh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fused_smem_bytes));
                if (dev_id >= 0 && dev_id < 64) smem_set[dev_id] = fused_smem_bytes;
            }
            { ipp::LaunchTimer _lt("chain_gh_kernel", st); chain_gh_kernel<<<gather_tiles, 256, fused_smem_bytes, st>>>(c); }
            IPP_LAUNCH_CHECK();
        }
    } else {
        if (gather_tiles > 0) {
            { ipp::LaunchTimer _lt("chain_gather_kernel", st); chain_gather_kernel<<<gather_tiles, 256, 0, st>>>(c); }
            IPP_LAUNCH_CHECK();
        }
        if (h_tiles > 0) {
            { ipp::LaunchTimer _lt("chain_hpass_kernel", st); chain_hpass_kernel<<<h_tiles, 256, 0, st>>>(c); }
            IPP_LAUNCH_CHECK();
        }
    }
    if (v_tiles > 0 && vstaged_smem_bytes > 0) {
        static int vsmem_set[64] = {};     // per device; the attribute only ever grows
        int dev_id = 0;
        IPP_CUDA(cudaGetDevice(&dev_id));
        if (dev_id < 0 || dev_id >= 64 || vstaged_smem_bytes > vsmem_set[dev_id]) {
            IPP_CUDA(cudaFuncSetAttribute(chain_vs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vstaged_smem_bytes));
            if (dev_id >= 0 && dev_id < 64) vsmem_set[dev_id] = vstaged_smem_bytes;
        }
        { ipp::LaunchTimer _lt("chain_vs_kernel", st); chain_vs_kernel<<<v_tiles, 256, vstaged_smem_bytes, st>>>(c); }
        IPP_LAUNCH_CHECK();
    } else if (v_tiles > 0) {
        { ipp::LaunchTimer _lt("chain_vpass_kernel", st); chain_vpass_kernel<<<v_tiles, 256, 0, st>>>(c); }
        IPP_LAUNCH_CHECK();
    }
    return IPP_OK;
}

int ipp_chain_tables(const ipp_chain_item* items, int n_items, int max_rows_per_item, int32_t* tables,
                     int32_t* amb_out, int amb_capacity, double eps, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    IPP_REQUIRE(items && tables && amb_out && n_items >= 0 && amb_capacity >= 0 && eps >= 0.0, IPP_ERR_BAD_ARG,
                "ipp_chain_tables: bad arguments");
    IPP_CUDA(cudaMemsetAsync(amb_out, 0, sizeof(int32_t), st));
    if (n_items == 0 || max_rows_per_item <= 0) return IPP_OK;
    TabParams p{items, tables, amb_out, amb_capacity, eps};
    dim3 grid(ceil_div(max_rows_per_item, 128), n_items);
    { ipp::LaunchTimer _lt("chain_tables_kernel", st); chain_tables_kernel<<<grid, 128, 0, st>>>(p); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

static int stage_c_launch(const uint8_t* backgrounds, int64_t bg_pitch, int64_t bg_stride, const ipp_chain_item* items, int n_items,
                          const uint8_t* sprite_ws, uint8_t* out, int64_t out_pitch, int64_t out_stride, int out_w, int out_h,
                          cudaStream_t st) {
    using namespace ipp;
    CParams c{};
    c.bgs = backgrounds; c.bg_pitch = bg_pitch; c.bg_stride = bg_stride;
    c.items = items; c.sprite_ws = sprite_ws;
    c.out = out; c.out_pitch = out_pitch; c.out_stride = out_stride; c.out_w = out_w; c.out_h = out_h;
    const uint32_t rowbytes = (uint32_t)out_w * 3u;
    c.smem_row_stride = staged_row_stride(rowbytes);
    IPP_REQUIRE(c.smem_row_stride <= 100 * 1024, IPP_ERR_UNSUPPORTED, "stage C: row too long");
    int rows = (int)(32768 / c.smem_row_stride);            // one warp per row: 8 rows when they fit
    if (rows > kPasteThreads / 32) rows = kPasteThreads / 32;
    if (rows < 1) rows = 1;
    if (rows > out_h) rows = out_h;
    c.rows_per_cta = rows;
    static const int chunks_env = getenv("IPP_PASTE_CHUNKS") ? atoi(getenv("IPP_PASTE_CHUNKS")) : 0;
    // two chunks per CTA (measured best: 1 / 2 / 4 / 8 chunks = 0.66 / 0.58 / 0.60 / 0.61 ms per 512 images) once that still leaves
    // ~8 waves of CTAs; small batches keep short CTAs (tail effect)
    const int64_t chunk_ctas = (int64_t)ceil_div(out_h, rows) * n_items;
    c.chunks_per_cta = chunks_env > 0 ? chunks_env : (int)std::max<int64_t>(1, std::min<int64_t>(2, chunk_ctas / ((int64_t)sm_count() * 32)));
    const size_t smem = 2 * (size_t)rows * c.smem_row_stride;          // two chunk buffers
    IPP_REQUIRE(smem <= 200 * 1024, IPP_ERR_UNSUPPORTED, "stage C: row too long");
    if (smem > 48 * 1024) {
        static size_t smem_set[64] = {};          // per device; the attribute only ever grows
        int dev_id = 0;
        IPP_CUDA(cudaGetDevice(&dev_id));
        if (dev_id < 0 || dev_id >= 64 || smem > smem_set[dev_id]) {
            IPP_CUDA(cudaFuncSetAttribute(chain_paste_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev_id >= 0 && dev_id < 64) smem_set[dev_id] = smem;
        }
    }
    dim3 grid(ceil_div(out_h, rows * c.chunks_per_cta), n_items);
    { ipp::LaunchTimer _lt("chain_paste_kernel", st); chain_paste_kernel<<<grid, kPasteThreads, smem, st>>>(c); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

int ipp_chain_stage_c(const uint8_t* backgrounds, int64_t bg_pitch, int64_t bg_stride,
                      const ipp_chain_geom* g, const ipp_chain_item* items, int n_items,
                      const uint8_t* sprite_ws,
                      uint8_t* out, int64_t out_pitch, int64_t out_stride, int out_size, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(g && backgrounds && items && sprite_ws && out && n_items >= 0, IPP_ERR_BAD_ARG, "ipp_chain_stage_c: bad arguments");
    IPP_REQUIRE(out_size > 0 && out_size <= g->bg_w && out_size <= g->bg_h && bg_pitch >= (int64_t)g->bg_w * 3 &&
                    out_pitch >= (int64_t)out_size * 3,
                IPP_ERR_BAD_SHAPE, "ipp_chain_stage_c: bad geometry");
    if (n_items == 0) return IPP_OK;
    return stage_c_launch(backgrounds, bg_pitch, bg_stride, items, n_items, sprite_ws, out, out_pitch, out_stride, out_size, out_size,
                          (cudaStream_t)stream);
}

int ipp_chain_stage_c_rect(const uint8_t* backgrounds, int64_t bg_pitch, int64_t bg_stride, int bg_w, int bg_h,
                           const ipp_chain_item* items, int n_items, const uint8_t* sprite_ws,
                           uint8_t* out, int64_t out_pitch, int64_t out_stride, int out_w, int out_h, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(backgrounds && items && sprite_ws && out && n_items >= 0, IPP_ERR_BAD_ARG, "ipp_chain_stage_c_rect: bad arguments");
    IPP_REQUIRE(out_w > 0 && out_h > 0 && out_w <= bg_w && out_h <= bg_h && bg_pitch >= (int64_t)bg_w * 3 && out_pitch >= (int64_t)out_w * 3,
                IPP_ERR_BAD_SHAPE, "ipp_chain_stage_c_rect: bad geometry");
    if (n_items == 0) return IPP_OK;
    return stage_c_launch(backgrounds, bg_pitch, bg_stride, items, n_items, sprite_ws, out, out_pitch, out_stride, out_w, out_h,
                          (cudaStream_t)stream);
}

int ipp_chain_upload_windows(const uint8_t* bgs_host, int64_t bg_pitch, int64_t bg_stride, int bg_w, int bg_h, int n_bgs,
                             const ipp_chain_item* items_host, int n_items, int out_size,
                             uint8_t* windows_dev, void* stream) {
    using namespace ipp;
    cudaStream_t st = (cudaStream_t)stream;
    IPP_REQUIRE(bgs_host && items_host && windows_dev && n_items >= 0 && n_bgs >= 1, IPP_ERR_BAD_ARG,
                "ipp_chain_upload_windows: bad arguments");
    IPP_REQUIRE(out_size > 0 && out_size <= bg_w && out_size <= bg_h && bg_pitch >= (int64_t)bg_w * 3, IPP_ERR_BAD_SHAPE,
                "ipp_chain_upload_windows: bad geometry");
    const size_t row = (size_t)out_size * 3;
    for (int i = 0; i < n_items; ++i) {
        const ipp_chain_item& it = items_host[i];
        IPP_REQUIRE(it.bg >= 0 && it.bg < n_bgs && it.sq_size == out_size && it.sq_x0 >= 0 && it.sq_y0 >= 0 &&
                        it.sq_x0 + out_size <= bg_w && it.sq_y0 + out_size <= bg_h,
                    IPP_ERR_BAD_ARG, "ipp_chain_upload_windows: item %d has a window outside its background", i);
    }
    // SM-driven gather: needs rows whose aligned 16-byte cover stays inside the row (pitch and base multiples of 16) and host
    // memory the device can address (pinned; unified addressing)
    // Short window rows are where the copy engines lose (1080p: 3240 of 5760 bytes per row, 36 GB/s next to the frame uploads;
    // gather 4.37 k -> 4.75 k images/s end to end); with 4K rows (6480 of 11520 bytes) they win (1.16 k vs 1.09-1.11 k).
    // IPP_WINDOWS_COPY=dma / sm forces one or the other.
    static const char* mode_env = getenv("IPP_WINDOWS_COPY");
    static const int ctas_env = getenv("IPP_WINDOWS_CTAS") ? atoi(getenv("IPP_WINDOWS_CTAS")) : 32;
    const bool want_sm = mode_env ? mode_env[0] == 's' : row <= 4096;
    const bool sm_copy = want_sm && (bg_pitch % 16) == 0 && (bg_stride % 16) == 0 && (((uintptr_t)bgs_host) & 15) == 0 &&
                         (int64_t)bg_w * 3 <= bg_pitch;
    const uint8_t* bgs_dev_view = nullptr;       // the host buffer as the device addresses it (pageable memory: none -> copy engines)
    if (sm_copy) {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, bgs_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer != nullptr)
            bgs_dev_view = static_cast<const uint8_t*>(pa.devicePointer);
        else
            (void)cudaGetLastError();
    }
    if (sm_copy && bgs_dev_view) {
        const uint32_t stride = staged_row_stride((uint32_t)row);
        const size_t smem = 8 * (size_t)stride;
        IPP_REQUIRE(smem <= 200 * 1024, IPP_ERR_UNSUPPORTED, "ipp_chain_upload_windows: row too long");
        static size_t set_w[64] = {};
        int dev_id = 0;
        IPP_CUDA(cudaGetDevice(&dev_id));
        if (smem > 48 * 1024 && (dev_id < 0 || dev_id >= 64 || smem > set_w[dev_id])) {
            IPP_CUDA(cudaFuncSetAttribute(window_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev_id >= 0 && dev_id < 64) set_w[dev_id] = smem;
        }
        for (int i0 = 0; i0 < n_items; i0 += kWinMaxItems) {
            WinGatherParams wp{};
            wp.n = std::min(kWinMaxItems, n_items - i0);
            for (int i = 0; i < wp.n; ++i) {
                const ipp_chain_item& it = items_host[i0 + i];
                wp.src[i] = bgs_dev_view + (int64_t)it.bg * bg_stride + (int64_t)it.sq_y0 * bg_pitch + (int64_t)it.sq_x0 * 3;
            }
            wp.bg_pitch = bg_pitch; wp.out_size = out_size; wp.smem_row_stride = stride;
            wp.dst = windows_dev + (size_t)i0 * row * out_size;
            const int grid = (int)std::min<int64_t>(ctas_env, ceil_div((int64_t)wp.n * out_size, 8));
            { ipp::LaunchTimer _lt("window_gather_kernel", st); window_gather_kernel<<<grid, 256, smem, st>>>(wp); }
            IPP_LAUNCH_CHECK();
        }
        return IPP_OK;
    }
    for (int i = 0; i < n_items; ++i) {
        const ipp_chain_item& it = items_host[i];
        const uint8_t* src = bgs_host + (int64_t)it.bg * bg_stride + (int64_t)it.sq_y0 * bg_pitch + (int64_t)it.sq_x0 * 3;
        IPP_CUDA(cudaMemcpy2DAsync(windows_dev + (size_t)i * row * out_size, row, src, (size_t)bg_pitch, row, (size_t)out_size,
                                   cudaMemcpyHostToDevice, st));
    }
    return IPP_OK;
}

}  // extern "C"
