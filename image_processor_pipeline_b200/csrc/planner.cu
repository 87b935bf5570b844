// planner.cu — the chain's HOST planner in C++ (no device code): the reference's per-step
// bookkeeping and random draws between the device stages, at native speed.
//
// What it restates (file:line of the reference, or of the Python runtime the reference
// runs on — the draws come from Python's `random`, a Mersenne Twister):
//   * Python random: MT19937 genrand_uint32 / random() (53-bit) / uniform / getrandbits /
//     _randbelow_with_getrandbits / randint / sample (pool branch, n <= 21)
//   * rotation step: transforms/rotations.py:77-119 + Pillow Image.rotate parameter math
//     (PIL/Image.py:2509-2585) + libImaging affine_fixed coefficients (Geometry.c)
//   * symmetry step: transforms/symmetry.py:97-126 (+ sorted file-name order of the
//     next step's directory listing, pipeline.py:137)
//   * overlay step: transforms/overlays.py:104-149 (size math, randint position,
//     xyxy2xywhn float32, "%.6f" label)
//   * square-crop step: transforms/crop_square.py:62-65,163-217 (label parse, float32
//     xywhn2xyxy, bounds, randint offset, shift/clip/valid, xyxy2xywhn, label rows)
// The Python planner in chain.py is the readable twin; tests require both to agree
// (records, labels and the RNG state afterwards).
#include "common.cuh"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static_assert(sizeof(ipp_chain_rot) == 64, "ipp_chain_rot layout (chain.py ROT_DTYPE)");
static_assert(sizeof(ipp_chain_item) == 184, "ipp_chain_item layout (chain.py ITEM_DTYPE)");
static_assert(sizeof(ipp_plan_params) == 88 && sizeof(ipp_plan_aux) == 16 && sizeof(ipp_plan_totals) == 64, "planner struct layout");

namespace ipp {
namespace plan {

// ------------------------------------------------------------------ Python's random.Random
struct PyRandom {
    uint32_t* mt;     // 624 words
    int32_t* idx;     // state index (0..624)

    uint32_t genrand_uint32() {
        static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
        if (*idx >= 624) {
            int kk;
            uint32_t y;
            for (kk = 0; kk < 624 - 397; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < 623; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
            *idx = 0;
        }
        uint32_t y = mt[(*idx)++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double random() {   // genrand_res53
        const uint32_t a = genrand_uint32() >> 5, b = genrand_uint32() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    double uniform(double a, double b) { return a + (b - a) * random(); }
    uint32_t randbelow(uint32_t n) {   // n > 0, n < 2^32
        int k = 0;
        for (uint32_t t = n; t; t >>= 1) ++k;
        uint32_t r = genrand_uint32() >> (32 - k);
        while (r >= n) r = genrand_uint32() >> (32 - k);
        return r;
    }
    // randint(a, b); ok=false when the range is empty (Python raises ValueError)
    long long randint(long long a, long long b, bool* ok) {
        const long long width = b + 1 - a;
        if (width <= 0) {
            *ok = false;
            return 0;
        }
        *ok = true;
        return a + (long long)randbelow((uint32_t)width);
    }
};

// round_half_even(|x| * 10^digits) on the EXACT binary value of x (128-bit integer arithmetic) — what printf("%.*f")
// and Python's round(x, n) compute, without going through a decimal string.  false: outside the fast range.
static bool scaled_round(double x, int digits, uint64_t* out) {
    static const uint64_t P10[16] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
                                     1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
                                     100000000000000ull, 1000000000000000ull};
    if (!(x == x) || digits < 0 || digits > 15) return false;
    const double ax = std::fabs(x);
    if (ax >= 4096.0) return false;
    if (ax == 0.0) { *out = 0; return true; }
    int ex;
    const double f = std::frexp(ax, &ex);                       // ax = f * 2^ex, f in [0.5, 1)
    const uint64_t m = (uint64_t)std::ldexp(f, 53);             // exact: 53-bit integer
    const int sh = 53 - ex;                                     // ax = m * 2^-sh, sh >= 41 here
    const unsigned __int128 N = (unsigned __int128)m * P10[digits];     // < 2^53 * 2^50
    if (sh >= 127) { *out = 0; return true; }                   // ax * 10^digits < 2^-24: rounds to 0
    const unsigned __int128 one = 1;
    unsigned __int128 q = N >> sh;
    const unsigned __int128 r = N & ((one << sh) - 1), half = one << (sh - 1);
    if (r > half || (r == half && (q & 1))) ++q;
    *out = (uint64_t)q;
    return true;
}

// round(x, 15) of Python: exact decimal rounding (half-even on the exact binary value), then the nearest double
// to that decimal = the correctly rounded quotient of two exactly representable numbers
static double round15(double x) {
    uint64_t q;
    if (scaled_round(x, 15, &q) && q < (1ull << 53)) return std::copysign((double)q / 1e15, x);
    char buf[400];
    snprintf(buf, sizeof(buf), "%.15f", x);
    return strtod(buf, nullptr);
}

// "%.6f" of a non-negative value + the float64 np.loadtxt reads back from that text.  Returns the number of
// characters written (no terminator).
static int fmt6(char* dst, double v, double* parsed) {
    uint64_t q;
    if (v >= 0.0 && scaled_round(v, 6, &q)) {
        const uint64_t ip = q / 1000000ull;
        uint32_t fp = (uint32_t)(q % 1000000ull);
        char tmp[24];
        int n = 0;
        uint64_t t = ip;
        do { tmp[n++] = (char)('0' + t % 10); t /= 10; } while (t);
        int len = 0;
        while (n) dst[len++] = tmp[--n];
        dst[len++] = '.';
        for (int i = 5; i >= 0; --i) { dst[len + i] = (char)('0' + fp % 10); fp /= 10; }
        len += 6;
        if (parsed) *parsed = (double)q / 1e6;
        return len;
    }
    const int len = snprintf(dst, 48, "%.6f", v);
    if (parsed) *parsed = strtod(dst, nullptr);
    return len;
}
static double py_fmod360(double a) {
    double m = std::fmod(a, 360.0);
    if (m != 0.0) {
        if (m < 0.0) m += 360.0;
    } else {
        m = std::copysign(0.0, 360.0);
    }
    return m;
}
static int fix16(double v) {
    const double t = v * 65536.0 + 0.5;
    return t >= 0.0 ? (int)t : (int)std::floor(t);
}

// Pillow Image.rotate(angle, expand=True) parameter math for a w x h image
static int rotation_plan(int w, int h, double angle, ipp_chain_rot* r, double* m_out = nullptr) {
    angle = py_fmod360(angle);
    r->a0 = r->a1 = r->a2 = r->a3 = r->a4 = r->a5 = 0;
    if (angle == 0.0) { r->kind = 1; r->canvas_w = w; r->canvas_h = h; return 0; }
    if (angle == 180.0) { r->kind = 2; r->canvas_w = w; r->canvas_h = h; return 0; }
    if (angle == 90.0 || angle == 270.0) { r->kind = angle == 90.0 ? 3 : 4; r->canvas_w = h; r->canvas_h = w; return 0; }
    const double cx = w / 2.0, cy = h / 2.0;
    const double a = -(angle * (M_PI / 180.0));
    const double rc = round15(std::cos(a)), rs = round15(std::sin(a));       // round(-x, 15) == -round(x, 15)
    double m[6] = {rc, rs, 0.0, -rs, rc, 0.0};
    auto tfx = [&](double x, double y) { return m[0] * x + m[1] * y + m[2]; };
    auto tfy = [&](double x, double y) { return m[3] * x + m[4] * y + m[5]; };
    {
        const double nx = tfx(-cx, -cy), ny = tfy(-cx, -cy);
        m[2] = nx;
        m[5] = ny;
    }
    m[2] += cx;
    m[5] += cy;
    const double xs[4] = {0.0, (double)w, (double)w, 0.0}, ys[4] = {0.0, 0.0, (double)h, (double)h};
    double minx = 0, maxx = 0, miny = 0, maxy = 0;
    for (int i = 0; i < 4; ++i) {
        const double tx = tfx(xs[i], ys[i]), ty = tfy(xs[i], ys[i]);
        if (i == 0) { minx = maxx = tx; miny = maxy = ty; }
        else { minx = std::fmin(minx, tx); maxx = std::fmax(maxx, tx); miny = std::fmin(miny, ty); maxy = std::fmax(maxy, ty); }
    }
    const int nw = (int)std::ceil(maxx) - (int)std::floor(minx);
    const int nh = (int)std::ceil(maxy) - (int)std::floor(miny);
    {
        const double nx = tfx(-(nw - w) / 2.0, -(nh - h) / 2.0), ny = tfy(-(nw - w) / 2.0, -(nh - h) / 2.0);
        m[2] = nx;
        m[5] = ny;
    }
    auto in_range = [&](double x, double y) {
        return std::fabs(x * m[0] + y * m[1] + m[2]) < 32768.0 && std::fabs(x * m[3] + y * m[4] + m[5]) < 32768.0;
    };
    if (!(in_range(0, 0) && in_range(nw, nh) && in_range(0, nh) && in_range(nw, 0)) || (m[1] == 0.0 && m[3] == 0.0))
        return IPP_ERR_UNSUPPORTED;      // outside libImaging's affine_fixed path
    r->kind = 0;
    if (m_out) for (int i = 0; i < 6; ++i) m_out[i] = m[i];
    r->a0 = fix16(m[0]); r->a1 = fix16(m[1]); r->a2 = fix16(m[2] + m[0] * 0.5 + m[1] * 0.5);
    r->a3 = fix16(m[3]); r->a4 = fix16(m[4]); r->a5 = fix16(m[5] + m[3] * 0.5 + m[4] * 0.5);
    r->canvas_w = nw; r->canvas_h = nh;
    return 0;
}

// math.hypot of CPython >= 3.10 for two finite arguments (Modules/mathmodule.c vector_norm): scaling by a power of two,
// lossless squaring (fma) and addition (two-sum), a square root and one differential correction step.  It is NOT the C
// library's hypot(): the two differ in the last bit now and then, and overlays.py:125 truncates a value derived from it with
// int() — for a sprite that is limited by the background's height the expression is mathematically an integer, so the last
// bit decides between h and h - 1.  (Found by tests/fuzz_parity.py: seed 705192.)
static double py_hypot2(double a, double b) {
    a = std::fabs(a);
    b = std::fabs(b);
    const double mx = a > b ? a : b;
    if (std::isinf(mx)) return mx;
    if (a != a || b != b) return NAN;
    if (mx == 0.0) return mx;
    int max_e;
    std::frexp(mx, &max_e);
    if (max_e < -1023) return DBL_MIN * py_hypot2(a / DBL_MIN, b / DBL_MIN);
    const double scale = std::ldexp(1.0, -max_e);
    double csum = 1.0, frac1 = 0.0, frac2 = 0.0;
    const double v[2] = {a, b};
    for (int i = 0; i < 2; ++i) {
        const double x = v[i] * scale;                       // lossless scaling
        const double hi = x * x, lo = std::fma(x, x, -hi);   // lossless squaring
        const double sx = csum + hi, sy = (csum - sx) + hi;  // lossless addition (|csum| >= |hi|)
        csum = sx;
        frac1 += lo;
        frac2 += sy;
    }
    double h = std::sqrt(csum - 1.0 + (frac1 + frac2));
    {
        const double hi = -h * h, lo = std::fma(-h, h, -hi);
        const double sx = csum + hi, sy = (csum - sx) + hi;
        csum = sx;
        frac1 += lo;
        frac2 += sy;
    }
    const double x = csum - 1.0 + (frac1 + frac2);
    h += x / (2.0 * h);                                      // differential correction
    return h / scale;
}

// reference transforms/overlays.py:106-126; false when PIL's resize would raise (size <= 0)
static bool overlay_size(int ow, int oh, int bw, int bh, double ratio, int* nw, int* nh) {
    if (oh == 0) return false;
    const double bg_diag = py_hypot2((double)bw, (double)bh);
    const double target = bg_diag * ratio;
    const double ar = (double)ow / (double)oh;
    const double h_max = std::fmin((double)bw / ar, (double)bh);
    const double max_diag = py_hypot2(ar * h_max, h_max);
    const double d = std::fmin(target, max_diag);
    // Python's `x ** 2` is libm's pow(x, 2.0), which is NOT always x * x (glibc's pow is within 1 ulp, not correctly rounded:
    // 1968.445400028335 ** 2 differs from the product in the last bit, and int() of the root then gives 1967 instead of 1968).
    // The call goes through a volatile pointer so that the compiler cannot fold it into a multiplication.
    static double (*volatile libm_pow)(double, double) = static_cast<double (*)(double, double)>(::pow);
    const double hh = std::sqrt(libm_pow(d, 2.0) / (libm_pow(ar, 2.0) + 1.0));
    if (!(hh == hh)) return false;
    *nh = (int)hh;
    *nw = (int)(ar * (double)*nh);
    return *nw > 0 && *nh > 0;
}

static int lanczos_ksize(int in_size, int out_size) {
    double fs = (double)((float)in_size - 0.0f) / out_size;
    if (fs < 1.0) fs = 1.0;
    return (int)std::ceil(3.0 * fs) * 2 + 1;
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace plan
}  // namespace ipp

extern "C" {

// Rotation step.  info: stage-A rows int32[n_frames][8] (first_frame = global index of row 0,
// for sharded runs pass the gathered rows and 0).  Outputs one ipp_chain_rot per sprite
// (original first when include_original, then the rotations), angle (NaN for the original)
// and rot_index (0 = original).  Returns IPP_ERR_WORKSPACE when capacity is too small.
int ipp_plan_rotations_host(const int32_t* info, int n_frames, const ipp_plan_params* p, uint32_t* rng_mt624, int32_t* rng_index,
                            ipp_chain_rot* rots_out, double* angles_out, int32_t* rot_index_out, int capacity,
                            int32_t* n_rot_out, int32_t* total_tiles_out) {
    using namespace ipp;
    using namespace ipp::plan;
    IPP_REQUIRE(info && p && rng_mt624 && rng_index && rots_out && angles_out && rot_index_out && n_rot_out && total_tiles_out,
                IPP_ERR_BAD_ARG, "ipp_plan_rotations_host: NULL pointer");
    PyRandom rng{rng_mt624, rng_index};
    int n = 0;
    long long tiles = 0;
    for (int f = 0; f < n_frames; ++f) {
        const int32_t* row = info + 8 * f;
        const int w = row[2], h = row[3];
        if (w <= 0 || h <= 0) continue;
        for (int i = (p->include_original ? 0 : 1); i <= p->num_rotations; ++i) {
            IPP_REQUIRE(n < capacity, IPP_ERR_WORKSPACE, "ipp_plan_rotations_host: capacity %d too small", capacity);
            ipp_chain_rot* r = rots_out + n;
            memset(r, 0, sizeof(*r));
            r->frame = f; r->cc_x = row[0]; r->cc_y = row[1]; r->cc_w = w; r->cc_h = h;
            if (i == 0) {
                r->kind = 1; r->canvas_w = w; r->canvas_h = h;
                angles_out[n] = NAN;
            } else {
                const double angle = rng.uniform(p->angle_min, p->angle_max);
                const int rc = rotation_plan(w, h, angle, r);
                IPP_REQUIRE(rc == 0, rc, "ipp_plan_rotations_host: rotation outside libImaging's 16.16 fixed-point range");
                angles_out[n] = angle;
            }
            rot_index_out[n] = i;
            r->tile_start = (int32_t)tiles;
            {   // stage B1 tile = 256 aligned quads of mask words of the component rect (chain_rot_bbox_kernel)
                const long long nq = ((((r->cc_x + r->cc_w + 31) >> 5) + 3) >> 2) - ((r->cc_x >> 5) >> 2);
                tiles += (nq * r->cc_h + 255) / 256;
            }
            IPP_REQUIRE(tiles < (1LL << 31), IPP_ERR_UNSUPPORTED, "ipp_plan_rotations_host: too many tiles");
            ++n;
        }
    }
    *n_rot_out = n;
    *total_tiles_out = (int32_t)tiles;
    return IPP_OK;
}

// Device-side schedule of an item array: coefficient-table offsets, banded int8 fragment tables for the
// tensor-core LANCZOS passes, flat tile starts, sprite-workspace offsets.  Shared by the planner and by
// callers that build item arrays themselves (tests).
static int layout_impl(ipp_chain_item* items, int n_items, ipp_plan_totals* totals_out, bool allow_fused);
int ipp_chain_layout_host(ipp_chain_item* items, int n_items, ipp_plan_totals* totals_out) {
    return layout_impl(items, n_items, totals_out, true);
}
int ipp_chain_layout_planar_host(ipp_chain_item* items, int n_items, ipp_plan_totals* totals_out) {
    return layout_impl(items, n_items, totals_out, false);
}
static int layout_impl(ipp_chain_item* items, int n_items, ipp_plan_totals* totals_out, bool allow_fused) {
    using namespace ipp;
    using namespace ipp::plan;
    IPP_REQUIRE((items || n_items == 0) && n_items >= 0 && totals_out, IPP_ERR_BAD_ARG, "ipp_chain_layout_host: bad arguments");
    int64_t toff = 0, tab_rows = 0, woff = 0, boff = 0;
    long long gt = 0, ht = 0, vt = 0;
    int max_rows = 0, max_band_warps = 0;
    for (int i = 0; i < n_items; ++i) {
        ipp_chain_item* it = items + i;
        const int sw = it->spr_w, sh = it->spr_h, nw = it->new_w, nh = it->new_h;
        IPP_REQUIRE(sw > 0 && sh > 0 && nw > 0 && nh > 0, IPP_ERR_BAD_SHAPE, "ipp_chain_layout_host: item %d has an empty sprite", i);
        it->h_ksize = it->v_ksize = 0;
        it->hb_off = it->hk_off = it->vb_off = it->vk_off = 0;
        it->tab_row_start = (int32_t)tab_rows;
        int rows = 0;
        if (nw != sw) {
            it->h_ksize = lanczos_ksize(sw, nw);
            it->hb_off = (int32_t)toff; it->hk_off = (int32_t)(toff + 2 * (int64_t)nw);
            toff += 2 * (int64_t)nw + (int64_t)nw * it->h_ksize;
            rows += nw;
        }
        if (nh != sh) {
            it->v_ksize = lanczos_ksize(sh, nh);
            it->vb_off = (int32_t)toff; it->vk_off = (int32_t)(toff + 2 * (int64_t)nh);
            toff += 2 * (int64_t)nh + (int64_t)nh * it->v_ksize;
            rows += nh;
        }
        tab_rows += rows;
        max_rows = std::max(max_rows, rows);
        IPP_REQUIRE(toff < (1LL << 31) && tab_rows < (1LL << 31), IPP_ERR_UNSUPPORTED, "ipp_chain_layout_host: coefficient tables exceed 2^31 entries");
        // banded fragment tables: an 8-output tile reads at most ceil(7 * in / out) + ksize + 16 source positions
        // (window drift over 7 outputs + one window + alignment of the first position to 16), in super-blocks of 64
        auto sblocks = [](int in, int out, int ks) { return (int)(((7LL * in + out - 1) / out + std::max(ks, 1) + 16 + 63) / 64); };
        const int nsb_h = sblocks(sw, nw, it->h_ksize), nsb_v = sblocks(sh, nh, it->v_ksize);
        const int nt_h = (nw + 7) / 8, nt_v = (nh + 7) / 8;
        it->nkb = nsb_h | (nsb_v << 16);
        it->hband_off = (int32_t)boff; boff += (nt_h + 1) / 2 + (int64_t)nt_h * nsb_h * 192;
        it->vband_off = (int32_t)boff; boff += (nt_v + 1) / 2 + (int64_t)nt_v * nsb_v * 192;
        IPP_REQUIRE(boff < (1LL << 31), IPP_ERR_UNSUPPORTED, "ipp_chain_layout_host: fragment tables exceed 16 GiB; use a smaller batch");
        max_band_warps = std::max(max_band_warps, nt_h * nsb_h + nt_v * nsb_v);
        it->gather_tile_start = (int32_t)gt; it->h_tile_start = (int32_t)ht; it->v_tile_start = (int32_t)vt;
        gt += (long long)((sw + 31) / 32) * ((sh + 63) / 64);          // chain_gather_kernel: 32 x 64 tiles
        ht += (long long)((nw + 63) / 64) * 4;
        vt += (long long)((nh + 63) / 64) * ((nw + 127) / 128);
        IPP_REQUIRE(gt < (1LL << 31) && ht < (1LL << 31) && vt < (1LL << 31), IPP_ERR_UNSUPPORTED, "ipp_chain_layout_host: too many tiles");
        // planar premultiplied sprite [4][sh][Wp], transposed planar h-pass result [4][nw][Hp], final RGBA sprite
        const int64_t Wp = (sw + 15) & ~15, Hp = (sh + 15) & ~15;
        // + 16 rows and 1 KB: the LANCZOS passes read whole 16-row tiles and 64-byte windows without clamping
        it->rot_off = woff; woff = align_up(woff + (4 * (int64_t)sh + 16) * Wp + 1024, 256);
        it->tmp_off = woff; woff = align_up(woff + (4 * (int64_t)nw + 16) * Hp + 1024, 256);
        it->spr_off = woff; woff = align_up(woff + (int64_t)nw * nh * 4, 256);
        it->out_slot = i;
    }
    // gather + h pass fused (16 gathered rows x 4 planes in shared memory) when the widest sprite fits; then the
    // gather tile list is "one tile per 16 sprite rows" and there are no h tiles
    int max_w = 0;
    for (int i = 0; i < n_items; ++i) max_w = std::max(max_w, items[i].spr_w);
    const int pitch = ((max_w + 64 + 127) / 128) * 128 + 64;              // gh_row_pitch (chain.cu)
    const int64_t fused_smem = n_items ? (int64_t)64 * pitch + 1024 : 0;   // 4 planes x 16 rows + tail slack
    const bool fused = allow_fused && n_items > 0 && fused_smem <= 200 * 1024 && getenv("IPP_B2_UNFUSED") == nullptr;
    if (fused) {
        gt = 0;
        ht = 0;
        for (int i = 0; i < n_items; ++i) {
            items[i].gather_tile_start = (int32_t)gt;
            items[i].h_tile_start = 0;
            gt += (items[i].spr_h + 15) / 16;
        }
    }
    totals_out->fused_smem_bytes = fused ? (int32_t)fused_smem : 0;
    // vertical pass staged per CTA (16 columns x 4 planes of the transposed intermediate) when the tallest sprite fits
    int max_h = 0;
    for (int i = 0; i < n_items; ++i) max_h = std::max(max_h, items[i].spr_h);
    const int vpitch = ((((max_h + 15) & ~15) + 32 - 64 + 127) / 128) * 128 + 64;   // vs_row_pitch (chain.cu)
    const int64_t vs_smem = n_items ? (int64_t)64 * vpitch + 1024 : 0;
    const bool vstaged = n_items > 0 && vs_smem <= 200 * 1024 && getenv("IPP_B2_UNFUSED") == nullptr;
    if (vstaged) {
        vt = 0;
        for (int i = 0; i < n_items; ++i) {
            items[i].v_tile_start = (int32_t)vt;
            vt += (items[i].new_w + 15) / 16;
        }
    }
    totals_out->vstaged_smem_bytes = vstaged ? (int32_t)vs_smem : 0;
    totals_out->table_ints = toff;
    totals_out->table_rows = tab_rows;
    totals_out->sprite_ws_bytes = woff;
    totals_out->band_units = boff;
    totals_out->gather_tiles = (int32_t)gt;
    totals_out->h_tiles = (int32_t)ht;
    totals_out->v_tiles = (int32_t)vt;
    totals_out->max_rows = max_rows;
    totals_out->max_band_warps = max_band_warps;
    return IPP_OK;
}

// Steps after the rotation: fit, symmetry, overlay, square crop.  bbox: stage-B1 rows
// int32[n_rot][4].  Fills item records (frame, bg, geometry, flips, sizes, positions, crop,
// table offsets, tile starts, workspace offsets), aux rows {rot (index into rots), sym
// (0 o,1 h,2 v,3 hv), ratio} and the label text of every item (labels_out, '\0'-separated).
int ipp_plan_items_host(const ipp_chain_rot* rots, const int32_t* rot_is_original, int n_rot, const int32_t* bbox,
                        const ipp_plan_params* p, uint32_t* rng_mt624, int32_t* rng_index,
                        ipp_chain_item* items_out, ipp_plan_aux* aux_out, int capacity, int32_t* n_items_out,
                        char* labels_out, int labels_capacity, ipp_plan_totals* totals_out) {
    using namespace ipp;
    using namespace ipp::plan;
    IPP_REQUIRE(rots && rot_is_original && bbox && p && rng_mt624 && rng_index && items_out && aux_out && n_items_out &&
                    labels_out && totals_out,
                IPP_ERR_BAD_ARG, "ipp_plan_items_host: NULL pointer");
    IPP_REQUIRE(p->pool_len >= 1 && p->pool_len <= 4 && p->choose_random >= 0 && p->choose_random <= p->pool_len && p->n_bgs >= 1,
                IPP_ERR_BAD_ARG, "ipp_plan_items_host: bad symmetry pool / choose_random / background count");
    PyRandom rng{rng_mt624, rng_index};
    const int BW = p->bg_w, BH = p->bg_h;
    static const int sym_file_order[4] = {2, 0, 3, 1};   // o, h, v, hv -> rank of "_o" < ... in file-name order: h < hv < o < v

    struct Var { int rot, sym, fx, fy, sw, sh; };
    std::vector<Var> variants;
    variants.reserve((size_t)n_rot * 2);
    // ---- symmetry step: one sample per sprite (symmetry.py:122-126), then sorted by file name
    for (int r = 0; r < n_rot; ++r) {
        const ipp_chain_rot& R = rots[r];
        const int32_t* bb = bbox + 4 * r;
        int fx = 0, fy = 0, sw = R.canvas_w, sh = R.canvas_h;
        if (!rot_is_original[r] && bb[2] > 0) { fx = bb[0]; fy = bb[1]; sw = bb[2] - bb[0]; sh = bb[3] - bb[1]; }
        int pool[4], n = p->pool_len, chosen[5], nc = 0;
        for (int i = 0; i < n; ++i) pool[i] = p->pool[i];
        for (int i = 0; i < p->choose_random; ++i) {
            const int j = (int)rng.randbelow((uint32_t)(n - i));
            chosen[nc++] = pool[j];
            pool[j] = pool[n - i - 1];
        }
        bool has_o = false;
        for (int i = 0; i < nc; ++i) has_o |= (chosen[i] == 0);
        if (p->sym_include_original && !has_o) chosen[nc++] = 0;
        bool present[4] = {false, false, false, false};
        for (int i = 0; i < nc; ++i) present[chosen[i]] = true;
        static const int by_name[4] = {1, 3, 0, 2};      // h, hv, o, v
        (void)sym_file_order;
        for (int k = 0; k < 4; ++k)
            if (present[by_name[k]]) variants.push_back({r, by_name[k], fx, fy, sw, sh});
    }
    // ---- overlay step (overlays.py:104-149)
    struct Ov { Var v; int bg, nw, nh, px, py; double ratio; double box[4]; };   // box: the label's numbers as np.loadtxt reads them
    std::vector<Ov> overlays;
    overlays.reserve(variants.size());
    for (size_t k = 0; k < variants.size(); ++k) {
        const Var& v = variants[k];
        Ov o;
        o.v = v;
        o.bg = (int)(k % (size_t)p->n_bgs);
        o.ratio = rng.uniform(p->scale_min, p->scale_max);
        if (!overlay_size(v.sw, v.sh, BW, BH, o.ratio, &o.nw, &o.nh)) continue;   // ValueError -> print, None
        bool ok;
        o.px = (int)rng.randint(0, BW - o.nw, &ok);
        if (!ok) continue;
        o.py = (int)rng.randint(0, BH - o.nh, &ok);
        if (!ok) continue;
        const long long x1 = o.px, y1 = o.py, x2 = o.px + o.nw, y2 = o.py + o.nh;
        const float cx = (float)(((double)(x1 + x2) / 2.0) / (double)BW), cy = (float)(((double)(y1 + y2) / 2.0) / (double)BH);
        const float wn = (float)((double)(x2 - x1) / (double)BW), hn = (float)((double)(y2 - y1) / (double)BH);
        {   // label text "{cls} {cx:.6f} {cy:.6f} {w:.6f} {h:.6f}" (overlays.py:149): only its parsed-back numbers are needed
            char scratch[48];
            fmt6(scratch, (double)cx, &o.box[0]); fmt6(scratch, (double)cy, &o.box[1]);
            fmt6(scratch, (double)wn, &o.box[2]); fmt6(scratch, (double)hn, &o.box[3]);
        }
        overlays.push_back(o);
    }
    // ---- square-crop step (crop_square.py:62-65,163-217) + records
    int n = 0;
    size_t lpos = 0;
    const int cs = BW < BH ? BW : BH;
    for (const Ov& o : overlays) {
        // np.loadtxt: class id and the four 6-decimal numbers back to float64
        const int cls = p->yolo_class_id;      // data[:, 0].astype(int)
        const double* b = o.box;
        // xywhn2xyxy: float64 expression, float32 store
        const float ax1 = (float)((double)BW * (b[0] - b[2] / 2.0)), ay1 = (float)((double)BH * (b[1] - b[3] / 2.0));
        const float ax2 = (float)((double)BW * (b[0] + b[2] / 2.0)), ay2 = (float)((double)BH * (b[1] + b[3] / 2.0));
        const int lbx = std::max(0, (int)(ax2 - (float)cs)), ubx = std::min((int)ax1, BW - cs);
        const int lby = std::max(0, (int)(ay2 - (float)cs)), uby = std::min((int)ay1, BH - cs);
        if (lbx > ubx || lby > uby) continue;                     // RuntimeError in the reference -> item logged as Error
        bool ok;
        const int x0 = (int)rng.randint(lbx, ubx, &ok);
        const int y0 = (int)rng.randint(lby, uby, &ok);
        auto clip = [&](double v) { return std::fmin(std::fmax(v, 0.0), (double)cs); };
        const double x1 = clip((double)ax1 - (double)x0), y1 = clip((double)ay1 - (double)y0);
        const double x2 = clip((double)ax2 - (double)x0), y2 = clip((double)ay2 - (double)y0);
        if (!(x1 < x2 && y1 < y2)) continue;                       // "Aucune bbox résiduelle après le crop."
        const float ncx = (float)(((x1 + x2) / 2.0) / (double)cs), ncy = (float)(((y1 + y2) / 2.0) / (double)cs);
        const float nwn = (float)((x2 - x1) / (double)cs), nhn = (float)((y2 - y1) / (double)cs);

        IPP_REQUIRE(n < capacity, IPP_ERR_WORKSPACE, "ipp_plan_items_host: capacity %d too small", capacity);
        char row[256];
        int len = snprintf(row, 16, "%d", cls);
        const float vals[4] = {ncx, ncy, nwn, nhn};
        for (int i = 0; i < 4; ++i) {
            row[len++] = ' ';
            len += fmt6(row + len, (double)vals[i], nullptr);
        }
        row[len++] = '\n';
        row[len] = 0;
        IPP_REQUIRE(lpos + (size_t)len + 1 <= (size_t)labels_capacity, IPP_ERR_WORKSPACE, "ipp_plan_items_host: label buffer too small");
        memcpy(labels_out + lpos, row, (size_t)len + 1);
        lpos += (size_t)len + 1;

        const ipp_chain_rot& R = rots[o.v.rot];
        ipp_chain_item* it = items_out + n;
        memset(it, 0, sizeof(*it));
        it->frame = R.frame; it->bg = o.bg;
        it->cc_x = R.cc_x; it->cc_y = R.cc_y; it->cc_w = R.cc_w; it->cc_h = R.cc_h;
        it->kind = R.kind; it->a0 = R.a0; it->a1 = R.a1; it->a2 = R.a2; it->a3 = R.a3; it->a4 = R.a4; it->a5 = R.a5;
        it->fit_x0 = o.v.fx; it->fit_y0 = o.v.fy; it->spr_w = o.v.sw; it->spr_h = o.v.sh;
        it->flip_h = (o.v.sym == 1 || o.v.sym == 3); it->flip_v = (o.v.sym == 2 || o.v.sym == 3);
        it->new_w = o.nw; it->new_h = o.nh; it->pos_x = o.px; it->pos_y = o.py;
        it->sq_x0 = x0; it->sq_y0 = y0; it->sq_size = cs;
        aux_out[n].ratio = o.ratio;
        aux_out[n].rot = o.v.rot;
        aux_out[n].sym = o.v.sym;
        ++n;
    }
    *n_items_out = n;
    const int lrc = ipp_chain_layout_host(items_out, n, totals_out);
    if (lrc != IPP_OK) return lrc;
    totals_out->labels_bytes = (int32_t)lpos;
    return IPP_OK;
}

// Batched rotation step (rotations.py:88-99) over n images of one size: one uniform() per image.
int ipp_plan_rotate_batch_host(int src_w, int src_h, int n, double angle_min, double angle_max,
                               uint32_t* rng_mt624, int32_t* rng_index,
                               ipp_rot_item* items_out, double* angles_out, int64_t* totals_out) {
    using namespace ipp;
    using namespace ipp::plan;
    IPP_REQUIRE(src_w > 0 && src_h > 0 && n >= 0 && rng_mt624 && rng_index && items_out && angles_out && totals_out, IPP_ERR_BAD_ARG,
                "ipp_plan_rotate_batch_host: bad arguments");
    PyRandom rng{rng_mt624, rng_index};
    int64_t off = 0, tiles = 0;
    for (int i = 0; i < n; ++i) {
        const double angle = rng.uniform(angle_min, angle_max);
        ipp_chain_rot r;
        memset(&r, 0, sizeof(r));
        ipp_rot_item* it = items_out + i;
        memset(it, 0, sizeof(*it));
        const int rc = rotation_plan(src_w, src_h, angle, &r, it->m);
        IPP_REQUIRE(rc == 0, rc, "ipp_plan_rotate_batch_host: rotation outside libImaging's 16.16 fixed-point range");
        angles_out[i] = angle;
        it->src = i; it->kind = r.kind;
        it->a0 = r.a0; it->a1 = r.a1; it->a2 = r.a2; it->a3 = r.a3; it->a4 = r.a4; it->a5 = r.a5;
        it->dst_w = r.canvas_w; it->dst_h = r.canvas_h;
        it->tiles_x = (r.canvas_w + kRotTileW - 1) / kRotTileW;    // rotate_batch_kernel's output tiles
        it->tile_start = (int32_t)tiles;
        tiles += (int64_t)it->tiles_x * ((r.canvas_h + kRotTileH - 1) / kRotTileH);
        IPP_REQUIRE(tiles < (1LL << 31), IPP_ERR_UNSUPPORTED, "ipp_plan_rotate_batch_host: too many tiles; use a smaller batch");
        it->dst_off = off;
        off = align_up(off + (int64_t)((r.canvas_w + 3) & ~3) * 4 * r.canvas_h, 256);
    }
    totals_out[0] = off;
    totals_out[1] = tiles;
    return IPP_OK;
}

// Batched overlay step (overlays.py:104-149) over n_pairs (sprite k % n_sprites, background k % n_bgs) pairs.
int ipp_plan_overlay_batch_host(int spr_w, int spr_h, int n_sprites, int n_pairs, const ipp_plan_params* p,
                                uint32_t* rng_mt624, int32_t* rng_index,
                                ipp_chain_item* items_out, int32_t* n_items_out,
                                char* labels_out, int labels_capacity, ipp_plan_totals* totals_out) {
    using namespace ipp;
    using namespace ipp::plan;
    IPP_REQUIRE(p && rng_mt624 && rng_index && items_out && n_items_out && labels_out && totals_out && spr_w > 0 && spr_h > 0 &&
                    n_sprites >= 1 && n_pairs >= 0 && p->n_bgs >= 1,
                IPP_ERR_BAD_ARG, "ipp_plan_overlay_batch_host: bad arguments");
    PyRandom rng{rng_mt624, rng_index};
    const int BW = p->bg_w, BH = p->bg_h;
    int n = 0;
    size_t lpos = 0;
    for (int k = 0; k < n_pairs; ++k) {
        const double ratio = rng.uniform(p->scale_min, p->scale_max);
        int nw, nh;
        if (!overlay_size(spr_w, spr_h, BW, BH, ratio, &nw, &nh)) continue;       // ValueError -> print, None
        bool ok;
        const int px = (int)rng.randint(0, BW - nw, &ok);
        if (!ok) continue;
        const int py = (int)rng.randint(0, BH - nh, &ok);
        if (!ok) continue;
        const long long x1 = px, y1 = py, x2 = px + nw, y2 = py + nh;
        const float vals[4] = {(float)(((double)(x1 + x2) / 2.0) / (double)BW), (float)(((double)(y1 + y2) / 2.0) / (double)BH),
                               (float)((double)(x2 - x1) / (double)BW), (float)((double)(y2 - y1) / (double)BH)};
        char row[256];
        int len = snprintf(row, 16, "%d", p->yolo_class_id);
        for (int i = 0; i < 4; ++i) {
            row[len++] = ' ';
            len += fmt6(row + len, (double)vals[i], nullptr);
        }
        row[len] = 0;                                                              // overlays.py:149: no trailing newline
        IPP_REQUIRE(lpos + (size_t)len + 1 <= (size_t)labels_capacity, IPP_ERR_WORKSPACE, "ipp_plan_overlay_batch_host: label buffer too small");
        memcpy(labels_out + lpos, row, (size_t)len + 1);
        lpos += (size_t)len + 1;
        ipp_chain_item* it = items_out + n;
        memset(it, 0, sizeof(*it));
        it->frame = k % n_sprites; it->bg = k % p->n_bgs;
        it->kind = 1; it->cc_w = spr_w; it->cc_h = spr_h; it->spr_w = spr_w; it->spr_h = spr_h;
        it->new_w = nw; it->new_h = nh; it->pos_x = px; it->pos_y = py;
        it->sq_size = BW < BH ? BW : BH;
        it->cc_x = k;                                                              // pair index (items of skipped pairs are absent)
        ++n;
    }
    *n_items_out = n;
    const int lrc = ipp_chain_layout_planar_host(items_out, n, totals_out);
    if (lrc != IPP_OK) return lrc;
    totals_out->labels_bytes = (int32_t)lpos;
    return IPP_OK;
}

// Photometric jitter step (tranfo.py:37-51): the draws of n images in order, the GaussianBlur box parameters exactly as
// libImaging derives them (BoxBlur.c: _gaussian_blur_radius with single-precision variables, ImagingHorizontalBoxBlur's
// 8.24 weights) and the channel tables as Image.point builds them (round-half-even of the clamped product).
int ipp_plan_enhance_host(int n, const uint8_t* do_blur, const uint8_t* do_rgb, uint32_t* rng_mt624, int32_t* rng_index,
                          ipp_enhance_item* items_out, uint8_t* luts_out, double* draws_out) {
    using namespace ipp;
    using namespace ipp::plan;
    IPP_REQUIRE(n >= 0 && do_blur && do_rgb && rng_mt624 && rng_index && items_out && luts_out && draws_out, IPP_ERR_BAD_ARG,
                "ipp_plan_enhance_host: bad arguments");
    PyRandom rng{rng_mt624, rng_index};
    for (int i = 0; i < n; ++i) {
        ipp_enhance_item* it = items_out + i;
        memset(it, 0, sizeof(*it));
        double* d = draws_out + 4 * i;
        d[0] = rng.uniform(0.7, 1.3); d[1] = rng.uniform(0.7, 1.3); d[2] = rng.uniform(0.7, 1.3); d[3] = NAN;
        it->brightness = (float)d[0]; it->contrast = (float)d[1]; it->color = (float)d[2];
        if (do_blur[i]) {
            d[3] = rng.uniform(0.5, 3.0);
            const int passes = 3;
            const float radius = (float)d[3];
            const float sigma2 = radius * radius / passes;
            const float L = (float)std::sqrt(12.0 * sigma2 + 1.0);
            const float l = (float)std::floor((L - 1.0) / 2.0);
            float a = (2 * l + 1) * (l * (l + 1) - 3 * sigma2);
            a /= 6 * (sigma2 - (l + 1) * (l + 1));
            const float box = l + a;
            if (box != 0.0f) {
                it->blur = 1;
                it->box_radius = (int)box;
                it->ww = (uint32_t)((uint32_t)(1 << 24) / (box * 2 + 1));
                it->fw = (uint32_t)(((1 << 24) - (it->box_radius * 2 + 1) * (long long)it->ww) / 2);
            }
        }
        uint8_t* lut = luts_out + (size_t)768 * i;
        if (do_rgb[i]) {
            it->lut = 1;
            for (int j = 0; j < 768; ++j) {
                const double v = (double)(j & 255) * rng.uniform(0.75, 1.25);
                lut[j] = (uint8_t)std::nearbyint(std::fmax(0.0, std::fmin(255.0, v)));       // Python round(): half to even
            }
        } else {
            for (int j = 0; j < 768; ++j) lut[j] = (uint8_t)(j & 255);
        }
    }
    return IPP_OK;
}

// the planner's exact-decimal helpers, exported so that tests can hold them against Python's round() / "%.6f"
double ipp_host_round15(double x) { return ipp::plan::round15(x); }
double ipp_host_hypot(double a, double b) { return ipp::plan::py_hypot2(a, b); }
int ipp_host_overlay_size(int ow, int oh, int bw, int bh, double ratio, int32_t* nw_nh_out) {
    if (!nw_nh_out) return -1;
    int nw = 0, nh = 0;
    const bool ok = ipp::plan::overlay_size(ow, oh, bw, bh, ratio, &nw, &nh);
    nw_nh_out[0] = nw; nw_nh_out[1] = nh;
    return ok ? 1 : 0;
}
int ipp_host_format6(double v, char* text_out48, double* parsed_out) {
    if (!text_out48) return -1;
    const int len = ipp::plan::fmt6(text_out48, v, parsed_out);
    text_out48[len] = 0;
    return len;
}

// one exact table row (libm sin) for the rare fix-up of rows the device flagged as ambiguous
int ipp_lanczos_row_host(int in_size, int out_size, int xx, int32_t* bounds2_host, int32_t* kk_row_host) {
    using namespace ipp;
    IPP_REQUIRE(in_size > 0 && out_size > 0 && xx >= 0 && xx < out_size && bounds2_host && kk_row_host, IPP_ERR_BAD_ARG,
                "ipp_lanczos_row_host: bad arguments");
    double scale, filterscale, support;
    int ksize;
    lanczos_scale(in_size, out_size, &scale, &filterscale, &support, &ksize);
    const double ss = 1.0 / filterscale;
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    std::vector<double> w((size_t)ksize, 0.0);
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        w[x] = lanczos_filter((x + xmin - center + 0.5) * ss);
        ww += w[x];
    }
    for (int x = 0; x < ksize; ++x) {
        double v = x < xmax ? w[x] : 0.0;
        if (x < xmax && ww != 0.0) v /= ww;
        kk_row_host[x] = (v < 0) ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
    }
    bounds2_host[0] = xmin;
    bounds2_host[1] = xmax;
    return IPP_OK;
}

}  // extern "C"
