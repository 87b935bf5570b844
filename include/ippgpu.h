/*
 * ippgpu.h — C ABI of the B200-native per-image transform chain.
 *
 * Drop-in boundary for the hot path of Tezahc/image_processor_pipeline: the
 * arithmetic executed inside one "process function" call
 * (reference pipeline.py:332-336 -> transforms/*.py) between "pixels decoded"
 * and "pixels ready to encode".  The reference has no FFI of its own (it is pure
 * Python over cv2 / Pillow); each entry point below replaces the library call
 * sequence of the cited reference lines.  INTEGRATION.md shows the ctypes stub
 * a maintainer of the reference would add.
 *
 * Conventions
 *  - plain C symbols, no C++/torch types; every image pointer is a DEVICE pointer
 *    owned by the caller (the library never allocates or frees user-visible
 *    memory); `*_host` parameters are HOST pointers read during the call;
 *  - images are interleaved HWC uint8 with an explicit row pitch in bytes;
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    all work is enqueued asynchronously on it; calls are re-entrant per stream;
 *  - return value: 0 = ok; negative = IPP_ERR_* (invalid argument class, mapped
 *    to the Python exception the reference function raises); positive = the
 *    cudaError_t of a failed runtime call.  ipp_last_error_string() describes
 *    the last failure on the calling thread.  No exceptions cross the ABI.
 *  - sm_100a only.  There is no CPU fallback.
 *  - memory access guarantee: 3-byte pixels are gathered through the aligned 32-bit word(s) that hold them
 *    (stage B2 / the batched rotation) and bit masks through aligned 16-byte quads (stage B1, inside the
 *    library-owned workspace).  Every word that is loaded contains at least one byte of the addressed pixel, so
 *    a kernel never reads beyond the 4-byte-aligned extent of a caller's buffer — i.e. never outside the memory
 *    page of its last valid byte; no padding is required of the caller.  (A buffer whose length is not a
 *    multiple of four is read up to 3 bytes past its end inside that last word; CUDA allocations are at least
 *    256-byte granular, so this is always backed memory.)
 */
#ifndef IPPGPU_H
#define IPPGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPP_VERSION 100

#define IPP_OK 0
#define IPP_ERR_BAD_ARG (-1)      /* -> ValueError   */
#define IPP_ERR_BAD_SHAPE (-2)    /* -> ValueError   */
#define IPP_ERR_WORKSPACE (-3)    /* -> MemoryError  */
#define IPP_ERR_UNSUPPORTED (-4)  /* -> NotImplementedError */

#define IPP_MAX_HSV_RANGES 16
#define IPP_MAX_TAPS_NOTE "resample taps are unbounded; tables are caller-provided"

int ipp_version(void);
const char* ipp_last_error_string(void);
/* number of SMs of the current device (grid sizing on the host side) */
int ipp_sm_count(int* out_host);
/* launch accounting: total kernels launched by this library in the process; optional
 * per-kernel device timing (CUDA events on the launching stream around every launch).
 * ipp_timing_report writes "kernel:launches:total_ms;..." for the launches since the
 * previous report into buf_host and synchronises on them. */
long long ipp_launch_count(void);
int ipp_timing_enable(int on);
int ipp_timing_report(char* buf_host, int buflen);
/* same launches as a timeline: "kernel:stream_id:start_ms:end_ms;..." in launch order, times relative to the
 * first launch (events of all streams share the device clock); synchronises and clears like the report. */
int ipp_timing_timeline(char* buf_host, int buflen);

/* ---------------------------------------------------------------------------------
 * 1. flip x crop x copy as one load-address remap.
 *    Replaces cv2.flip / ndarray.copy (reference transforms/symmetry.py:114-119),
 *    ndarray slicing crops (recadrages.py:46, crop_square.py:196, pixels_isolés.py:81)
 *    and PIL crop (rotations.py:101, recadrages.py:77).
 *    dst[y, x, :] = src[y0 + (flip_v ? out_h-1-y : y), x0 + (flip_h ? out_w-1-x : x), :]
 *    bpp = bytes per pixel (1..16).  n images of identical geometry are processed
 *    when n > 1 (image i at src + i*src_image_stride / dst + i*dst_image_stride).
 * --------------------------------------------------------------------------------- */
int ipp_remap_u8(const uint8_t* src, int64_t src_pitch, int src_w, int src_h, int bpp,
                 int x0, int y0, int out_w, int out_h, int flip_h, int flip_v,
                 uint8_t* dst, int64_t dst_pitch,
                 int n, int64_t src_image_stride, int64_t dst_image_stride, void* stream);

/*    Fused symmetries: ONE read of the source, up to four writes
 *    (reference transforms/symmetry.py:128-137 writes each selected variant).
 *    Any of dst_o / dst_h / dst_v / dst_hv may be NULL (variant not wanted). */
int ipp_symmetries_u8(const uint8_t* src, int64_t src_pitch, int w, int h, int bpp,
                      uint8_t* dst_o, uint8_t* dst_h, uint8_t* dst_v, uint8_t* dst_hv,
                      int64_t dst_pitch,
                      int n, int64_t src_image_stride, int64_t dst_image_stride, void* stream);

/*    gray/RGB -> RGBA with A=255 (PIL convert("RGBA"), reference rotations.py:55,
 *    overlays.py:85) and channel-order swap BGR(A)<->RGB(A) between cv2 and PIL
 *    conventions.  src_channels in {1,3,4}; dst is always 4 channels. */
int ipp_to_rgba_u8(const uint8_t* src, int64_t src_pitch, int w, int h, int src_channels,
                   int swap_rb, uint8_t* dst, int64_t dst_pitch, void* stream);

/*    Quarter turns of an RGBA image: PIL Image.rotate fast paths for 90 / 270 degrees
 *    (PIL/Image.py:2509-2521 -> transpose(ROTATE_90 | ROTATE_270)).  dst is h x w. */
int ipp_rot90_rgba_u8(const uint8_t* src, int64_t src_pitch, int w, int h, int counter_clockwise,
                      uint8_t* dst, int64_t dst_pitch, void* stream);

/* ---------------------------------------------------------------------------------
 * 2. bounding box of non-zero pixels.
 *    PIL getbbox() (rotations.py:99, recadrages.py:73: alpha band for 2/4-channel
 *    images, any band otherwise) and cv2.findNonZero+boundingRect on the alpha
 *    plane (pixels_isolés.py:77-79).
 *    bbox_out (device int32[4]) = {x0, y0, x1, y1} exclusive; empty <=> x1 == 0
 *    (then x0 = y0 = INT_MAX).
 * --------------------------------------------------------------------------------- */
int ipp_alpha_bbox_u8(const uint8_t* img, int64_t pitch, int w, int h, int channels,
                      int alpha_only, int32_t* bbox_out, void* stream);

/* ---------------------------------------------------------------------------------
 * 3. inverse-mapped rotation warp (reference transforms/rotations.py:96,
 *    PIL Image.rotate(expand=True) -> libImaging affine_fixed / bilinear).
 *    mode 0: NEAREST, 16.16 fixed point: xin=(a2+y*a1+x*a0)>>16, yin=(a5+y*a4+x*a3)>>16
 *            (coef_fixed_host = int32[6] {a0..a5}); bit-exact.
 *    mode 1: BILINEAR on premultiplied alpha, double-precision coordinates
 *            (coef_f64_host = double[6] PIL matrix); <= 1 LSB (documented).
 *    The canvas (dst_w x dst_h, 4 channels) is fully written (0 outside the
 *    source).  If bbox_out != NULL the bbox {x0,y0,x1,y1} of alpha != 0 is reduced
 *    in the same launch (PIL getbbox, rotations.py:99); it must be pre-set by the
 *    caller to {INT_MAX, INT_MAX, 0, 0} (ipp_bbox_reset does that).
 * --------------------------------------------------------------------------------- */
int ipp_bbox_reset(int32_t* bbox, int n, void* stream);
int ipp_rotate_rgba_u8(const uint8_t* src, int64_t src_pitch, int src_w, int src_h,
                       int mode, const int32_t* coef_fixed_host, const double* coef_f64_host,
                       uint8_t* dst, int64_t dst_pitch, int dst_w, int dst_h,
                       int32_t* bbox_out, void* stream);

/* ---------------------------------------------------------------------------------
 * 4. Pillow-compatible 8bpc separable resample of an RGBA image
 *    (reference transforms/overlays.py:129, Image.resize(LANCZOS): RGBA -> RGBa
 *    premultiply, horizontal pass, vertical pass, un-premultiply; bit-exact).
 *    Coefficient tables are computed on the host exactly as libImaging
 *    precompute_coeffs + normalize_coeffs_8bpc do and passed as DEVICE arrays:
 *      h_bounds int32[dst_w][2] = {xmin, count}, h_kk int32[dst_w][h_ksize]
 *      v_bounds int32[dst_h][2],                  v_kk int32[dst_h][v_ksize]
 *    A pass is skipped when its size is unchanged (h_ksize == 0 / v_ksize == 0).
 *    tmp: device scratch of at least src_h * dst_w * 4 bytes.
 * --------------------------------------------------------------------------------- */
/* host helpers: ksize = 2*ceil(3*max(in/out,1))+1; tables as Pillow computes them
 * (C library sin(), double precision, same operation order).  bounds_host is
 * int32[out_size][2], kk_host int32[out_size][ksize] — both HOST arrays. */
int ipp_lanczos_ksize(int in_size, int out_size);
int ipp_lanczos_coeffs_host(int in_size, int out_size, int32_t* bounds_host, int32_t* kk_host);
int ipp_resample_rgba_u8(const uint8_t* src, int64_t src_pitch, int src_w, int src_h,
                         uint8_t* dst, int64_t dst_pitch, int dst_w, int dst_h,
                         const int32_t* h_bounds, const int32_t* h_kk, int h_ksize,
                         const int32_t* v_bounds, const int32_t* v_kk, int v_ksize,
                         uint8_t* tmp, void* stream);

/* ---------------------------------------------------------------------------------
 * 5. alpha paste of an RGBA sprite onto an RGB background, optionally cropped
 *    (reference transforms/overlays.py:138-139: bg.copy().paste(spr, pos, spr);
 *    libImaging paste_mask_RGBA: out = DIV255(bg*(255-a) + spr*a); bit-exact).
 *    dst (crop_w x crop_h x 3) = composite[crop_y0 : crop_y0+crop_h, crop_x0 : ...]
 *    — crop = (0, 0, bg_w, bg_h) reproduces the plain paste; a square crop
 *    (reference crop_square.py:196) folds into the same pass.
 * --------------------------------------------------------------------------------- */
int ipp_paste_rgba_on_rgb_u8(const uint8_t* bg, int64_t bg_pitch, int bg_w, int bg_h,
                             const uint8_t* sprite, int64_t sprite_pitch, int sprite_w, int sprite_h,
                             int pos_x, int pos_y,
                             int crop_x0, int crop_y0, int crop_w, int crop_h,
                             uint8_t* dst, int64_t dst_pitch, void* stream);

/* ---------------------------------------------------------------------------------
 * 6. HSV range keying (reference transforms/filtres_liste.py:90-134:
 *    cvtColor(BGR2HSV) -> OR_k(inRange_k & zone_k) -> alpha = ~mask; bit-exact).
 *    Bounds are integers already rounded like cv2.inRange rounds scalars on CV_8U
 *    (half-to-even); zone = margins {top, bottom, left, right}.
 * --------------------------------------------------------------------------------- */
typedef struct ipp_hsv_range {
    int32_t lo[3];   /* H, S, V lower bounds (inclusive) */
    int32_t hi[3];   /* H, S, V upper bounds (inclusive) */
    int32_t zone[4]; /* top, bottom, left, right margins in pixels */
} ipp_hsv_range;

int ipp_hsv_ranges_to_alpha_u8(const uint8_t* bgr, int64_t src_pitch, int w, int h,
                               const ipp_hsv_range* ranges_host, int n_ranges,
                               uint8_t* bgra, int64_t dst_pitch,
                               int n, int64_t src_image_stride, int64_t dst_image_stride, void* stream);

/* ---------------------------------------------------------------------------------
 * 7. keep the largest 8-connected component of alpha > 1
 *    (reference transforms/pixels_isolés.py:32-55 + _crop_fit :74-81:
 *    cv2.threshold + connectedComponentsWithStats(8) + argmax area (ties: first
 *    label in OpenCV's order = smallest 2x2-block raster key) + alpha[not kept]=0
 *    + bounding rect of alpha != 0; bit-exact).
 *    The alpha channel of the 4-channel image is rewritten in place; colour bytes
 *    are untouched.  info_out (device int32[8] per image) =
 *    {x, y, w, h of alpha != 0 after cleaning (0,0,0,0 if none), area, n_fg_pixels, 0, 0}.
 *    workspace: ipp_keep_largest_workspace_bytes(w, h, n) bytes of device scratch.
 * --------------------------------------------------------------------------------- */
int64_t ipp_keep_largest_workspace_bytes(int w, int h, int n);
int ipp_keep_largest_cc_u8(uint8_t* bgra, int64_t pitch, int w, int h,
                           int32_t* info_out, void* workspace, int64_t workspace_bytes,
                           int n, int64_t image_stride, void* stream);

/* ---------------------------------------------------------------------------------
 * 8. YOLO box arithmetic, one warp per image (float32 store semantics of
 *    ultralytics.utils.ops as the reference uses them).
 *    ipp_boxes_square_crop_f32: reference crop_square.py:167,201-217 —
 *      xywhn -> xyxy (float32) -> shift by (x0,y0) -> clip to [0,cs] -> keep
 *      x1<x2 && y1<y2 -> xyxy -> xywhn(cs,cs) (float32).
 *      boxes_in/out: float64[n_boxes][4] in, float32[n_boxes][4] out, valid_out
 *      int32[n_boxes]; boxes of image i are rows box_offsets[i] .. box_offsets[i+1].
 *      params int32[n_images][5] = {img_w, img_h, x0, y0, crop_size}.
 *    ipp_boxes_xyxy2xywhn_f32: reference overlays.py:143-146 — int box
 *      {x1,y1,x2,y2} -> float32 {cx,cy,w,h} normalised by (img_w, img_h).
 * --------------------------------------------------------------------------------- */
int ipp_boxes_square_crop_f32(const double* boxes_in, const int32_t* box_offsets, const int32_t* params,
                              float* boxes_out, int32_t* valid_out, int n_images, void* stream);
int ipp_boxes_xyxy2xywhn_f32(const int32_t* xyxy, const int32_t* img_wh, float* xywhn_out,
                             int n_boxes, void* stream);

/* ---------------------------------------------------------------------------------
 * 9. Fused chain (batched, device-resident): frame -> border crop -> HSV key ->
 *    keep largest component + fit -> rotate (nearest) + fit -> flip -> LANCZOS
 *    resize -> paste on background -> square crop.
 *    Reference order of steps: SURVEY.md §7.2 (video.py:42 -> recadrages.py:23 ->
 *    filtres_liste.py:138 -> pixels_isolés.py:16 -> rotations.py:55 -> symmetry.py ->
 *    overlays.py -> crop_square.py).  Intermediate images are never materialised at
 *    full size: the frame is read once into a 1-bit keep-mask, the sprite is gathered
 *    once (rotate∘fit∘flip folded into the addressing), and the background crop,
 *    the alpha blend and the final square crop are one streaming pass.
 *    The chain is split only where data-dependent sizes must reach the host RNG
 *    (randint bounds depend on the sprite size): stage A | B1 | B2+C.
 *    All frames of a batch share one geometry; all backgrounds share one geometry.
 * --------------------------------------------------------------------------------- */
typedef struct ipp_chain_geom {
    int32_t frame_w, frame_h;               /* decoded frame (BGR, 3 bytes/px)           */
    int32_t crop_l, crop_t, crop_w, crop_h; /* border crop rect (recadrages.py:46)        */
    int32_t bg_w, bg_h;                     /* backgrounds (RGB, 3 bytes/px)              */
    int32_t n_ranges;
    ipp_hsv_range ranges[IPP_MAX_HSV_RANGES];
} ipp_chain_geom;

/* one rotated sprite whose alpha bbox is wanted (stage B1); 16 x int32 */
typedef struct ipp_chain_rot {
    int32_t frame;                      /* source frame index                               */
    int32_t cc_x, cc_y, cc_w, cc_h;     /* kept component's rect inside the cropped frame   */
    int32_t kind;                       /* 0 affine, 1 copy, 2 rot180, 3 rot90, 4 rot270    */
    int32_t a0, a1, a2, a3, a4, a5;     /* libImaging 16.16 coefficients (kind 0)           */
    int32_t canvas_w, canvas_h;
    int32_t tile_start;                 /* first stage-B1 tile of this sprite in the flat grid; a tile =
                                           256 aligned quads of mask words of the component rect, whose rows span
                                           nq = ((((cc_x+cc_w+31)>>5)+3)>>2) - ((cc_x>>5)>>2) quads:
                                           ceil(nq*cc_h/256) tiles */
    int32_t _pad;
} ipp_chain_rot;

/* one output image (stages B2 and C); 40 x int32 + 3 x int64 */
typedef struct ipp_chain_item {
    int32_t frame, bg;
    int32_t cc_x, cc_y, cc_w, cc_h;
    int32_t kind, a0, a1, a2, a3, a4, a5;
    int32_t fit_x0, fit_y0, spr_w, spr_h;   /* getbbox crop inside the canvas, sprite size     */
    int32_t flip_h, flip_v;
    int32_t new_w, new_h, pos_x, pos_y;     /* resized sprite and paste position (overlays.py) */
    int32_t sq_x0, sq_y0, sq_size;          /* final square crop (crop_square.py)              */
    int32_t h_ksize, v_ksize;               /* 0 = pass skipped (size unchanged)               */
    int32_t hb_off, hk_off, vb_off, vk_off; /* int32 offsets of this item's tables             */
    int32_t gather_tile_start, h_tile_start, v_tile_start;
    int32_t tab_row_start;                  /* first coefficient-table row of this item (h rows, then v rows) */
    int32_t out_slot;                       /* index of the output image this item writes            */
    int32_t hband_off, vband_off;           /* 8-byte offsets of this item's banded int8 fragment tables (h, v) */
    int32_t nkb;                            /* 64-wide super-blocks (two MMA k-blocks) per 8-output tile: h | v << 16 */
    int64_t rot_off, tmp_off, spr_off;      /* byte offsets in the sprite workspace: planar premultiplied sprite
                                               [4][spr_h][Wp], transposed planar h-pass result [4][new_w][Hp],
                                               final RGBA sprite [new_h][new_w][4]; Wp/Hp = spr_w/spr_h rounded up to 16;
                                               the two planar buffers carry 16 extra rows + 1 KB (the passes read
                                               whole 16-row tiles and 64-byte windows without clamping) */
} ipp_chain_item;

/* ---- host planner (C++ twin of chain.py's planner; no device work) ------------------
 * The reference draws its random parameters from Python's `random` (MT19937), step-major;
 * rng_mt624 / rng_index are random.getstate()[1][:624] / [624] and are advanced in place. */
typedef struct ipp_plan_params {
    int32_t num_rotations, include_original;      /* rotations.py:9-10                         */
    double angle_min, angle_max;                  /* rotations.py:11-12                        */
    int32_t pool[4], pool_len;                    /* symmetry codes 0 o, 1 h, 2 v, 3 hv        */
    int32_t choose_random, sym_include_original;  /* symmetry.py:17-18 (choose_random resolved) */
    int32_t yolo_class_id;                        /* overlays.py:30                            */
    double scale_min, scale_max;                  /* overlays.py:31-32                         */
    int32_t bg_w, bg_h, n_bgs;
    int32_t _pad;
} ipp_plan_params;

typedef struct ipp_plan_aux {
    double ratio;        /* overlay target ratio drawn for this item          */
    int32_t rot;         /* index of the sprite in the rots array             */
    int32_t sym;         /* 0 o, 1 h, 2 v, 3 hv                               */
} ipp_plan_aux;

typedef struct ipp_plan_totals {
    int64_t table_ints, table_rows, sprite_ws_bytes;
    int64_t band_units;                     /* size of the fragment-table buffer in 8-byte units      */
    int32_t gather_tiles, h_tiles, v_tiles, labels_bytes;
    int32_t max_rows, max_band_warps;       /* per-item maxima: coefficient rows, fragment-table warps */
    int32_t fused_smem_bytes;               /* > 0: the fused gather + h pass kernel is used with this much shared memory */
    int32_t vstaged_smem_bytes;             /* > 0: the CTA-staged vertical pass is used with this much shared memory       */
} ipp_plan_totals;

int ipp_plan_rotations_host(const int32_t* info, int n_frames, const ipp_plan_params* p, uint32_t* rng_mt624, int32_t* rng_index,
                            ipp_chain_rot* rots_out, double* angles_out, int32_t* rot_index_out, int capacity,
                            int32_t* n_rot_out, int32_t* total_tiles_out);
int ipp_plan_items_host(const ipp_chain_rot* rots, const int32_t* rot_is_original, int n_rot, const int32_t* bbox,
                        const ipp_plan_params* p, uint32_t* rng_mt624, int32_t* rng_index,
                        ipp_chain_item* items_out, ipp_plan_aux* aux_out, int capacity, int32_t* n_items_out,
                        char* labels_out, int labels_capacity, ipp_plan_totals* totals_out);
int ipp_lanczos_row_host(int in_size, int out_size, int xx, int32_t* bounds2_host, int32_t* kk_row_host);
/* Exact-decimal helpers of the planner (128-bit integer arithmetic on the binary value, no printf on the hot
 * path): Python's round(x, 15) (PIL:Image.py:2541-2548 rounds the rotation matrix) and "%.6f" of a label number
 * together with the float64 np.loadtxt reads back from that text (overlays.py:149, crop_square.py:62-65,98-101). */
double ipp_host_round15(double x);
/* CPython's math.hypot (two arguments) and the overlay size math of overlays.py:106-126 built on it, exported for tests:
 * returns 1 and {new_w, new_h} in nw_nh_out, or 0 when the resize would fail (size <= 0). */
double ipp_host_hypot(double a, double b);
int ipp_host_overlay_size(int ow, int oh, int bw, int bh, double ratio, int32_t* nw_nh_out);
int ipp_host_format6(double v, char* text_out48, double* parsed_out);
/* Device-side schedule of an item array (sizes, kind, flips, positions already set): coefficient-table
 * and fragment-table offsets, flat tile starts, sprite-workspace offsets, out_slot = index.  Called by
 * ipp_plan_items_host; exported for callers that assemble item arrays themselves. */
int ipp_chain_layout_host(ipp_chain_item* items_host, int n_items, ipp_plan_totals* totals_out);

/* LANCZOS coefficient tables of all items computed ON THE DEVICE (double precision, the
 * libImaging operation order).  The C library's sin() cannot be reproduced bit for bit on the
 * GPU, so every row with a coefficient whose rounding decision lies within `eps` (in units of
 * 2^-22) of a boundary is reported together with its device-computed content:
 * amb_out[0] = count; record r < amb_capacity at amb_out + 64 * (r + 1) =
 * {item, axis, row, xmin, xmax, ksize, k[0 .. min(ksize, 58) - 1]}  (amb_out holds
 * 64 * (amb_capacity + 1) int32).  The caller re-derives those rows with ipp_lanczos_row_host
 * and patches `tables` where a row really differs BEFORE ipp_chain_stage_b2 reads them —
 * results stay bit-exact. */
int ipp_chain_tables(const ipp_chain_item* items, int n_items, int max_rows_per_item, int32_t* tables,
                     int32_t* amb_out, int amb_capacity, double eps, void* stream);

int64_t ipp_chain_mask_workspace_bytes(const ipp_chain_geom* geom_host, int n_frames);
/* stage A: frames (n x frame_h x frame_w x 3, BGR) -> keep-mask (inside workspace) and
 * info_out int32[n][8] = {x, y, w, h of the kept component (cropped-frame coords), area,
 * n_fg, 0, 0} (all zero when the key removed everything). */
int ipp_chain_stage_a(const uint8_t* frames, int64_t frame_pitch, int64_t frame_stride,
                      const ipp_chain_geom* geom_host, int32_t* info_out,
                      void* workspace, int64_t workspace_bytes, int n_frames, void* stream);
/* stage B1: alpha bbox of every rotated sprite -> bbox_out int32[n_rot][4]
 * {x0,y0,x1,y1}; empty <=> x1 == 0.  total_tiles = sum over sprites of their tiles (see tile_start).
 * Only the 8-boundary pixels of the kept component are mapped (the extreme pixels of the rotated
 * sprite are images of boundary pixels), each checked with the exact fixed-point inverse map. */
int ipp_chain_stage_b1(const ipp_chain_geom* geom_host, const ipp_chain_rot* rots, int n_rot, int total_tiles,
                       const void* workspace, int32_t* bbox_out, void* stream);
/* stage B2: rotate∘fit∘flip gather (premultiplied, four byte planes) -> LANCZOS h pass -> v pass +
 * un-premultiply -> straight-alpha RGBA sprites in sprite_ws at items[i].spr_off.
 * The two passes are exact integer banded GEMMs on the tensor cores (mma.sync m16n8k32, u8 data x
 * coefficients split into three byte planes, int32 accumulation = Pillow's own arithmetic); `band`
 * (totals.band_units * 8 bytes) receives the coefficient bands in fragment order, derived from
 * `tables` by this call (max_band_warps = totals.max_band_warps).
 * (Fragment format per axis: 64-position super-blocks for the passes that stream from global memory, 32-position blocks
 * with a per-tile block count for the passes that work from shared memory; it follows the two *_smem_bytes arguments below.)
 * fused_smem_bytes = totals.fused_smem_bytes: when > 0 (every sprite of the batch is narrow enough) gather and h
 * pass are ONE kernel that keeps 16 gathered rows x 4 planes in shared memory (gather_tiles = its tiles: one per 16
 * sprite rows, h_tiles = 0); when 0 they are two kernels exchanging the planar sprite through sprite_ws
 * (gather_tiles of 32 x 64 pixels, h_tiles of 64 outputs x plane).  gather_tiles = 0 skips the gather in the
 * two-kernel mode only (h / v passes re-run on the sprites already gathered).
 * vstaged_smem_bytes = totals.vstaged_smem_bytes: when > 0 the vertical pass stages 16 columns x 4 planes of the
 * transposed intermediate per CTA in shared memory (v_tiles: one per 16 sprite columns); when 0 every warp streams
 * its windows through a private cp.async ring (v_tiles of 64 output rows x 128 columns). */
int ipp_chain_stage_b2(const uint8_t* frames, int64_t frame_pitch, int64_t frame_stride,
                       const ipp_chain_geom* geom_host, const ipp_chain_item* items, int n_items,
                       int gather_tiles, int h_tiles, int v_tiles,
                       const int32_t* tables, void* band, int max_band_warps, int fused_smem_bytes, int vstaged_smem_bytes,
                       const void* workspace, uint8_t* sprite_ws, void* stream);
/* stage C: out[i] (sq x sq x 3, RGB) = crop(paste(background[bg], sprite_i, pos), square) */
int ipp_chain_stage_c(const uint8_t* backgrounds, int64_t bg_pitch, int64_t bg_stride,
                      const ipp_chain_geom* geom_host, const ipp_chain_item* items, int n_items,
                      const uint8_t* sprite_ws,
                      uint8_t* out, int64_t out_pitch, int64_t out_stride, int out_size, void* stream);

/* Host-background variant of stage C's input: copies, for every item, ONLY the square-crop window
 * [sq_y0, sq_y0+out_size) x [sq_x0, sq_x0+out_size) of its background from (pinned) host memory to
 * windows_dev[i] (out_size x out_size x 3, dense) with one cudaMemcpy2DAsync per item on `stream`
 * — the final square crop (crop_square.py:196) folded into the host->device copy, so the
 * columns/rows the crop throws away never cross PCIe.  Stage C is then called on windows_dev with
 * items whose bg = i, sq_x0 = sq_y0 = 0 and pos_x/pos_y shifted by the window origin.
 * items_host is a HOST array.  bgs_host must be PINNED host memory: short window rows (<= 4 KB) are gathered by a small kernel that
 * reads the host rows over PCIe itself (zero-copy loads; the copy engines move such 2-D windows at ~70 % of their contiguous
 * rate), longer rows go through cudaMemcpy2DAsync. */
int ipp_chain_upload_windows(const uint8_t* bgs_host, int64_t bg_pitch, int64_t bg_stride, int bg_w, int bg_h, int n_bgs,
                             const ipp_chain_item* items_host, int n_items, int out_size,
                             uint8_t* windows_dev, void* stream);

/* =================================================================================
 * Batched entry points of the stand-alone operators: n images per call, per-item parameter
 * arrays on the device, flat tile lists (one launch per batch whatever the item sizes).
 * These are what a ProcessingStep that maps one transform over a whole directory calls
 * (reference pipeline.py:304-437 runs the same process function over every file).
 * ================================================================================= */

/* -- crop with per-image offsets (crop_square.py:192-196: x0, y0 are drawn per image).
 *    xy_offsets: device int32[n][2] = {x0, y0} of image i; every crop is out_w x out_h. */
int ipp_crop_batch_u8(const uint8_t* src, int64_t src_pitch, int64_t src_image_stride, int src_w, int src_h, int bpp,
                      const int32_t* xy_offsets, int out_w, int out_h,
                      uint8_t* dst, int64_t dst_pitch, int64_t dst_image_stride, int n, void* stream);

/* -- rotation (rotations.py:55,96-99: convert("RGBA"), rotate(expand=True), getbbox) of n images.
 *    One record per output canvas; 104 bytes. */
typedef struct ipp_rot_item {
    int32_t src;                        /* source image index                                       */
    int32_t kind;                       /* 0 affine, 1 copy, 2 rot180, 3 rot90, 4 rot270            */
    int32_t a0, a1, a2, a3, a4, a5;     /* libImaging 16.16 coefficients (NEAREST, kind 0)          */
    int32_t dst_w, dst_h;               /* expanded canvas                                          */
    int32_t tile_start;                 /* first 64 x 256 output tile of this canvas in the flat grid */
    int32_t tiles_x;                    /* ceil(dst_w / 64); the canvas has tiles_x * ceil(dst_h / 256) tiles */
    int64_t dst_off;                    /* byte offset of the canvas in dst_pool (256-byte aligned);
                                           rows are ((dst_w + 3) & ~3) * 4 bytes apart              */
    double m[6];                        /* PIL inverse matrix (BILINEAR, kind 0)                    */
} ipp_rot_item;

/*    Host plan: one `random.uniform(angle_min, angle_max)` per image from the Python generator state
 *    (rotations.py:89), PIL's rotate parameter math, canvas offsets and tile starts.  angles_out: double[n].
 *    totals_out[0] = bytes of dst_pool, totals_out[1] = tiles. */
int ipp_plan_rotate_batch_host(int src_w, int src_h, int n, double angle_min, double angle_max,
                               uint32_t* rng_mt624, int32_t* rng_index,
                               ipp_rot_item* items_out, double* angles_out, int64_t* totals_out);

/*    src: n_src images of src_w x src_h, src_channels 3 (RGB, alpha = 255 supplied: the convert("RGBA") of
 *    rotations.py:55 folded into the load) or 4 (RGBA).  mode 0: NEAREST, bit-exact (16.16 fixed point);
 *    mode 1: BILINEAR on premultiplied alpha, double-precision source coordinates, single-precision
 *    interpolation: within 1 LSB per channel of Pillow (documented tolerance).  Every canvas is fully written
 *    (0 outside the source, padding columns included); bbox_out int32[n_items][4] = PIL getbbox of canvas i
 *    ({x0,y0,x1,y1}, empty <=> x1 == 0), reduced in the same launch. */
int ipp_rotate_batch_u8(const uint8_t* src, int64_t src_pitch, int64_t src_image_stride, int src_w, int src_h,
                        int src_channels, int n_src, const ipp_rot_item* items, int n_items, int total_tiles, int mode,
                        uint8_t* dst_pool, int32_t* bbox_out, void* stream);

/* -- overlay (overlays.py:104-149: LANCZOS resize of an RGBA sprite, paste onto an RGB background, YOLO label)
 *    of n (sprite, background) pairs.  Items are ipp_chain_item records: frame = sprite index, bg = background
 *    index, spr_w / spr_h = sprite size, new_w / new_h, pos_x / pos_y, sq_x0 = sq_y0 = 0.
 *    Host plan: per pair `uniform(scale_min, scale_max)`, `randint(0, W - nw)`, `randint(0, H - nh)` in the
 *    reference's order; pairs whose resize would fail are skipped (overlays.py:151-153 prints and returns None).
 *    labels_out: '\0'-separated label texts "{cls} {cx:.6f} {cy:.6f} {w:.6f} {h:.6f}" (overlays.py:149). */
int ipp_plan_overlay_batch_host(int spr_w, int spr_h, int n_sprites, int n_pairs, const ipp_plan_params* p,
                                uint32_t* rng_mt624, int32_t* rng_index,
                                ipp_chain_item* items_out, int32_t* n_items_out,
                                char* labels_out, int labels_capacity, ipp_plan_totals* totals_out);
/*    Like ipp_chain_layout_host with the gather and the horizontal pass as separate kernels (the planar
 *    premultiplied sprite lives in the sprite workspace): the layout of the batched resize / overlay path. */
int ipp_chain_layout_planar_host(ipp_chain_item* items_host, int n_items, ipp_plan_totals* totals_out);
/*    Straight-alpha RGBA sprites -> premultiplied byte planes [R, G, B, A][spr_h][Wp] at items[i].rot_off
 *    (Pillow's RGBA -> RGBa, MULDIV255): the input of ipp_chain_stage_b2 called with fused_smem_bytes = 0 and
 *    gather_tiles = 0.  sprites: n_sprites images of spr_w x spr_h x 4 (items[i].frame selects one). */
int ipp_sprite_planes_u8(const uint8_t* sprites, int64_t sprite_pitch, int64_t sprite_stride, int spr_w, int spr_h,
                         int n_sprites, const ipp_chain_item* items, int n_items, uint8_t* sprite_ws, void* stream);
/*    stage C on a rectangular window: out[i] (out_h x out_w x 3) = paste(background[bg], sprite_i, pos)[sq_y0:, sq_x0:]
 *    (out_w x out_h = the whole background for the plain overlay step). */
int ipp_chain_stage_c_rect(const uint8_t* backgrounds, int64_t bg_pitch, int64_t bg_stride, int bg_w, int bg_h,
                           const ipp_chain_item* items, int n_items, const uint8_t* sprite_ws,
                           uint8_t* out, int64_t out_pitch, int64_t out_stride, int out_w, int out_h, void* stream);
/*    Resized straight-alpha RGBA sprites out of the workspace: dst + i * dst_stride = sprite i
 *    (new_h x new_w x 4, dense); used by the stand-alone batched resize. */
int ipp_sprites_export_u8(const ipp_chain_item* items_host, int n_items, const uint8_t* sprite_ws,
                          uint8_t* dst, const int64_t* dst_offsets_host, void* stream);

/* -- HSV key + largest component of n frames without the BGRA image in between (filtres_liste.py:90-134 then
 *    pixels_isolés.py:29-58): call ipp_chain_stage_a (mask + keep-mask + info = bbox of the kept component, what
 *    _crop_fit of pixels_isolés.py:74-81 crops to), then this: out[i] (crop_h x crop_w x 4) = B, G, R of the
 *    border-cropped frame, alpha = 255 on the kept component and 0 elsewhere.  `workspace` = stage A's. */
int ipp_bgra_from_keep_u8(const uint8_t* frames, int64_t frame_pitch, int64_t frame_stride, const ipp_chain_geom* geom_host,
                          const void* workspace, uint8_t* out, int64_t out_pitch, int64_t out_stride, int n, void* stream);

/* -- photometric jitter of n RGB images (tranfo.py:37-51; last step of the workflow, pairing mode 'sample'
 *    pipeline.py:209-225): Brightness / Contrast / Color enhance (ImagingBlend in single precision), optional
 *    GaussianBlur (three 8.24 fixed-point box passes per axis), optional per-channel 256-entry tables.  Bit-exact. */
typedef struct ipp_enhance_item {
    float brightness, contrast, color;  /* enhance factors (tranfo.py:38-40: uniform(0.7, 1.3) each)            */
    int32_t blur;                       /* 1: GaussianBlur; box_radius / ww / fw as ImagingHorizontalBoxBlur derives them */
    int32_t box_radius;
    uint32_t ww, fw;
    int32_t lut;                        /* 1: per-channel tables at luts[i][3][256] (tranfo.py:46-51)            */
} ipp_enhance_item;
/*    Host plan, image by image in order: three uniform(0.7, 1.3), then uniform(0.5, 3) when do_blur[i], then
 *    3 x 256 uniform(0.75, 1.25) when do_rgb[i] (Image.point calls the lambda once per table entry), from the
 *    Python generator state.  draws_out: double[n][4] = the three factors and the blur radius (NaN: none). */
int ipp_plan_enhance_host(int n, const uint8_t* do_blur, const uint8_t* do_rgb, uint32_t* rng_mt624, int32_t* rng_index,
                          ipp_enhance_item* items_out, uint8_t* luts_out, double* draws_out);
int64_t ipp_enhance_workspace_bytes(int w, int h, int n, int n_blur);
/*    src / dst: n dense images (h x w x 3), may be the same buffer; items / luts / blur_list (indices of the images
 *    with blur == 1, n_blur of them) are DEVICE arrays. */
int ipp_enhance_batch_u8(const uint8_t* src, uint8_t* dst, int w, int h, int n, const ipp_enhance_item* items,
                         const uint8_t* luts, const int32_t* blur_list, int n_blur,
                         void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IPPGPU_H */
