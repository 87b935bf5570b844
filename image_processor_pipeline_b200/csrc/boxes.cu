// boxes.cu — YOLO box arithmetic on the device, one warp per image.
//
// Reference behaviour replaced (float32 store semantics of ultralytics.utils.ops,
// float64 expression evaluation — SURVEY.md §8c):
//   transforms/crop_square.py:167      xywhn2xyxy(bboxes, width, height)     -> float32
//   transforms/crop_square.py:201-208  shift by (x0,y0), clip to [0, cs]     -> float64
//   transforms/crop_square.py:210      valid = x1 < x2 & y1 < y2
//   transforms/crop_square.py:217      xyxy2xywhn(valid boxes, cs, cs)       -> float32
//   transforms/overlays.py:143-146     xyxy2xywhn(int box, bg_w, bg_h)       -> float32
#include "common.cuh"

namespace ipp {

__device__ __forceinline__ double clipd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

__global__ void __launch_bounds__(128) boxes_square_crop_kernel(const double* __restrict__ boxes_in,
                                                                const int32_t* __restrict__ offs,
                                                                const int32_t* __restrict__ params,
                                                                float* __restrict__ boxes_out,
                                                                int32_t* __restrict__ valid_out, int n_images) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_images) return;
    const double W = (double)params[5 * warp + 0], H = (double)params[5 * warp + 1];
    const double x0 = (double)params[5 * warp + 2], y0 = (double)params[5 * warp + 3];
    const double cs = (double)params[5 * warp + 4];
    for (int i = offs[warp] + lane; i < offs[warp + 1]; i += 32) {
        const double xc = boxes_in[4 * i + 0], yc = boxes_in[4 * i + 1];
        const double hw = __ddiv_rn(boxes_in[4 * i + 2], 2.0), hh = __ddiv_rn(boxes_in[4 * i + 3], 2.0);
        // xywhn2xyxy: float64 expression, float32 store
        const float ax1 = (float)__dmul_rn(W, __dsub_rn(xc, hw));
        const float ay1 = (float)__dmul_rn(H, __dsub_rn(yc, hh));
        const float ax2 = (float)__dmul_rn(W, __dadd_rn(xc, hw));
        const float ay2 = (float)__dmul_rn(H, __dadd_rn(yc, hh));
        // float32 - int64 -> float64; clip
        const double x1 = clipd(__dsub_rn((double)ax1, x0), 0.0, cs), y1 = clipd(__dsub_rn((double)ay1, y0), 0.0, cs);
        const double x2 = clipd(__dsub_rn((double)ax2, x0), 0.0, cs), y2 = clipd(__dsub_rn((double)ay2, y0), 0.0, cs);
        valid_out[i] = (x1 < x2) && (y1 < y2);
        boxes_out[4 * i + 0] = (float)__ddiv_rn(__ddiv_rn(__dadd_rn(x1, x2), 2.0), cs);
        boxes_out[4 * i + 1] = (float)__ddiv_rn(__ddiv_rn(__dadd_rn(y1, y2), 2.0), cs);
        boxes_out[4 * i + 2] = (float)__ddiv_rn(__dsub_rn(x2, x1), cs);
        boxes_out[4 * i + 3] = (float)__ddiv_rn(__dsub_rn(y2, y1), cs);
    }
}

__global__ void __launch_bounds__(128) boxes_xyxy2xywhn_kernel(const int32_t* __restrict__ xyxy,
                                                               const int32_t* __restrict__ wh,
                                                               float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double W = (double)wh[2 * i], H = (double)wh[2 * i + 1];
    const long long x1 = xyxy[4 * i], y1 = xyxy[4 * i + 1], x2 = xyxy[4 * i + 2], y2 = xyxy[4 * i + 3];
    out[4 * i + 0] = (float)__ddiv_rn(__ddiv_rn((double)(x1 + x2), 2.0), W);
    out[4 * i + 1] = (float)__ddiv_rn(__ddiv_rn((double)(y1 + y2), 2.0), H);
    out[4 * i + 2] = (float)__ddiv_rn((double)(x2 - x1), W);
    out[4 * i + 3] = (float)__ddiv_rn((double)(y2 - y1), H);
}

}  // namespace ipp

extern "C" {

int ipp_boxes_square_crop_f32(const double* boxes_in, const int32_t* box_offsets, const int32_t* params,
                              float* boxes_out, int32_t* valid_out, int n_images, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(boxes_in && box_offsets && params && boxes_out && valid_out, IPP_ERR_BAD_ARG,
                "ipp_boxes_square_crop_f32: NULL pointer");
    IPP_REQUIRE(n_images >= 0, IPP_ERR_BAD_ARG, "ipp_boxes_square_crop_f32: n_images=%d", n_images);
    if (n_images == 0) return IPP_OK;
    { ipp::LaunchTimer _lt("boxes_square_crop_kernel", (cudaStream_t)stream); boxes_square_crop_kernel<<<ceil_div((int64_t)n_images * 32, 128), 128, 0, (cudaStream_t)stream>>>(
        boxes_in, box_offsets, params, boxes_out, valid_out, n_images); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

int ipp_boxes_xyxy2xywhn_f32(const int32_t* xyxy, const int32_t* img_wh, float* xywhn_out, int n_boxes, void* stream) {
    using namespace ipp;
    IPP_REQUIRE(xyxy && img_wh && xywhn_out, IPP_ERR_BAD_ARG, "ipp_boxes_xyxy2xywhn_f32: NULL pointer");
    IPP_REQUIRE(n_boxes >= 0, IPP_ERR_BAD_ARG, "ipp_boxes_xyxy2xywhn_f32: n_boxes=%d", n_boxes);
    if (n_boxes == 0) return IPP_OK;
    { ipp::LaunchTimer _lt("boxes_xyxy2xywhn_kernel", (cudaStream_t)stream); boxes_xyxy2xywhn_kernel<<<ceil_div(n_boxes, 128), 128, 0, (cudaStream_t)stream>>>(xyxy, img_wh, xywhn_out, n_boxes); }
    IPP_LAUNCH_CHECK();
    return IPP_OK;
}

}  // extern "C"
