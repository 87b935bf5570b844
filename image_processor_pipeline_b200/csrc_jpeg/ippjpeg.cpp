// ippjpeg.cpp — batched JPEG decode / encode on the GPU through nvJPEG behind a C ABI (include/ippjpeg.h).
#include "../../include/ippjpeg.h"

#include <cuda_runtime.h>
#include <nvjpeg.h>

#include <cstdarg>
#include <cstdio>
#include <vector>

namespace {

thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct Codec {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    nvjpegEncoderState_t enc_state = nullptr;
    nvjpegEncoderParams_t enc_params = nullptr;
    int backend = 0;
    int cpu_threads = 8;            // host threads of the Huffman stage when it runs on the CPU
    int batch_init = 0;             // batch size / format the batched decoder is initialised for
    int fmt_init = -1;
    std::vector<nvjpegImage_t> outs;
};

#define NVJ(call)                                                        \
    do {                                                                 \
        nvjpegStatus_t _s = (call);                                      \
        if (_s != NVJPEG_STATUS_SUCCESS) {                               \
            set_error("%s failed: nvjpegStatus %d", #call, (int)_s);     \
            return (int)_s;                                              \
        }                                                                \
    } while (0)

}  // namespace

extern "C" {

const char* ippjpeg_last_error_string(void) { return g_err; }

int ippjpeg_create(void** handle_out, int backend) {
    if (!handle_out) { set_error("ippjpeg_create: handle_out is NULL"); return IPPJPEG_ERR_BAD_ARG; }
    Codec* c = new Codec();
    nvjpegStatus_t s = NVJPEG_STATUS_NOT_INITIALIZED;
    // bit 8 of `backend`: chroma up-sampling with interpolation (closer to libjpeg-turbo's "fancy" up-sampling)
    const unsigned int flags = (backend & 256) ? NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION : NVJPEG_FLAGS_DEFAULT;
    backend &= 255;
    if (backend == 1) {
        s = nvjpegCreateEx(NVJPEG_BACKEND_HARDWARE, nullptr, nullptr, flags, &c->handle);
        if (s == NVJPEG_STATUS_SUCCESS) c->backend = 1;
    }
    if (s != NVJPEG_STATUS_SUCCESS && backend == 2) {
        s = nvjpegCreateEx(NVJPEG_BACKEND_GPU_HYBRID, nullptr, nullptr, flags, &c->handle);
        if (s == NVJPEG_STATUS_SUCCESS) c->backend = 2;
    }
    if (s != NVJPEG_STATUS_SUCCESS) {
        s = nvjpegCreateEx(NVJPEG_BACKEND_DEFAULT, nullptr, nullptr, flags, &c->handle);
        c->backend = 0;
    }
    if (s != NVJPEG_STATUS_SUCCESS) {
        set_error("nvjpegCreateEx failed: nvjpegStatus %d", (int)s);
        delete c;
        return (int)s;
    }
    s = nvjpegJpegStateCreate(c->handle, &c->state);
    if (s != NVJPEG_STATUS_SUCCESS) {
        set_error("nvjpegJpegStateCreate failed: nvjpegStatus %d", (int)s);
        nvjpegDestroy(c->handle);
        delete c;
        return (int)s;
    }
    *handle_out = c;
    return IPPJPEG_OK;
}

int ippjpeg_destroy(void* handle) {
    Codec* c = (Codec*)handle;
    if (!c) return IPPJPEG_OK;
    if (c->enc_params) nvjpegEncoderParamsDestroy(c->enc_params);
    if (c->enc_state) nvjpegEncoderStateDestroy(c->enc_state);
    if (c->state) nvjpegJpegStateDestroy(c->state);
    if (c->handle) nvjpegDestroy(c->handle);
    delete c;
    return IPPJPEG_OK;
}

int ippjpeg_backend(void* handle) { return handle ? ((Codec*)handle)->backend : -1; }

int ippjpeg_info(void* handle, const uint8_t* data, size_t size, int* width, int* height, int* components) {
    Codec* c = (Codec*)handle;
    if (!c || !data || !width || !height || !components) { set_error("ippjpeg_info: NULL argument"); return IPPJPEG_ERR_BAD_ARG; }
    int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t ss;
    NVJ(nvjpegGetImageInfo(c->handle, data, size, &nc, &ss, ws, hs));
    *width = ws[0]; *height = hs[0]; *components = nc;
    return IPPJPEG_OK;
}

int ippjpeg_decode_batch(void* handle, const uint8_t* const* data, const size_t* sizes, int n, int w, int h, int bgr,
                         uint8_t* dst, int64_t pitch, int64_t image_stride, void* stream) {
    Codec* c = (Codec*)handle;
    if (!c || !data || !sizes || !dst || n < 0) { set_error("ippjpeg_decode_batch: bad arguments"); return IPPJPEG_ERR_BAD_ARG; }
    if (w <= 0 || h <= 0 || pitch < (int64_t)w * 3) { set_error("ippjpeg_decode_batch: bad geometry"); return IPPJPEG_ERR_BAD_SHAPE; }
    if (n == 0) return IPPJPEG_OK;
    for (int i = 0; i < n; ++i) {
        int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
        nvjpegChromaSubsampling_t ss;
        NVJ(nvjpegGetImageInfo(c->handle, data[i], sizes[i], &nc, &ss, ws, hs));
        if (ws[0] != w || hs[0] != h || nc != 3) {
            set_error("ippjpeg_decode_batch: image %d is %dx%d with %d components, expected %dx%dx3", i, ws[0], hs[0], nc, w, h);
            return IPPJPEG_ERR_BAD_SHAPE;
        }
    }
    const nvjpegOutputFormat_t fmt = bgr ? NVJPEG_OUTPUT_BGRI : NVJPEG_OUTPUT_RGBI;
    if (c->batch_init != n || c->fmt_init != (int)fmt) {
        NVJ(nvjpegDecodeBatchedInitialize(c->handle, c->state, n, c->cpu_threads, fmt));
        c->batch_init = n;
        c->fmt_init = (int)fmt;
    }
    c->outs.resize(n);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < NVJPEG_MAX_COMPONENT; ++k) { c->outs[i].channel[k] = nullptr; c->outs[i].pitch[k] = 0; }
        c->outs[i].channel[0] = dst + (int64_t)i * image_stride;
        c->outs[i].pitch[0] = (size_t)pitch;
    }
    NVJ(nvjpegDecodeBatched(c->handle, c->state, data, sizes, c->outs.data(), (cudaStream_t)stream));
    return IPPJPEG_OK;
}

static int ensure_encoder(Codec* c, cudaStream_t st) {
    if (!c->enc_state) NVJ(nvjpegEncoderStateCreate(c->handle, &c->enc_state, st));
    if (!c->enc_params) {
        NVJ(nvjpegEncoderParamsCreate(c->handle, &c->enc_params, st));
        NVJ(nvjpegEncoderParamsSetSamplingFactors(c->enc_params, NVJPEG_CSS_420, st));
    }
    return IPPJPEG_OK;
}

size_t ippjpeg_encode_bound(void* handle, int w, int h) {
    Codec* c = (Codec*)handle;
    if (!c || w <= 0 || h <= 0) return 0;
    if (ensure_encoder(c, nullptr) != IPPJPEG_OK) return 0;
    size_t len = 0;
    if (nvjpegEncodeGetBufferSize(c->handle, c->enc_params, w, h, &len) != NVJPEG_STATUS_SUCCESS) return 0;
    return len;
}

int ippjpeg_encode(void* handle, const uint8_t* src, int64_t pitch, int w, int h, int bgr, int quality,
                   uint8_t* out_host, size_t capacity, size_t* size_out, void* stream) {
    Codec* c = (Codec*)handle;
    cudaStream_t st = (cudaStream_t)stream;
    if (!c || !src || !out_host || !size_out) { set_error("ippjpeg_encode: NULL argument"); return IPPJPEG_ERR_BAD_ARG; }
    if (w <= 0 || h <= 0 || pitch < (int64_t)w * 3 || quality < 1 || quality > 100) { set_error("ippjpeg_encode: bad geometry / quality"); return IPPJPEG_ERR_BAD_SHAPE; }
    int rc = ensure_encoder(c, st);
    if (rc != IPPJPEG_OK) return rc;
    NVJ(nvjpegEncoderParamsSetQuality(c->enc_params, quality, st));
    nvjpegImage_t img;
    for (int k = 0; k < NVJPEG_MAX_COMPONENT; ++k) { img.channel[k] = nullptr; img.pitch[k] = 0; }
    img.channel[0] = const_cast<unsigned char*>(src);
    img.pitch[0] = (size_t)pitch;
    NVJ(nvjpegEncodeImage(c->handle, c->enc_state, c->enc_params, &img, bgr ? NVJPEG_INPUT_BGRI : NVJPEG_INPUT_RGBI, w, h, st));
    size_t len = 0;
    NVJ(nvjpegEncodeRetrieveBitstream(c->handle, c->enc_state, nullptr, &len, st));
    if (len > capacity) { set_error("ippjpeg_encode: output buffer too small (%zu > %zu)", len, capacity); return IPPJPEG_ERR_BAD_ARG; }
    NVJ(nvjpegEncodeRetrieveBitstream(c->handle, c->enc_state, out_host, &len, st));
    if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("ippjpeg_encode: stream synchronisation failed"); return IPPJPEG_ERR_UNSUPPORTED; }
    *size_out = len;
    return IPPJPEG_OK;
}

}  // extern "C"
