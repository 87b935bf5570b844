// ippvideo.cu — NVDEC frame source behind include/ippvideo.h (reference transforms/video.py:23-45).
//
// libnvcuvid is part of the driver, not of the toolkit, and this image carries neither its headers nor a link stub: the
// library is opened with dlopen and the handful of cuvid structures used here are declared below, field for field, from the
// public Video Codec SDK interface (nvcuvid.h / cuviddec.h, API 11-12; the layouts have been append-only since API 8).
// Every structure passed in is zero-filled inside an oversized buffer, so that a library built against a newer header
// that appended fields reads zeros there.
//
// Flow of one packet: cuvidParseVideoData -> (sequence callback: create the decoder for the stream's size) ->
// (decode callback: cuvidDecodePicture) -> (display callback: map the NV12 surface, convert it to interleaved BGR
// in the caller's buffer with the kernel below, unmap).  All callbacks run synchronously inside the parse call.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/ippvideo.h"

namespace {

thread_local char g_err[512] = "";
void set_err(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------ cuvid interface (subset)
typedef int CUresult_t;                 // CUresult
typedef void* CUvideoparser;
typedef void* CUvideodecoder;
typedef void* CUvideoctxlock;
typedef long long CUvideotimestamp;

struct CUVIDEOFORMAT {
    int codec;
    struct { unsigned int numerator, denominator; } frame_rate;
    unsigned char progressive_sequence, bit_depth_luma_minus8, bit_depth_chroma_minus8, min_num_decode_surfaces;
    unsigned int coded_width, coded_height;
    struct { int left, top, right, bottom; } display_area;
    int chroma_format;
    unsigned int bitrate;
    struct { int x, y; } display_aspect_ratio;
    struct {
        unsigned char video_format : 3;
        unsigned char video_full_range_flag : 1;
        unsigned char reserved_zero_bits : 4;
        unsigned char color_primaries, transfer_characteristics, matrix_coefficients;
    } video_signal_description;
    unsigned int seqhdr_data_length;
};
static_assert(sizeof(CUVIDEOFORMAT) == 64, "CUVIDEOFORMAT layout");

struct CUVIDPARSERDISPINFO {
    int picture_index, progressive_frame, top_field_first, repeat_first_field;
    CUvideotimestamp timestamp;
};

typedef int (*PFNVIDSEQUENCECALLBACK)(void*, CUVIDEOFORMAT*);
typedef int (*PFNVIDDECODECALLBACK)(void*, void* /* CUVIDPICPARAMS* */);
typedef int (*PFNVIDDISPLAYCALLBACK)(void*, CUVIDPARSERDISPINFO*);

struct CUVIDPARSERPARAMS {
    int CodecType;
    unsigned int ulMaxNumDecodeSurfaces, ulClockRate, ulErrorThreshold, ulMaxDisplayDelay;
    unsigned int bAnnexb : 1;
    unsigned int uReserved : 31;
    unsigned int uReserved1[4];
    void* pUserData;
    PFNVIDSEQUENCECALLBACK pfnSequenceCallback;
    PFNVIDDECODECALLBACK pfnDecodePicture;
    PFNVIDDISPLAYCALLBACK pfnDisplayPicture;
    void* pfnGetOperatingPoint;
    void* pfnGetSEIMsg;
    void* pvReserved2[5];
    void* pExtVideoInfo;
};
static_assert(offsetof(CUVIDPARSERPARAMS, pUserData) == 40 && sizeof(CUVIDPARSERPARAMS) == 136, "CUVIDPARSERPARAMS layout");

struct CUVIDSOURCEDATAPACKET {
    unsigned long flags, payload_size;
    const unsigned char* payload;
    CUvideotimestamp timestamp;
};
constexpr unsigned long kPktEndOfStream = 0x01, kPktEndOfPicture = 0x08;

struct CUVIDDECODECREATEINFO {
    unsigned long ulWidth, ulHeight, ulNumDecodeSurfaces;
    int CodecType, ChromaFormat;
    unsigned long ulCreationFlags, bitDepthMinus8, ulIntraDecodeOnly, ulMaxWidth, ulMaxHeight, Reserved1;
    struct { short left, top, right, bottom; } display_area;
    int OutputFormat, DeinterlaceMode;
    unsigned long ulTargetWidth, ulTargetHeight, ulNumOutputSurfaces;
    CUvideoctxlock vidLock;
    struct { short left, top, right, bottom; } target_rect;
    unsigned long enableHistogram;
    unsigned long Reserved2[4];
};
static_assert(offsetof(CUVIDDECODECREATEINFO, display_area) == 80 && offsetof(CUVIDDECODECREATEINFO, vidLock) == 120, "CUVIDDECODECREATEINFO layout");

struct CUVIDPROCPARAMS {
    int progressive_frame, second_field, top_field_first, unpaired_field;
    unsigned int reserved_flags, reserved_zero;
    unsigned long long raw_input_dptr;
    unsigned int raw_input_pitch, raw_input_format;
    unsigned long long raw_output_dptr;
    unsigned int raw_output_pitch, Reserved1;
    void* output_stream;
    unsigned int Reserved[46];
    unsigned long long* histogram_dptr;
    void* Reserved2[1];
};
static_assert(offsetof(CUVIDPROCPARAMS, output_stream) == 56, "CUVIDPROCPARAMS layout");

struct CUVIDDECODECAPS {
    int eCodecType, eChromaFormat;
    unsigned int nBitDepthMinus8;
    unsigned int reserved1[3];
    unsigned char bIsSupported, nNumNVDECs;
    unsigned short nOutputFormatMask;
    unsigned int nMaxWidth, nMaxHeight, nMaxMBCount;
    unsigned short nMinWidth, nMinHeight;
    unsigned char bIsHistogramSupported, nCounterBitDepth;
    unsigned short nMaxHistogramBins;
    unsigned int reserved3[10];
};

struct Api {
    void* so = nullptr;
    CUresult_t (*CreateVideoParser)(CUvideoparser*, void*) = nullptr;
    CUresult_t (*ParseVideoData)(CUvideoparser, CUVIDSOURCEDATAPACKET*) = nullptr;
    CUresult_t (*DestroyVideoParser)(CUvideoparser) = nullptr;
    CUresult_t (*CreateDecoder)(CUvideodecoder*, void*) = nullptr;
    CUresult_t (*DestroyDecoder)(CUvideodecoder) = nullptr;
    CUresult_t (*DecodePicture)(CUvideodecoder, void*) = nullptr;
    CUresult_t (*MapVideoFrame64)(CUvideodecoder, int, unsigned long long*, unsigned int*, void*) = nullptr;
    CUresult_t (*UnmapVideoFrame64)(CUvideodecoder, unsigned long long) = nullptr;
    CUresult_t (*GetDecoderCaps)(void*) = nullptr;
    bool ok = false;
};

Api& api() {
    static Api a = [] {
        Api x;
        for (const char* name : {"libnvcuvid.so.1", "libnvcuvid.so"}) {
            x.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (x.so) break;
        }
        if (!x.so) return x;
        auto sym = [&](const char* n) { return dlsym(x.so, n); };
        x.CreateVideoParser = (decltype(x.CreateVideoParser))sym("cuvidCreateVideoParser");
        x.ParseVideoData = (decltype(x.ParseVideoData))sym("cuvidParseVideoData");
        x.DestroyVideoParser = (decltype(x.DestroyVideoParser))sym("cuvidDestroyVideoParser");
        x.CreateDecoder = (decltype(x.CreateDecoder))sym("cuvidCreateDecoder");
        x.DestroyDecoder = (decltype(x.DestroyDecoder))sym("cuvidDestroyDecoder");
        x.DecodePicture = (decltype(x.DecodePicture))sym("cuvidDecodePicture");
        x.MapVideoFrame64 = (decltype(x.MapVideoFrame64))sym("cuvidMapVideoFrame64");
        x.UnmapVideoFrame64 = (decltype(x.UnmapVideoFrame64))sym("cuvidUnmapVideoFrame64");
        x.GetDecoderCaps = (decltype(x.GetDecoderCaps))sym("cuvidGetDecoderCaps");
        x.ok = x.CreateVideoParser && x.ParseVideoData && x.DestroyVideoParser && x.CreateDecoder && x.DestroyDecoder && x.DecodePicture &&
               x.MapVideoFrame64 && x.UnmapVideoFrame64;
        return x;
    }();
    return a;
}

// zero-filled, oversized home of a structure handed to the library
template <typename T>
struct Padded {
    alignas(16) unsigned char raw[sizeof(T) + 512];
    Padded() { memset(raw, 0, sizeof(raw)); }
    T* operator->() { return reinterpret_cast<T*>(raw); }
    void* ptr() { return raw; }
};

// ------------------------------------------------------------------ NV12 -> interleaved BGR
// BT.601, limited range, integer: C = Y - 16, D = U - 128, E = V - 128;
// R = clip((298 C + 409 E + 128) >> 8), G = clip((298 C - 100 D - 208 E + 128) >> 8), B = clip((298 C + 516 D + 128) >> 8);
// chroma of a 2 x 2 block from its one sample (nearest).  thread = 4 pixels of a row: 4 luma bytes + 2 chroma pairs in,
// 12 bytes out (three aligned words when the row allows it).
__device__ __forceinline__ uint32_t clip8i(int v) { return (uint32_t)min(max(v >> 8, 0), 255); }
__global__ void __launch_bounds__(256) nv12_to_bgr_kernel(const uint8_t* __restrict__ luma, const uint8_t* __restrict__ chroma, int src_pitch,
                                                          int x0, int y0, int w, int h, uint8_t* __restrict__ dst, int64_t dst_pitch) {
    const int q = (w + 3) >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= q * h) return;
    const int y = idx / q, x = (idx - y * q) * 4;
    const int sy = y0 + y;
    uint8_t o[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int sx = x0 + min(x + k, w - 1);
        const int C = (int)luma[(int64_t)sy * src_pitch + sx] - 16;
        const uint8_t* uv = chroma + (int64_t)(sy >> 1) * src_pitch + (sx & ~1);
        const int D = (int)uv[0] - 128, E = (int)uv[1] - 128;
        o[3 * k + 0] = (uint8_t)clip8i(298 * C + 516 * D + 128);
        o[3 * k + 1] = (uint8_t)clip8i(298 * C - 100 * D - 208 * E + 128);
        o[3 * k + 2] = (uint8_t)clip8i(298 * C + 409 * E + 128);
    }
    uint8_t* d = dst + (int64_t)y * dst_pitch + (int64_t)x * 3;
    if (x + 3 < w && ((((uintptr_t)d) & 3) == 0)) {
        uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
        d32[0] = o[0] | (o[1] << 8) | (o[2] << 16) | ((uint32_t)o[3] << 24);
        d32[1] = o[4] | (o[5] << 8) | (o[6] << 16) | ((uint32_t)o[7] << 24);
        d32[2] = o[8] | (o[9] << 8) | (o[10] << 16) | ((uint32_t)o[11] << 24);
    } else {
        for (int k = 0; k < 12 && x + k / 3 < w; ++k) d[k] = o[k];
    }
}

// ------------------------------------------------------------------ decoder object
struct Decoder {
    int codec = 0;
    CUvideoparser parser = nullptr;
    CUvideodecoder decoder = nullptr;
    int width = 0, height = 0;                 // display size
    int disp_left = 0, disp_top = 0;
    unsigned int surface_h = 0;                // rows of the luma plane of a mapped surface (chroma follows it)
    // output of the call in progress
    uint8_t* out = nullptr;
    int64_t pitch = 0, frame_stride = 0;
    int capacity = 0, produced = 0;
    cudaStream_t stream = nullptr;
    int error = 0;                             // first failure inside a callback of the call in progress
};

int on_sequence(void* user, CUVIDEOFORMAT* f) {
    Decoder* d = static_cast<Decoder*>(user);
    Api& a = api();
    const int w = f->display_area.right - f->display_area.left, h = f->display_area.bottom - f->display_area.top;
    if (f->bit_depth_luma_minus8 != 0 || f->chroma_format != 1 /* 4:2:0 */) {
        set_err("ippvideo: only 8-bit 4:2:0 streams are converted (bit depth %d, chroma format %d)", 8 + f->bit_depth_luma_minus8, f->chroma_format);
        d->error = IPPVIDEO_ERR_UNSUPPORTED;
        return 0;
    }
    if (d->decoder) {
        if (w == d->width && h == d->height) return f->min_num_decode_surfaces > 0 ? f->min_num_decode_surfaces : 1;
        a.DestroyDecoder(d->decoder);          // a new sequence of another size: start over
        d->decoder = nullptr;
    }
    int surfaces = f->min_num_decode_surfaces > 0 ? f->min_num_decode_surfaces : 8;
    if (surfaces < 4) surfaces = 4;
    Padded<CUVIDDECODECREATEINFO> ci;
    ci->ulWidth = f->coded_width;
    ci->ulHeight = f->coded_height;
    ci->ulNumDecodeSurfaces = (unsigned long)surfaces;
    ci->CodecType = f->codec;
    ci->ChromaFormat = f->chroma_format;
    ci->ulCreationFlags = 4;                   // cudaVideoCreate_PreferCUVID: the dedicated engine
    ci->bitDepthMinus8 = 0;
    ci->ulMaxWidth = f->coded_width;
    ci->ulMaxHeight = f->coded_height;
    ci->display_area.left = 0; ci->display_area.top = 0;
    ci->display_area.right = (short)f->coded_width; ci->display_area.bottom = (short)f->coded_height;
    ci->OutputFormat = 0;                      // cudaVideoSurfaceFormat_NV12
    ci->DeinterlaceMode = f->progressive_sequence ? 0 /* weave */ : 2 /* adaptive */;
    ci->ulTargetWidth = f->coded_width;
    ci->ulTargetHeight = f->coded_height;
    ci->ulNumOutputSurfaces = 2;
    ci->vidLock = nullptr;
    const CUresult_t rc = a.CreateDecoder(&d->decoder, ci.ptr());
    if (rc != 0) {
        set_err("cuvidCreateDecoder failed for a %u x %u stream of codec %d: CUresult %d", f->coded_width, f->coded_height, f->codec, rc);
        d->error = rc;
        d->decoder = nullptr;
        return 0;
    }
    d->width = w; d->height = h;
    d->disp_left = f->display_area.left; d->disp_top = f->display_area.top;
    d->surface_h = f->coded_height;
    return surfaces;
}

int on_decode(void* user, void* pic_params) {
    Decoder* d = static_cast<Decoder*>(user);
    if (!d->decoder) return 0;
    const CUresult_t rc = api().DecodePicture(d->decoder, pic_params);
    if (rc != 0) {
        if (!d->error) { set_err("cuvidDecodePicture: CUresult %d", rc); d->error = rc; }
        return 0;
    }
    return 1;
}

int on_display(void* user, CUVIDPARSERDISPINFO* info) {
    Decoder* d = static_cast<Decoder*>(user);
    if (!info || !d->decoder) return 1;        // end-of-stream notification
    Api& a = api();
    if (d->produced >= d->capacity || d->out == nullptr) {
        if (!d->error) { set_err("ippvideo_decode: more than %d frames became displayable", d->capacity); d->error = IPPVIDEO_ERR_CAPACITY; }
        return 0;
    }
    if (d->pitch < (int64_t)d->width * 3) {
        if (!d->error) { set_err("ippvideo_decode: pitch %lld smaller than a row of %d pixels", (long long)d->pitch, d->width); d->error = IPPVIDEO_ERR_BAD_ARG; }
        return 0;
    }
    Padded<CUVIDPROCPARAMS> pp;
    pp->progressive_frame = info->progressive_frame;
    pp->second_field = info->repeat_first_field + 1;
    pp->top_field_first = info->top_field_first;
    pp->unpaired_field = info->repeat_first_field < 0;
    pp->output_stream = d->stream;
    unsigned long long src = 0;
    unsigned int src_pitch = 0;
    CUresult_t rc = a.MapVideoFrame64(d->decoder, info->picture_index, &src, &src_pitch, pp.ptr());
    if (rc != 0) {
        if (!d->error) { set_err("cuvidMapVideoFrame64: CUresult %d", rc); d->error = rc; }
        return 0;
    }
    const uint8_t* luma = reinterpret_cast<const uint8_t*>(src);
    const uint8_t* chroma = luma + (size_t)src_pitch * ((d->surface_h + 1u) & ~1u);
    const int q = (d->width + 3) >> 2;
    const int threads = q * d->height;
    nv12_to_bgr_kernel<<<(threads + 255) / 256, 256, 0, d->stream>>>(luma, chroma, (int)src_pitch, d->disp_left, d->disp_top, d->width, d->height,
                                                                    d->out + (int64_t)d->produced * d->frame_stride, d->pitch);
    cudaError_t ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(d->stream);      // the surface goes back to the decoder below
    rc = a.UnmapVideoFrame64(d->decoder, src);
    if (ce != cudaSuccess) {
        if (!d->error) { set_err("nv12_to_bgr_kernel: %s", cudaGetErrorString(ce)); d->error = 1000 + (int)ce; }
        return 0;
    }
    if (rc != 0) {
        if (!d->error) { set_err("cuvidUnmapVideoFrame64: CUresult %d", rc); d->error = rc; }
        return 0;
    }
    ++d->produced;
    return 1;
}

}  // namespace

extern "C" {

const char* ippvideo_last_error_string(void) { return g_err; }

int ippvideo_supported(int codec, int* max_w, int* max_h) {
    if (max_w) *max_w = 0;
    if (max_h) *max_h = 0;
    Api& a = api();
    if (!a.ok) { set_err("ippvideo: libnvcuvid.so.1 cannot be opened (%s)", dlerror() ? dlerror() : "missing symbols"); return 0; }
    if (cudaFree(nullptr) != cudaSuccess) { set_err("ippvideo: no CUDA context"); return 0; }
    if (!a.GetDecoderCaps) return 1;           // old driver: find out by creating the decoder
    Padded<CUVIDDECODECAPS> caps;
    caps->eCodecType = codec;
    caps->eChromaFormat = 1;
    caps->nBitDepthMinus8 = 0;
    const CUresult_t rc = a.GetDecoderCaps(caps.ptr());
    if (rc != 0) { set_err("cuvidGetDecoderCaps: CUresult %d", rc); return 0; }
    if (max_w) *max_w = (int)caps->nMaxWidth;
    if (max_h) *max_h = (int)caps->nMaxHeight;
    if (!caps->bIsSupported) set_err("ippvideo: this GPU's NVDEC does not decode codec %d (8-bit 4:2:0)", codec);
    return caps->bIsSupported ? 1 : 0;
}

int ippvideo_create(void** handle_out, int codec) {
    if (!handle_out) { set_err("ippvideo_create: NULL handle"); return IPPVIDEO_ERR_BAD_ARG; }
    *handle_out = nullptr;
    Api& a = api();
    if (!a.ok) { set_err("ippvideo_create: libnvcuvid.so.1 cannot be opened on this machine"); return IPPVIDEO_ERR_UNAVAILABLE; }
    if (cudaFree(nullptr) != cudaSuccess) { set_err("ippvideo_create: no CUDA context"); return IPPVIDEO_ERR_UNAVAILABLE; }
    if (!ippvideo_supported(codec, nullptr, nullptr)) return IPPVIDEO_ERR_UNAVAILABLE;
    Decoder* d = new Decoder();
    d->codec = codec;
    Padded<CUVIDPARSERPARAMS> pp;
    pp->CodecType = codec;
    pp->ulMaxNumDecodeSurfaces = 1;            // the sequence callback returns the real number
    pp->ulClockRate = 0;
    pp->ulErrorThreshold = 0;
    pp->ulMaxDisplayDelay = 0;                 // frames are displayed as soon as they are decodable in order
    pp->pUserData = d;
    pp->pfnSequenceCallback = on_sequence;
    pp->pfnDecodePicture = on_decode;
    pp->pfnDisplayPicture = on_display;
    const CUresult_t rc = a.CreateVideoParser(&d->parser, pp.ptr());
    if (rc != 0) {
        set_err("cuvidCreateVideoParser (codec %d): CUresult %d", codec, rc);
        delete d;
        return rc > 0 ? rc : IPPVIDEO_ERR_UNAVAILABLE;
    }
    *handle_out = d;
    return IPPVIDEO_OK;
}

int ippvideo_destroy(void* handle) {
    Decoder* d = static_cast<Decoder*>(handle);
    if (!d) return IPPVIDEO_OK;
    Api& a = api();
    if (d->parser) a.DestroyVideoParser(d->parser);
    if (d->decoder) a.DestroyDecoder(d->decoder);
    delete d;
    return IPPVIDEO_OK;
}

int ippvideo_size(void* handle, int* width, int* height) {
    Decoder* d = static_cast<Decoder*>(handle);
    if (!d) { set_err("ippvideo_size: NULL handle"); return IPPVIDEO_ERR_BAD_ARG; }
    if (width) *width = d->width;
    if (height) *height = d->height;
    return IPPVIDEO_OK;
}

int ippvideo_decode(void* handle, const uint8_t* packet_host, size_t size, int end_of_stream,
                    uint8_t* out, int64_t pitch, int64_t frame_stride, int capacity_frames,
                    int* n_frames_out, int* width, int* height, void* stream) {
    Decoder* d = static_cast<Decoder*>(handle);
    if (n_frames_out) *n_frames_out = 0;
    if (!d || (size > 0 && !packet_host) || capacity_frames < 0 || (capacity_frames > 0 && !out)) {
        set_err("ippvideo_decode: bad arguments");
        return IPPVIDEO_ERR_BAD_ARG;
    }
    d->out = out; d->pitch = pitch; d->frame_stride = frame_stride; d->capacity = capacity_frames; d->produced = 0;
    d->stream = static_cast<cudaStream_t>(stream);
    d->error = 0;
    CUVIDSOURCEDATAPACKET pkt;
    memset(&pkt, 0, sizeof(pkt));
    pkt.payload = packet_host;
    pkt.payload_size = (unsigned long)size;
    pkt.flags = kPktEndOfPicture;              // the demuxer hands out whole access units
    if (end_of_stream || size == 0) pkt.flags |= kPktEndOfStream;
    const CUresult_t rc = api().ParseVideoData(d->parser, &pkt);
    if (n_frames_out) *n_frames_out = d->produced;
    if (width) *width = d->width;
    if (height) *height = d->height;
    d->out = nullptr; d->capacity = 0;
    if (d->error) return d->error;
    if (rc != 0) { set_err("cuvidParseVideoData: CUresult %d", rc); return rc; }
    return IPPVIDEO_OK;
}

}  // extern "C"
