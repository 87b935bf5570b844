/*
 * ippvideo.h — C ABI of the on-GPU video frame source in front of the transform chain.
 *
 * Replaces the host decode of reference transforms/video.py:23-45 (cv2.VideoCapture.read() of every frame, then
 * cv2.imwrite of a .jpg that the next step reads back): the compressed packets of the video cross PCIe, NVDEC decodes
 * them, and the frames are born in HBM as the interleaved BGR images the chain's stage A consumes.  The decoder is
 * NVIDIA's NVDEC behind the driver's libnvcuvid (library code, like nvJPEG for the JPEG layer); this layer owns the
 * parser / decoder plumbing, the NV12 -> BGR conversion kernel and the C boundary.  libnvcuvid ships with the driver,
 * not with the toolkit: it is opened at run time (dlopen), so this library builds and loads without it and
 * ippvideo_create reports IPPVIDEO_ERR_UNAVAILABLE where it is missing.
 *
 * Demuxing stays with whoever owns the container (cv2.VideoCapture with CAP_PROP_FORMAT = -1 hands out the raw packets
 * of a file; Python: nvdec.frame_batches).
 *
 * Conventions as in ippgpu.h: plain C, device pointers owned by the caller, `stream` = cudaStream_t as void*,
 * return 0 = ok, negative = IPPVIDEO_ERR_*, positive = CUresult of a failed driver / cuvid call;
 * ippvideo_last_error_string() describes the last failure of the calling thread.
 *
 * Parity: the decoded Y / U / V samples of a conforming bitstream are normative (identical to libvpx / libavcodec);
 * the reference's BGR frames come out of swscale's yuv420p -> bgr24 conversion (BT.601, limited range, its own chroma
 * up-sampling and rounding), which is not reproduced bit for bit: the conversion here is the integer BT.601 limited-range
 * matrix with nearest chroma, and the measured tolerance against cv2.VideoCapture on the same file is stated in
 * tests/test_gpu_video.py.
 */
#ifndef IPPVIDEO_H
#define IPPVIDEO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define IPPVIDEO_OK 0
#define IPPVIDEO_ERR_BAD_ARG (-1)
#define IPPVIDEO_ERR_UNAVAILABLE (-2)   /* no libnvcuvid on this machine, or the GPU cannot decode this codec / size */
#define IPPVIDEO_ERR_CAPACITY (-3)      /* more frames became displayable than the output buffer holds */
#define IPPVIDEO_ERR_UNSUPPORTED (-4)   /* bit depth / chroma format the conversion does not handle (8-bit 4:2:0 only) */

/* cudaVideoCodec values of the cuvid API */
#define IPPVIDEO_CODEC_MPEG1 0
#define IPPVIDEO_CODEC_MPEG2 1
#define IPPVIDEO_CODEC_MPEG4 2
#define IPPVIDEO_CODEC_H264 4
#define IPPVIDEO_CODEC_HEVC 8
#define IPPVIDEO_CODEC_VP8 9
#define IPPVIDEO_CODEC_VP9 10
#define IPPVIDEO_CODEC_AV1 11

const char* ippvideo_last_error_string(void);

/* 1 when libnvcuvid can be opened and this GPU decodes `codec` (8-bit 4:2:0), else 0; max_w / max_h (optional) receive
 * the largest supported frame. */
int ippvideo_supported(int codec, int* max_w, int* max_h);

/* One parser + decoder for one elementary stream of `codec` on the current device / context. */
int ippvideo_create(void** handle_out, int codec);
int ippvideo_destroy(void* handle);

/* Feed one packet (host memory; size 0 with end_of_stream != 0 flushes the decoder).  Every frame that becomes displayable
 * during the call is converted to interleaved BGR (cv2 order) and written to out + k * frame_stride (row pitch `pitch`
 * bytes, k = 0 .. *n_frames_out - 1, at most capacity_frames).  *width / *height = display size of the stream (0 before
 * the first sequence header).  The frames are complete when the call returns (it synchronises `stream`). */
int ippvideo_decode(void* handle, const uint8_t* packet_host, size_t size, int end_of_stream,
                    uint8_t* out, int64_t pitch, int64_t frame_stride, int capacity_frames,
                    int* n_frames_out, int* width, int* height, void* stream);

/* display size of the stream (0, 0 before the first sequence header) */
int ippvideo_size(void* handle, int* width, int* height);

#ifdef __cplusplus
}
#endif
#endif
