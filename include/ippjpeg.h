/*
 * ippjpeg.h — C ABI of the on-GPU JPEG codec either side of the transform chain.
 *
 * Replaces the host codec calls the reference makes around the hot path for JPEG files:
 * cv2.imread of the frames video.py writes (reference transforms/video.py:44 writes `.jpg`,
 * recadrages.py:29 reads them) and cv2.imwrite / PIL save of `.jpg` outputs.  The codec itself is NVIDIA's
 * nvJPEG (libnvjpeg, shipped with the CUDA toolkit) — library code, like cuBLAS would be for a GEMM; this
 * layer only owns batching, buffers and the C boundary.  Built as its own shared library (libippjpeg.so) so
 * that libippgpu.so has no dependency on libnvjpeg.
 *
 * Conventions as in ippgpu.h: plain C, device pointers owned by the caller, `stream` = cudaStream_t as void*,
 * return 0 = ok, negative = IPPJPEG_ERR_*, positive = nvjpegStatus_t of a failed library call;
 * ippjpeg_last_error_string() describes the last failure of the calling thread.
 *
 * Parity: decoded pixels are compared with libjpeg-turbo's (cv2.imdecode) on the SAME bitstream; the two decoders
 * round the inverse DCT / colour conversion differently and upsample chroma differently, so equality is not
 * bit-exact: the measured tolerance is stated in tests/test_gpu_jpeg.py.
 */
#ifndef IPPJPEG_H
#define IPPJPEG_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define IPPJPEG_OK 0
#define IPPJPEG_ERR_BAD_ARG (-1)
#define IPPJPEG_ERR_BAD_SHAPE (-2)   /* an image of the batch does not have the announced size / is not 3-component */
#define IPPJPEG_ERR_UNSUPPORTED (-4)

const char* ippjpeg_last_error_string(void);

/* backend (low byte): 0 = nvJPEG default, 1 = the hardware JPEG engine, 2 = GPU-assisted Huffman decode (nvJPEG's
 * GPU_HYBRID, for batches of baseline bitstreams); an unavailable backend falls back to 0 and ippjpeg_backend() tells
 * which one is active.  Bit 8 (| 256): chroma up-sampling with interpolation (closer to libjpeg-turbo's default). */
int ippjpeg_create(void** handle_out, int backend);
int ippjpeg_destroy(void* handle);
int ippjpeg_backend(void* handle);

/* width / height / components of one JPEG bitstream (host memory) */
int ippjpeg_info(void* handle, const uint8_t* data_host, size_t size, int* width, int* height, int* components);

/* Decode n baseline JPEG bitstreams (host memory) of identical size w x h into interleaved 3-channel images
 * dst + i * image_stride (row pitch in bytes), byte order BGR (cv2) when bgr != 0 else RGB (PIL).  Asynchronous on
 * `stream` (the Huffman stage of the default backend may run partly on the calling thread). */
int ippjpeg_decode_batch(void* handle, const uint8_t* const* data_host, const size_t* sizes_host, int n, int w, int h, int bgr,
                         uint8_t* dst, int64_t pitch, int64_t image_stride, void* stream);

/* Encode one interleaved 3-channel device image (BGR when bgr != 0) to a baseline JPEG, chroma 4:2:0 like cv2.imwrite's
 * default, quality 1..100.  out_host must hold ippjpeg_encode_bound(w, h) bytes; *size_out = bytes written.
 * Synchronises `stream`. */
size_t ippjpeg_encode_bound(void* handle, int w, int h);
int ippjpeg_encode(void* handle, const uint8_t* src, int64_t pitch, int w, int h, int bgr, int quality,
                   uint8_t* out_host, size_t capacity, size_t* size_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
