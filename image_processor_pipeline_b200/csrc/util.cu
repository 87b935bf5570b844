// util.cu — error plumbing, version, device queries of libippgpu.
#include "common.cuh"

namespace ipp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace ipp

extern "C" {

int ipp_version(void) { return IPP_VERSION; }

const char* ipp_last_error_string(void) { return ipp::g_err; }

int ipp_sm_count(int* out_host) {
    IPP_REQUIRE(out_host != nullptr, IPP_ERR_BAD_ARG, "ipp_sm_count: out is NULL");
    int dev = 0;
    IPP_CUDA(cudaGetDevice(&dev));
    IPP_CUDA(cudaDeviceGetAttribute(out_host, cudaDevAttrMultiProcessorCount, dev));
    return IPP_OK;
}

}  // extern "C"
