// util.cu — error plumbing, version, device queries of libippgpu.
#include "common.cuh"

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <string.h>
#include <vector>

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

// ---------------------------------------------------------------- launch accounting

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_timing{0};
struct TimedLaunch {
    const char* name;
    cudaEvent_t a, b;
    cudaStream_t st;
};
static std::mutex g_tmu;
static std::vector<TimedLaunch> g_timed;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_event_pool;

LaunchTimer::LaunchTimer(const char* name, cudaStream_t st) : name_(name), st_(st), slot_(-1) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_timing.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lock(g_tmu);
    TimedLaunch t;
    t.name = name;
    if (!g_event_pool.empty()) {
        t.a = g_event_pool.back().first;
        t.b = g_event_pool.back().second;
        g_event_pool.pop_back();
    } else {
        if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) return;
    }
    t.st = st;
    cudaEventRecord(t.a, st);
    g_timed.push_back(t);
    slot_ = (int)g_timed.size() - 1;
}

LaunchTimer::~LaunchTimer() {
    if (slot_ < 0) return;
    std::lock_guard<std::mutex> lock(g_tmu);
    if (slot_ < (int)g_timed.size()) cudaEventRecord(g_timed[slot_].b, st_);
}

}  // namespace ipp

extern "C" {

long long ipp_launch_count(void) { return ipp::g_launches.load(); }

int ipp_timing_enable(int on) {
    ipp::g_timing.store(on ? 1 : 0);
    return IPP_OK;
}

// "name:launches:total_ms;..." for every kernel launched since the last report; synchronises.
int ipp_timing_report(char* buf_host, int buflen) {
    using namespace ipp;
    IPP_REQUIRE(buf_host && buflen > 0, IPP_ERR_BAD_ARG, "ipp_timing_report: bad buffer");
    std::lock_guard<std::mutex> lock(g_tmu);
    std::map<std::string, std::pair<int, double>> acc;
    for (auto& t : g_timed) {
        float ms = 0.f;
        if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            auto& e = acc[t.name];
            e.first += 1;
            e.second += ms;
        }
        g_event_pool.push_back({t.a, t.b});
    }
    g_timed.clear();
    std::string out;
    for (auto& kv : acc) {
        char line[256];
        snprintf(line, sizeof(line), "%s:%d:%.6f;", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    snprintf(buf_host, buflen, "%s", out.c_str());
    return IPP_OK;
}

// "name:stream:start_ms:end_ms;..." per launch since the last report (times relative to the first launch's
// start event; events of different streams share the device clock), in launch order; synchronises and clears.
int ipp_timing_timeline(char* buf_host, int buflen) {
    using namespace ipp;
    IPP_REQUIRE(buf_host && buflen > 0, IPP_ERR_BAD_ARG, "ipp_timing_timeline: bad buffer");
    std::lock_guard<std::mutex> lock(g_tmu);
    std::string out;
    std::map<cudaStream_t, int> ids;
    for (auto& t : g_timed) {
        float t0 = 0.f, t1 = 0.f;
        if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventSynchronize(g_timed[0].a) == cudaSuccess &&
            cudaEventElapsedTime(&t0, g_timed[0].a, t.a) == cudaSuccess && cudaEventElapsedTime(&t1, g_timed[0].a, t.b) == cudaSuccess) {
            const int id = ids.emplace(t.st, (int)ids.size()).first->second;
            char line[256];
            snprintf(line, sizeof(line), "%s:%d:%.4f:%.4f;", t.name, id, t0, t1);
            if ((int)(out.size() + strlen(line)) < buflen - 1) out += line;
        }
    }
    for (auto& t : g_timed) g_event_pool.push_back({t.a, t.b});
    g_timed.clear();
    snprintf(buf_host, buflen, "%s", out.c_str());
    return IPP_OK;
}

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
