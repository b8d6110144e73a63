// Status strings, device query, pinned-memory helpers of the C ABI.
#include <string.h>
#include <utility>
#include <vector>
#include "sfa_common.cuh"

namespace sfa {
static thread_local char g_last_error[512] = "";
void set_last_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(e),
             cudaGetErrorString(e));
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
static long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }
long long launches() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// ---- dominant-kernel timing (event pairs on the launching stream) ----
static bool g_prof = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static cudaEvent_t g_prof_open = nullptr;
bool prof_enabled() { return g_prof; }
void prof_begin(cudaStream_t st) {
    if (!g_prof) return;
    cudaEvent_t a;
    cudaEventCreate(&a);
    cudaEventRecord(a, st);
    g_prof_open = a;
}
void prof_end(cudaStream_t st) {
    if (!g_prof || !g_prof_open) return;
    cudaEvent_t b;
    cudaEventCreate(&b);
    cudaEventRecord(b, st);
    g_prof_events.emplace_back(g_prof_open, b);
    g_prof_open = nullptr;
}
static void prof_clear() {
    for (auto& e : g_prof_events) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    g_prof_events.clear();
}
}  // namespace sfa

extern "C" {

long long sfa_launch_count(void) { return sfa::launches(); }

int sfa_prof_enable(int on) {
    sfa::prof_clear();
    sfa::g_prof = on != 0;
    return SFA_OK;
}

// Sum of the recorded durations (ms) of the dominant kernel and the number of launches recorded.
int sfa_prof_read(double* total_ms, long long* count) {
    double tot = 0.0;
    for (auto& e : sfa::g_prof_events) {
        SFA_CUDA_CHECK(cudaEventSynchronize(e.second));
        float ms = 0.f;
        SFA_CUDA_CHECK(cudaEventElapsedTime(&ms, e.first, e.second));
        tot += ms;
    }
    if (total_ms) *total_ms = tot;
    if (count) *count = (long long)sfa::g_prof_events.size();
    return SFA_OK;
}

const char* sfa_strerror(int status) {
    switch (status) {
        case SFA_OK: return "ok";
        case SFA_ERR_INVALID: return "invalid argument";
        case SFA_ERR_CUDA: return "CUDA call failed";
        case SFA_ERR_UNSUPPORTED: return "unsupported size or mode";
        case SFA_ERR_WORKSPACE: return "workspace too small";
        case SFA_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        default: return "unknown status";
    }
}

const char* sfa_last_cuda_error(void) { return sfa::g_last_error; }

int sfa_version(void) { return 100; }

int sfa_device_info(int* sms, int* cc_major, int* cc_minor) {
    int dev = 0;
    SFA_CUDA_CHECK(cudaGetDevice(&dev));
    int a = 0, b = 0, c = 0;
    SFA_CUDA_CHECK(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
    SFA_CUDA_CHECK(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
    SFA_CUDA_CHECK(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
    if (sms) *sms = a;
    if (cc_major) *cc_major = b;
    if (cc_minor) *cc_minor = c;
    return SFA_OK;
}

int sfa_host_alloc(void** p, size_t bytes) {
    if (!p) return SFA_ERR_INVALID;
    SFA_CUDA_CHECK(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return SFA_OK;
}

int sfa_host_free(void* p) {
    if (p) SFA_CUDA_CHECK(cudaFreeHost(p));
    return SFA_OK;
}
}
