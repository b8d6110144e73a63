// Status strings, device query, pinned-memory helpers of the C ABI.
#include <stdlib.h>
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
// ---- tuning knobs: environment read once (static initialisation at library load) ----
static double env_num(const char* name, double dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atof(e) : dflt;
}
static int env_flag(const char* name) {
    const char* e = getenv(name);
    return (e && *e && strcmp(e, "0") != 0) ? 1 : 0;
}
static Options make_options() {
    Options o;
    o.pwd_chunk_mb = env_num("SFA_PWD_CHUNK_MB", 256.0);
    o.no_kloop = env_flag("SFA_NO_KLOOP");
    o.no_cluster = env_flag("SFA_NO_CLUSTER");
    o.no_fold = env_flag("SFA_NO_FOLD");
    o.no_tma_store = env_flag("SFA_NO_TMA_STORE");
    o.kloop_block = (int)env_num("SFA_KLOOP_BLOCK", 0.0);
    o.kloop_no_hint = env_flag("SFA_KLOOP_NO_HINT");
    o.issuers = ((int)env_num("SFA_ISSUERS", 1.0) == 2) ? 2 : 1;
    o.old_build = env_flag("SFA_OLD_BUILD");
    o.no_fused = env_flag("SFA_NO_FUSED");
    o.fused_prepass = env_flag("SFA_FUSED_PREPASS");
    o.no_legendre = env_flag("SFA_NO_LEGENDRE");
    o.no_dmma = env_flag("SFA_NO_DMMA");
    o.fft_generic = env_flag("SFA_FFT_GENERIC");
    o.reserve_sms = (int)env_num("SFA_RESERVE_SMS", 0.0);
    o.fused_D = (int)env_num("SFA_FUSED_D", 0.0);
    o.fused_pair = (int)env_num("SFA_FUSED_PAIR", 0.0);
    o.fused_ring = (int)env_num("SFA_FUSED_RING", 0.0);
    o.fused_debug = (int)env_num("SFA_FUSED_DEBUG", 0.0);
    return o;
}
static Options g_options = make_options();
Options& options() { return g_options; }

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
// secondary spans (kind 1: steering build) for the timeline tool
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_aux;
static cudaEvent_t g_prof_aux_open = nullptr;
void prof_aux_begin(cudaStream_t st) {
    if (!g_prof) return;
    cudaEvent_t a;
    cudaEventCreate(&a);
    cudaEventRecord(a, st);
    g_prof_aux_open = a;
}
void prof_aux_end(cudaStream_t st) {
    if (!g_prof || !g_prof_aux_open) return;
    cudaEvent_t b;
    cudaEventCreate(&b);
    cudaEventRecord(b, st);
    g_prof_aux.emplace_back(g_prof_aux_open, b);
    g_prof_aux_open = nullptr;
}
static void prof_clear() {
    for (auto& e : g_prof_aux) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    g_prof_aux.clear();
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

// Elapsed time (ms) from the start of the first recorded launch to the end of the last one:
// span - sum(durations) = idle gaps between the launches of the dominant kernel.
int sfa_prof_span(double* span_ms) {
    float ms = 0.f;
    if (!sfa::g_prof_events.empty()) {
        SFA_CUDA_CHECK(cudaEventSynchronize(sfa::g_prof_events.back().second));
        SFA_CUDA_CHECK(cudaEventElapsedTime(&ms, sfa::g_prof_events.front().first, sfa::g_prof_events.back().second));
    }
    if (span_ms) *span_ms = ms;
    return SFA_OK;
}

// Timeline of the recorded launches relative to the first GEMM start: rows of (kind, start_ms, end_ms),
// kind 0 = tensor-core GEMM, 1 = steering build.  Returns the number of rows written.
int sfa_prof_timeline(double* rows, int max_rows) {
    if (sfa::g_prof_events.empty() || !rows) return 0;
    cudaEvent_t t0 = sfa::g_prof_events.front().first;
    int n = 0;
    for (int kind = 0; kind < 2; ++kind) {
        auto& v = kind == 0 ? sfa::g_prof_events : sfa::g_prof_aux;
        for (auto& e : v) {
            if (n >= max_rows) return n;
            float a = 0.f, b = 0.f;
            if (cudaEventSynchronize(e.second) != cudaSuccess) return n;
            if (cudaEventElapsedTime(&a, t0, e.first) != cudaSuccess) {   // started before t0
                cudaGetLastError();
                float neg = 0.f;
                cudaEventElapsedTime(&neg, e.first, t0);
                a = -neg;
            }
            if (cudaEventElapsedTime(&b, t0, e.second) != cudaSuccess) {
                cudaGetLastError();
                float neg = 0.f;
                cudaEventElapsedTime(&neg, e.second, t0);
                b = -neg;
            }
            rows[3 * n] = kind;
            rows[3 * n + 1] = a;
            rows[3 * n + 2] = b;
            ++n;
        }
    }
    return n;
}

// Tuning knobs (process-global).  Unknown key -> SFA_ERR_INVALID / NaN.
static double* option_slot(const char* key, int** islot) {
    sfa::Options& o = sfa::options();
    *islot = nullptr;
    if (!key) return nullptr;
    if (!strcmp(key, "pwd_chunk_mb")) return &o.pwd_chunk_mb;
    if (!strcmp(key, "no_kloop")) *islot = &o.no_kloop;
    else if (!strcmp(key, "no_cluster")) *islot = &o.no_cluster;
    else if (!strcmp(key, "no_fold")) *islot = &o.no_fold;
    else if (!strcmp(key, "no_tma_store")) *islot = &o.no_tma_store;
    else if (!strcmp(key, "kloop_block")) *islot = &o.kloop_block;
    else if (!strcmp(key, "kloop_no_hint")) *islot = &o.kloop_no_hint;
    else if (!strcmp(key, "issuers")) *islot = &o.issuers;
    else if (!strcmp(key, "old_build")) *islot = &o.old_build;
    else if (!strcmp(key, "no_fused")) *islot = &o.no_fused;
    else if (!strcmp(key, "fused_prepass")) *islot = &o.fused_prepass;
    else if (!strcmp(key, "no_legendre")) *islot = &o.no_legendre;
    else if (!strcmp(key, "no_dmma")) *islot = &o.no_dmma;
    else if (!strcmp(key, "fft_generic")) *islot = &o.fft_generic;
    else if (!strcmp(key, "reserve_sms")) *islot = &o.reserve_sms;
    else if (!strcmp(key, "fused_D")) *islot = &o.fused_D;
    else if (!strcmp(key, "fused_pair")) *islot = &o.fused_pair;
    else if (!strcmp(key, "fused_ring")) *islot = &o.fused_ring;
    else if (!strcmp(key, "fused_debug")) *islot = &o.fused_debug;
    return nullptr;
}
int sfa_set_option(const char* key, double value) {
    int* is = nullptr;
    double* ds = option_slot(key, &is);
    if (ds) *ds = value;
    else if (is) *is = (int)value;
    else return SFA_ERR_INVALID;
    return SFA_OK;
}
double sfa_get_option(const char* key) {
    int* is = nullptr;
    double* ds = option_slot(key, &is);
    if (ds) return *ds;
    if (is) return (double)*is;
    return __builtin_nan("");
}

const char* sfa_strerror(int status) {
    switch (status) {
        case SFA_OK: return "ok";
        case SFA_ERR_INVALID: return "invalid argument";
        case SFA_ERR_CUDA: return "CUDA call failed";
        case SFA_ERR_UNSUPPORTED: return "unsupported size or mode";
        case SFA_ERR_WORKSPACE: return "workspace too small";
        case SFA_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        case SFA_ERR_RANGE: return "radial filters exceed the float32 range of the complex64 paths";
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

// Peer-to-peer DMA copy (copy engines over NVLink / NVSwitch; no SM involved): `bytes` from `src` on device
// `src_dev` to `dst` on device `dst_dev` -- `dst` may be another process's buffer mapped over CUDA IPC --
// asynchronously on `stream` (a stream of src_dev).  Peer access is enabled in both directions on first use.
int sfa_peer_copy(void* dst, int dst_dev, const void* src, int src_dev, size_t bytes, void* stream) {
    if (bytes == 0) return SFA_OK;
    if (!dst || !src || dst_dev < 0 || src_dev < 0 || dst_dev >= 64 || src_dev >= 64) return SFA_ERR_INVALID;
    static unsigned char enabled[64][64];
    if (dst_dev != src_dev && !enabled[src_dev][dst_dev]) {
        int can = 0, cur = -1;
        SFA_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, src_dev, dst_dev));
        if (!can) return SFA_ERR_UNSUPPORTED;
        SFA_CUDA_CHECK(cudaGetDevice(&cur));
        const int pairs[2][2] = {{src_dev, dst_dev}, {dst_dev, src_dev}};
        for (int i = 0; i < 2; ++i) {
            SFA_CUDA_CHECK(cudaSetDevice(pairs[i][0]));
            cudaError_t e = cudaDeviceEnablePeerAccess(pairs[i][1], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            else if (e != cudaSuccess) {
                (void)cudaSetDevice(cur);
                ::sfa::set_last_cuda_error(e, "cudaDeviceEnablePeerAccess");
                return SFA_ERR_CUDA;
            }
        }
        SFA_CUDA_CHECK(cudaSetDevice(cur));
        enabled[src_dev][dst_dev] = 1;
    }
    SFA_CUDA_CHECK(cudaMemcpyPeerAsync(dst, dst_dev, src, src_dev, bytes, (cudaStream_t)stream));
    return SFA_OK;
}
}
