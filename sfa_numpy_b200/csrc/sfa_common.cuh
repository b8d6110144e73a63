// Shared host-side helpers for the C ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sfa_b200.h"

namespace sfa {

void set_last_cuda_error(cudaError_t e, const char* where);

#define SFA_CUDA_CHECK(expr)                                             \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) {                                         \
            ::sfa::set_last_cuda_error(_e, #expr);                       \
            return SFA_ERR_CUDA;                                         \
        }                                                                \
    } while (0)

#define SFA_LAUNCH_CHECK()                                               \
    do {                                                                 \
        ::sfa::count_launch();                                           \
        cudaError_t _e = cudaGetLastError();                             \
        if (_e != cudaSuccess) {                                         \
            ::sfa::set_last_cuda_error(_e, "kernel launch");             \
            return SFA_ERR_CUDA;                                         \
        }                                                                \
    } while (0)

int sm_count();   // cached multiprocessor count of the current device
void count_launch();                 // bumps the library-wide kernel-launch counter
// Optional per-kernel timing of the dominant kernel (bench.py roofline): when enabled, the
// launcher brackets the kernel with CUDA events on the launching stream.
bool prof_enabled();
void prof_begin(cudaStream_t st);
void prof_end(cudaStream_t st);
void prof_aux_begin(cudaStream_t st);   // secondary spans (steering build), timeline only
void prof_aux_end(cudaStream_t st);

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Function attributes are PER DEVICE: the dynamic-shared-memory opt-in is remembered per (call site,
// device), so a process that drives several GPUs opts in on each of them.
static inline int current_device() {
    int dev = -1;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}
#define SFA_ENSURE_DYN_SMEM(kern, bytes)                                                                   \
    do {                                                                                                   \
        static size_t _optin[64] = {0};                                                                    \
        const int _dev = ::sfa::current_device();                                                          \
        const bool _known = _dev >= 0 && _dev < 64;                                                        \
        if ((size_t)(bytes) > 48 * 1024 && (!_known || (size_t)(bytes) > _optin[_dev])) {                  \
            SFA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (_known) _optin[_dev] = (size_t)(bytes);                                                    \
        }                                                                                                  \
    } while (0)

// Library-wide tuning knobs (sfa_set_option / sfa_get_option of the C ABI).  Initialised ONCE, at load
// time, from the SFA_* environment variables of the same meaning; launches never call getenv().
struct Options {
    double pwd_chunk_mb;   // per-slab scratch budget of the chunked stage-3 paths (MiB)
    int no_kloop;          // K > 64 with a shared left operand: split-K launches instead of the K-loop kernel
    int no_cluster;        // no CTA pairs / TMA multicast
    int no_fold;           // resident-K kernel: do not fold bins into the frame tiles
    int no_tma_store;      // resident-K kernel: register stores instead of the TMA-store epilogue
    int kloop_no_hint;     // K-loop kernel: plain output stores (no evict_first policy on large outputs)
    int kloop_block;       // accumulation block of the K-loop kernel in 32-k chunks (0 = default)
    int issuers;           // MMA issuer warps of the resident-K kernel (1 or 2)
    int old_build;         // steering build: generic kernel instead of the few-groups fast kernel
    int no_fused;          // steering form: slab build + GEMM launches instead of the fused persistent kernel
    int no_legendre;       // fused kernel: always build the steering tile from the group-product table, even when the
                           // operators are spherical-harmonic matrices (addition theorem, see pwd_fused_sm100.cu)
    int fused_prepass;     // fused kernel: split ALL spectra in the pre-pass kernel (no in-kernel stream-ahead conversion)
    int no_dmma;           // complex128 path: CUDA-core FMA tile kernel instead of the FP64 tensor-core (DMMA) kernel
    int fft_generic;       // irfft / stft: the generic radix-4 ping-pong kernel instead of the radix-16 fast path
    int reserve_sms;       // SMs the persistent tensor-core kernels leave free (for NCCL's copy kernels when a collective
                           // is meant to run under them)
    int fused_pair;        // fused kernel: the two CTAs of a cluster as ONE tcgen05.mma.cta_group::2 pair
    int fused_D, fused_ring; // stream-ahead distance / plane-ring size of the fused kernel in bins (0 = automatic)
    int fused_debug;       // timing-attribution switches of the fused kernel; honoured by SFA_DEBUG_HOOKS builds only
};
Options& options();
// SMs a persistent kernel may occupy: all of them minus options().reserve_sms (at least 2)
static inline int persistent_sms() {
    int n = sm_count() - options().reserve_sms;
    return n < 2 ? 2 : n;
}

}  // namespace sfa
