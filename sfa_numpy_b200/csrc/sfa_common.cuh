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

}  // namespace sfa
