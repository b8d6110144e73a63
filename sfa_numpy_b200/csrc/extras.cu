// "Next" rows of SURVEY section 8f: the rest of the reference's angular/util surface.
//   sfa_legendre_matrix   <- angular.Legendre_matrix        micarray/modal/angular.py:73-103
//   sfa_matdiagmul        <- util.matdiagmul                micarray/util.py:65-93
//   sfa_norm_of_columns   <- util.norm_of_columns           micarray/util.py:5-21
//   sfa_coherence_of_columns <- util.coherence_of_columns   micarray/util.py:24-45
//   sfa_grid              <- angular.grid_equal_angle / grid_gauss / grid_equal_polar_angle   angular.py:153-235
#include "sfa_common.cuh"

namespace sfa {

// L[n][q] = (2n+1)/(4 pi) P_n(x_q).  The reference evaluates np.polyval(scipy.special.legendre(n), x)
// (coefficient form, ill-conditioned for large n); here the three-term recurrence
// (n+1) P_{n+1} = (2n+1) x P_n - n P_{n-1}, which is stable for |x| <= 1.
__global__ void legendre_matrix_kernel(int N, long long Q, const double* __restrict__ x, double2* __restrict__ L) {
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < Q;
         q += (long long)gridDim.x * blockDim.x) {
        const double xq = x[q];
        const double inv4pi = 0.07957747154594767;      // 1 / (4 pi)
        double pm = 1.0, p = xq;
        L[q] = make_double2(inv4pi, 0.0);
        if (N >= 1) L[Q + q] = make_double2(3.0 * inv4pi * xq, 0.0);
        for (int n = 1; n < N; ++n) {
            const double pn = ((2 * n + 1) * xq * p - n * pm) / (n + 1);
            pm = p;
            p = pn;
            L[(long long)(n + 1) * Q + q] = make_double2((2 * n + 3) * inv4pi * pn, 0.0);
        }
    }
}

// Sampling grids (angular.py:153-235), one thread per point.
//   kind 0  grid_equal_angle(n):  (2n+2)^2 points, azi = j * 2pi/(2n+2), colat = (i + 1/2) * pi/(2n+2) (np.linspace
//           arithmetic reproduced: i * step, then + step/2), weights 2pi/(n+1) sin(th) sum_{p odd} sin(p th)/p / (n+1)
//   kind 1  grid_gauss(n):        (n+1)(2n+2) points, colat = arccos of the Gauss-Legendre nodes of order n+1 in
//           ascending x (np.polynomial.legendre.leggauss order), weights w_i * pi/(n+1).  Nodes by Newton on the
//           three-term recurrence from the Chebyshev-like guess, weights 2 / ((1 - x^2) P'(x)^2): both at rounding level.
//   kind 2  grid_equal_polar_angle(n): 2n+1 points on the circle, weights 1/(2n+1); colat is not written.
// Layout of the outputs as in the reference: azi = np.tile(azi, rows), colat/weights = np.repeat(., 2n+2).
__device__ double gauss_node(int m, int i, double* weight) {
    const double kPi = 3.14159265358979323846;
    // nodes ascending in x: the i-th smallest root of P_m
    double x = -cos(kPi * (i + 0.75) / (m + 0.5));
    double dp = 1.0;
    for (int it = 0; it < 100; ++it) {
        double p0 = 1.0, p1 = x;
        for (int k = 1; k < m; ++k) {
            const double p2 = ((2 * k + 1) * x * p1 - k * p0) / (k + 1);
            p0 = p1;
            p1 = p2;
        }
        if (m == 0) p1 = 1.0;
        dp = m * (x * p1 - p0) / (x * x - 1.0);
        const double dx = p1 / dp;
        x -= dx;
        if (fabs(dx) <= 1e-16 * fmax(fabs(x), 1e-300) || (it > 3 && fabs(dx) < 4e-16)) break;
    }
    // derivative at the converged node
    {
        double p0 = 1.0, p1 = x;
        for (int k = 1; k < m; ++k) {
            const double p2 = ((2 * k + 1) * x * p1 - k * p0) / (k + 1);
            p0 = p1;
            p1 = p2;
        }
        dp = m * (x * p1 - p0) / (x * x - 1.0);
    }
    *weight = 2.0 / ((1.0 - x * x) * dp * dp);
    return x;
}

__global__ void grid_kernel(int kind, int n, double* __restrict__ azi, double* __restrict__ colat,
                            double* __restrict__ weights) {
    const double kPi = 3.14159265358979323846;
    const int m = 2 * n + 2;
    const long long total = kind == 0 ? (long long)m * m : kind == 1 ? (long long)(n + 1) * m : 2 * n + 1;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        if (kind == 2) {
            const int num = 2 * n + 1;
            const double step = (2 * kPi) / num;
            azi[idx] = idx * step;
            weights[idx] = 1.0 / num * 1.0;
            continue;
        }
        const int i = (int)(idx / m), j = (int)(idx - (long long)i * m);      // row (colatitude), column (azimuth)
        const double astep = (2 * kPi) / m;
        azi[idx] = j * astep;
        if (kind == 0) {
            const double cstep = kPi / m;
            const double th = i * cstep + cstep / 2;
            double s = 0.0;
            for (int p = 1; p < m; p += 2) s += sin(p * th) / p;
            colat[idx] = th;
            weights[idx] = 2 * kPi / (n + 1) * sin(th) * s / (n + 1);
        } else {
            double wgt;
            const double x = gauss_node(n + 1, i, &wgt);
            colat[idx] = acos(x);
            weights[idx] = wgt * (kPi / (n + 1));
        }
    }
}

// C[k][m][n] = A[m][n] * b[k][n]
__global__ void matdiagmul_kernel(long long K, long long M, long long N, const double2* __restrict__ A,
                                  const double2* __restrict__ b, double2* __restrict__ C) {
    const long long total = K * M * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long k = i / (M * N), r = i - k * M * N;
        const long long n = r % N;
        const double2 a = A[r], v = b[k * N + n];
        C[i] = make_double2(a.x * v.x - a.y * v.y, a.x * v.y + a.y * v.x);
    }
}

// one warp per column: (sum_m |A[m][j]|^p)^(1/p); p = 2 uses a scaled sum of squares
__global__ void norm_of_columns_kernel(long long M, long long N, const double2* __restrict__ A, double p,
                                       double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long j = warp; j < N; j += nwarps) {
        double acc = 0.0;
        for (long long m = lane; m < M; m += 32) {
            const double2 v = A[m * N + j];
            const double a = hypot(v.x, v.y);
            acc += (p == 2.0) ? a * a : pow(a, p);
        }
        for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) out[j] = (p == 2.0) ? sqrt(acc) : pow(acc, 1.0 / p);
    }
}

// Beam power map: P[r] = mean_t |q[r][t]|^2 for the R = F*L rows of a beam tensor [F, L, T] with row
// pitch ld (the quantity the reference's examples plot, db(q) over look direction and frequency, and
// the small collective of the multi-GPU path: [F, L] float32 instead of [F, L, T] complex64).
// HBM-read-bound: 8 B in per beam output.  One warp per row, 16-byte loads when the row is aligned.
__global__ void __launch_bounds__(256)
power_map_kernel(long long R, long long T, long long ld, const float2* __restrict__ q, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float inv = 1.0f / (float)T;
    for (long long r = warp; r < R; r += nwarps) {
        const float2* row = q + r * ld;
        float a0 = 0.f, a1 = 0.f;
        if ((((size_t)row) & 15) == 0) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            const long long T2 = T >> 1;
            for (long long t = lane; t < T2; t += 32) {
                const float4 v = __ldcs(row4 + t);           // streamed once
                a0 = fmaf(v.x, v.x, fmaf(v.y, v.y, a0));
                a1 = fmaf(v.z, v.z, fmaf(v.w, v.w, a1));
            }
            if ((T & 1) && lane == 0) {
                const float2 v = row[T - 1];
                a0 = fmaf(v.x, v.x, fmaf(v.y, v.y, a0));
            }
        } else {
            for (long long t = lane; t < T; t += 32) {
                const float2 v = __ldcs(row + t);
                a0 = fmaf(v.x, v.x, fmaf(v.y, v.y, a0));
            }
        }
        float acc = a0 + a1;
        for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) out[r] = acc * inv;
    }
}

// mutual coherence: max_{i != j} |<a_i, a_j>| / (|a_i| |a_j|); one CTA per column i
__global__ void coherence_kernel(long long M, long long N, const double2* __restrict__ A,
                                 const double* __restrict__ norms, unsigned long long* __restrict__ best_bits) {
    __shared__ double red[32];
    const long long i = blockIdx.x;
    double best = 0.0;
    for (long long j = threadIdx.x; j < N; j += blockDim.x) {
        if (j == i) continue;
        double sr = 0.0, si = 0.0;
        for (long long m = 0; m < M; ++m) {
            const double2 a = A[m * N + i], b = A[m * N + j];
            sr += a.x * b.x + a.y * b.y;          // conj(a) * b
            si += a.x * b.y - a.y * b.x;
        }
        const double c = hypot(sr, si) / (norms[i] * norms[j]);
        best = fmax(best, c);
    }
    for (int off = 16; off; off >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        best = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
        for (int off = 16; off; off >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, off));
        // non-negative doubles order like their bit patterns
        if (threadIdx.x == 0) atomicMax(best_bits, (unsigned long long)__double_as_longlong(best));
    }
}

static unsigned grid_for(long long items, int threads) {
    long long b = ceil_div(items, threads);
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace sfa

using namespace sfa;

extern "C" int sfa_legendre_matrix(int N, int64_t Q, const double* ctheta, sfa_c128* L, void* stream) {
    if (N < 0 || Q < 0) return SFA_ERR_INVALID;
    if (Q == 0) return SFA_OK;
    if (!ctheta || !L) return SFA_ERR_INVALID;
    legendre_matrix_kernel<<<grid_for(Q, 256), 256, 0, (cudaStream_t)stream>>>(N, Q, ctheta, reinterpret_cast<double2*>(L));
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

extern "C" int64_t sfa_grid_points(int kind, int n) {
    if (n < 0 || kind < 0 || kind > 2) return -1;
    const int64_t m = 2 * (int64_t)n + 2;
    return kind == 0 ? m * m : kind == 1 ? ((int64_t)n + 1) * m : 2 * (int64_t)n + 1;
}

extern "C" int sfa_grid(int kind, int n, double* azi, double* colat, double* weights, void* stream) {
    const int64_t P = sfa_grid_points(kind, n);
    if (P < 0) return SFA_ERR_INVALID;
    if (!azi || !weights || (kind != 2 && !colat)) return SFA_ERR_INVALID;
    grid_kernel<<<grid_for(P, 128), 128, 0, (cudaStream_t)stream>>>(kind, n, azi, colat, weights);
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

extern "C" int sfa_matdiagmul(int64_t K, int64_t M, int64_t N, const sfa_c128* A, const sfa_c128* b, sfa_c128* C,
                              void* stream) {
    if (K < 0 || M < 0 || N < 0) return SFA_ERR_INVALID;
    if (K * M * N == 0) return SFA_OK;
    if (!A || !b || !C) return SFA_ERR_INVALID;
    matdiagmul_kernel<<<grid_for(K * M * N, 256), 256, 0, (cudaStream_t)stream>>>(
        K, M, N, reinterpret_cast<const double2*>(A), reinterpret_cast<const double2*>(b), reinterpret_cast<double2*>(C));
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

extern "C" int sfa_norm_of_columns(int64_t M, int64_t N, const sfa_c128* A, double p, double* out, void* stream) {
    if (M < 0 || N < 0 || p <= 0.0) return SFA_ERR_INVALID;
    if (N == 0) return SFA_OK;
    if (!A || !out) return SFA_ERR_INVALID;
    norm_of_columns_kernel<<<grid_for(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        M, N, reinterpret_cast<const double2*>(A), p, out);
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

// norms: scratch [N] doubles; result: one double (device)
extern "C" int sfa_coherence_of_columns(int64_t M, int64_t N, const sfa_c128* A, double* norms, double* result,
                                        void* stream) {
    if (M < 0 || N < 0) return SFA_ERR_INVALID;
    if (!result) return SFA_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    SFA_CUDA_CHECK(cudaMemsetAsync(result, 0, sizeof(double), st));
    if (N == 0 || M == 0) return SFA_OK;
    if (!A || !norms) return SFA_ERR_INVALID;
    norm_of_columns_kernel<<<grid_for(N * 32, 256), 256, 0, st>>>(M, N, reinterpret_cast<const double2*>(A), 2.0, norms);
    SFA_LAUNCH_CHECK();
    coherence_kernel<<<(unsigned)N, 128, 0, st>>>(M, N, reinterpret_cast<const double2*>(A), norms,
                                                   reinterpret_cast<unsigned long long*>(result));
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

extern "C" int sfa_power_map(int64_t rows, int64_t T, int64_t ld, const sfa_c64* q, float* out, void* stream) {
    if (rows < 0 || T <= 0 || ld < T) return SFA_ERR_INVALID;
    if (rows == 0) return SFA_OK;
    if (!q || !out) return SFA_ERR_INVALID;
    long long blocks = sfa::ceil_div(rows, 8);
    const long long cap = (long long)sfa::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    sfa::power_map_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rows, T, ld, reinterpret_cast<const float2*>(q), out);
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}
