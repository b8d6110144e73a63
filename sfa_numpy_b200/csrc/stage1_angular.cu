// Stage 1: spherical- and circular-harmonic matrices (FP64, HBM-store-bound).
//
//   sfa_sht_matrix  <- micarray/modal/angular.py:12-70   (sht_matrix)
//   sfa_cht_matrix  <- micarray/modal/angular.py:106-150 (cht_matrix)
//
// sht kernel: "warp-parallel associated-Legendre recurrence".  A group of G
// lanes (G = 32, or the next power of two >= N+1 for small orders, so that a
// warp carries 32/G points) owns one direction; lane l owns azimuthal order
// m = l (+G, +2G, ... for N >= G).  All lanes step the degree n in lockstep, so
// at step n lanes 0..n hold Pb_n^0..Pb_n^n and the group stores the whole
// degree-n block  Y[q, n^2 .. n^2+2n]  as two contiguous runs of 16-byte
// complex128 values (st.global.v2.f64): every store instruction of the warp is
// a coalesced segment.  Recurrence coefficients a(n,m), b(n,m) and the seeds
// C_m live in shared memory, built once per CTA.  Algorithmic traffic: 16 B per
// SH value out, 24 B per direction in.
#include "sfa_common.cuh"
#include "sfa_math.cuh"

namespace sfa {

__device__ __forceinline__ int tri(int n) { return (n * (n + 1)) >> 1; }

template <int G>
__global__ void __launch_bounds__(256)
sht_matrix_kernel(int N, long long Q, const double* __restrict__ azi,
                  const double* __restrict__ colat, int colat_scalar,
                  const double* __restrict__ wts, double2* __restrict__ Y) {
    extern __shared__ double2 sh_ab[];              // [(N+1)(N+2)/2] (a, b), then C_m as doubles
    const int ntab = tri(N + 1);
    double* sh_c = reinterpret_cast<double*>(sh_ab + ntab);
    for (int t = threadIdx.x; t < ntab; t += blockDim.x) {
        // invert t = n(n+1)/2 + m
        int n = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while (tri(n + 1) <= t) ++n;
        while (tri(n) > t) --n;
        const int m = t - tri(n);
        double2 ab = make_double2(0.0, 0.0);
        if (n > m) {
            ab.x = sh_a(n, m);
            ab.y = (n > m + 1) ? sh_b(n, m) : 0.0;
        }
        sh_ab[t] = ab;
    }
    if (threadIdx.x == 0) {
        double c = 0.28209479177387814347;
        sh_c[0] = c;
        for (int i = 1; i <= N; ++i) {
            c = -c * sqrt((2.0 * i + 1.0) / (2.0 * i));
            sh_c[i] = c;
        }
    }
    __syncthreads();

    constexpr int PPW = 32 / G;                      // directions per warp
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int ml = lane % G;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long M = (long long)(N + 1) * (N + 1);

    for (long long q0 = warp * PPW; q0 < Q; q0 += nwarps * PPW) {
        const long long q = q0 + sub;
        const bool valid = q < Q;
        const long long qs = valid ? q : (Q - 1);
        const double phi = azi[qs];
        const double th = colat[colat_scalar ? 0 : qs];
        const double w = wts ? wts[qs] : 1.0;
        double st, ct;
        sincos(th, &st, &ct);
        double2* __restrict__ row = Y + qs * M;
        for (int m0 = 0; m0 <= N; m0 += G) {
            const int m = m0 + ml;
            const bool lane_on = valid && (m <= N);
            double sm = 0.0, cm = 1.0, p1 = 0.0, p2 = 0.0;
            if (lane_on) {
                sincos((double)m * phi, &sm, &cm);
                p1 = w * sh_c[m] * ipow(st, m);
            }
            const double sg = (m & 1) ? -1.0 : 1.0;
            for (int n = m0; n <= N; ++n) {
                if (lane_on && n >= m) {
                    double p = p1;
                    if (n > m) {
                        const double2 ab = sh_ab[tri(n) + m];
                        p = ab.x * (ct * p1 - ab.y * p2);
                        p2 = p1;
                        p1 = p;
                    }
                    const int base = n * n + n;
                    const double re = p * cm, im = p * sm;
                    row[base + m] = make_double2(re, im);
                    if (m > 0) row[base - m] = make_double2(sg * re, -sg * im);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
cht_matrix_kernel(int N, long long Q, const double* __restrict__ pol,
                  const double* __restrict__ wts, double2* __restrict__ Psi) {
    const int C = 2 * N + 1;
    const long long total = Q * (long long)(N + 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long q = i / (N + 1);
        const int n = (int)(i - q * (N + 1));
        const double w = wts ? wts[q] : 1.0;
        double s, c;
        sincos((double)n * pol[q], &s, &c);
        double2* row = Psi + q * C;
        row[n] = make_double2(w * c, w * s);
        if (n > 0) row[C - n] = make_double2(w * c, w * -s);
    }
}

// Row-block variant for large Q: a warp owns 32 consecutive rows.  Lane = row: one sincos, then
// e^{i n phi} by the angle-addition recurrence (z_n = z_{n-1} z_1; error ~ n ulp, n <= a few hundred),
// written into a shared-memory image of the 32 rows (row pitch C = 2N+1 double2 is odd, so 16-byte
// accesses of 8 consecutive lanes fall into distinct banks).  The 32 rows are contiguous in global memory
// (32 C x 16 B), so the write-out is one stream of fully coalesced 16-byte stores: HBM-store-bound,
// ~0.3 FP64 instructions per value instead of a sincos per value pair.
__global__ void __launch_bounds__(128)
cht_matrix_rows_kernel(int N, long long Q, const double* __restrict__ pol,
                       const double* __restrict__ wts, double2* __restrict__ Psi) {
    extern __shared__ double2 cht_smem[];
    const int C = 2 * N + 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double2* tile = cht_smem + (size_t)wib * 32 * C;
    const long long nblocks = (Q + 31) >> 5;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + wib;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long rb = warp; rb < nblocks; rb += nwarps) {
        const long long q = rb * 32 + lane;
        if (q < Q) {
            const double w = wts ? wts[q] : 1.0;
            double s1, c1;
            sincos(pol[q], &s1, &c1);
            double2* row = tile + (size_t)lane * C;
            row[0] = make_double2(w, 0.0);
            double zr = c1, zi = s1;
            for (int n = 1; n <= N; ++n) {
                row[n] = make_double2(w * zr, w * zi);
                row[C - n] = make_double2(w * zr, -(w * zi));
                const double t = zr * c1 - zi * s1;
                zi = fma(zr, s1, zi * c1);
                zr = t;
            }
        }
        __syncwarp();
        const long long rows = (Q - rb * 32 < 32) ? (Q - rb * 32) : 32;
        const long long count = rows * C;
        double2* dst = Psi + rb * 32 * C;
        for (long long e = lane; e < count; e += 32) __stcs(dst + e, tile[e]);
        __syncwarp();
    }
}

template <int G>
static int launch_sht(int N, long long Q, const double* azi, const double* colat, int cs,
                      const double* w, double2* Y, cudaStream_t st) {
    const size_t smem = (size_t)((N + 1) * (N + 2) / 2) * sizeof(double2) + (size_t)(N + 1) * sizeof(double);
    if (smem > 200 * 1024) return SFA_ERR_UNSUPPORTED;
    SFA_ENSURE_DYN_SMEM(sht_matrix_kernel<G>, smem);
    const int threads = 256;
    const long long warps_needed = ceil_div(Q, 32 / G);
    long long blocks = ceil_div(warps_needed, threads / 32);
    const long long cap = (long long)sm_count() * 8;
    // grid-stride with an even share per CTA (2048 CTAs of work on a cap of 1184 -> 1024 CTAs x 2, not 1184 x 1.7)
    if (blocks > cap) blocks = ceil_div(blocks, ceil_div(blocks, cap));
    if (blocks < 1) blocks = 1;
    sht_matrix_kernel<G><<<(unsigned)blocks, threads, smem, st>>>(N, Q, azi, colat, cs, w, Y);
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

}  // namespace sfa

extern "C" int sfa_sht_matrix(int N, int64_t Q, const double* azi, const double* colat,
                              int colat_is_scalar, const double* weights, sfa_c128* Y, void* stream) {
    if (N < 0 || Q < 0 || (Q > 0 && (!azi || !colat || !Y))) return SFA_ERR_INVALID;
    if (Q == 0) return SFA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    double2* y = reinterpret_cast<double2*>(Y);
    const int need = N + 1;
    if (need <= 1) return sfa::launch_sht<1>(N, Q, azi, colat, colat_is_scalar, weights, y, st);
    if (need <= 2) return sfa::launch_sht<2>(N, Q, azi, colat, colat_is_scalar, weights, y, st);
    if (need <= 4) return sfa::launch_sht<4>(N, Q, azi, colat, colat_is_scalar, weights, y, st);
    if (need <= 8) return sfa::launch_sht<8>(N, Q, azi, colat, colat_is_scalar, weights, y, st);
    if (need <= 16) return sfa::launch_sht<16>(N, Q, azi, colat, colat_is_scalar, weights, y, st);
    return sfa::launch_sht<32>(N, Q, azi, colat, colat_is_scalar, weights, y, st);
}

extern "C" int sfa_cht_matrix(int N, int64_t Q, const double* pol, const double* weights,
                              sfa_c128* Psi, void* stream) {
    if (N < 0 || Q < 0 || (Q > 0 && (!pol || !Psi))) return SFA_ERR_INVALID;
    if (Q == 0) return SFA_OK;
    // many rows: shared-memory row blocks with the recurrence (see cht_matrix_rows_kernel)
    const size_t per_warp = (size_t)32 * (2 * N + 1) * sizeof(double2);
    if (Q >= 4096 && N >= 1 && N <= 256 && per_warp * 4 <= 200 * 1024) {
        const size_t smem = per_warp * 4;
        SFA_ENSURE_DYN_SMEM(sfa::cht_matrix_rows_kernel, smem);
        long long rblocks = sfa::ceil_div(sfa::ceil_div(Q, 32), 4);
        const long long rcap = (long long)sfa::sm_count() * 8;
        if (rblocks > rcap) rblocks = rcap;
        sfa::cht_matrix_rows_kernel<<<(unsigned)rblocks, 128, smem, (cudaStream_t)stream>>>(
            N, Q, pol, weights, reinterpret_cast<double2*>(Psi));
        SFA_LAUNCH_CHECK();
        return SFA_OK;
    }
    const long long total = (long long)Q * (N + 1);
    long long blocks = sfa::ceil_div(total, 256);
    const long long cap = (long long)sfa::sm_count() * 8;
    if (blocks > cap) blocks = cap;
    sfa::cht_matrix_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        N, Q, pol, weights, reinterpret_cast<double2*>(Psi));
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}
