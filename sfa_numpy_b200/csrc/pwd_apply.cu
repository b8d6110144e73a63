// Stage 3 orchestration: q_f = Y_L diag(d_f) Y_Q^H X_f for every frequency bin.
//
// Reference idiom being replaced (dense diagonal stack + three np.matmul):
//   doc/examples/modal_beamforming_rigid_spherical_array.py:36-39
//   doc/examples/modal_beamforming_open_circular_array.py:35-39
//   radial.diagonal_mode_mat / repeat_n_m              micarray/modal/radial.py:269-314
//
// Two algebraically identical evaluation orders, chosen per shape:
//   form 1 "steering":  A_f = sum_g d_f[g] G_g  with  G_g = sum_{m in g} Y_L[:,m] Y_Q[:,m]^H
//                       (g = SH order n, or the column itself), then q_f = A_f X_f.
//                       Cheapest when Q <= M and T is large (the headline config).
//   form 2 "factored":  Z_f = d_f * (Y_Q^H X_f), q_f = Y_L Z_f.  Cheapest when T is small
//                       or Q > M; never materialises an [F, L, Q] tensor.
// The dense [F, M, M] diagonal stack of the reference is never built.
//
// c64 path: the GEMMs run on tcgen05 (cgemm3x_sm100.cu); c128 path: FP64 FMA kernels below.
#include <stdlib.h>
#include <mutex>
#include "sfa_common.cuh"

namespace sfa {

int launch_cgemm3x(long long B, long long Nn, long long Mm, int K, const float* Pplanar, long long Kp,
                   int p_shared, const float2* R, long long ldR, long long strideR, float* out, long long ldo,
                   long long strideO, int out_mode, const float2* rowscale, long long strideS, int hybrid,
                   cudaStream_t st);

int launch_cgemm3x_kloop(long long B, long long Nn, long long Mm, long long K, const float* Pplanar, long long Kp,
                         const float2* R, long long ldR, long long strideR, float2* out, long long ldo,
                         long long strideO, const float2* rowscale, long long strideS, int hybrid,
                         void* scratch, size_t scratch_bytes, cudaStream_t st);
size_t cgemm3x_kloop_scratch_bytes(long long K, int hybrid);

// fused steering kernel (pwd_fused_sm100.cu): A_f built in tensor memory, one persistent launch
bool pwd_fused_supported(long long L, long long Q, long long T, int Cd, long long out_ld, const void* out);
size_t pwd_fused_g2_bytes(long long L, long long Q, int Cd);
size_t pwd_fused_xs_bytes(long long nf, long long L, long long Q, long long T);
int launch_group_products_g2(const double2* YL, const double2* YQ, long long L, long long Q, int M, int Cd,
                             int per_order, float* G2, cudaStream_t st);
int launch_pwd_fused(long long nf, long long L, long long Q, long long T, int Cd, const float* G2, const double2* d,
                     const float2* X, float* Xs, float2* out, long long ldo, float* power, int hybrid,
                     cudaStream_t st, int pad_writable);

static inline bool use_kloop(long long K, int p_shared) { return K > 64 && p_shared && !options().no_kloop; }

constexpr int kOutC64 = 0, kOutFirstOfMany = 1, kOutAccum = 2;

__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// B-side slot of the hybrid mode: two bf16 in 32 bits, (lo, hi) = (even k, odd k) of the
// interleaved bf16 K axis (the A side holds (hi, lo), see cgemm3x_sm100.cu)
__device__ __forceinline__ float pack_bf16_lo_hi(float hi, float lo) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return __uint_as_float(r);
}
// x (double) -> hi = tf32(x), lo = float(x - hi)
__device__ __forceinline__ void split_from_double(double x, float& hi, float& lo) {
    hi = __uint_as_float(tf32_rna((float)x));
    lo = (float)(x - (double)hi);
}

__device__ __forceinline__ int order_of_column(int i) {     // radial.repeat_n_m: column i -> order n
    int n = (int)sqrtf((float)i);
    while ((n + 1) * (n + 1) <= i) ++n;
    while (n * n > i) --n;
    return n;
}

// S [rows x cols] complex128 (ld) -> planar split P[4][Nn][Kp]; P(n,k) = S(n,k) or conj(S(k,n)).
__global__ void split_planar_kernel(const double2* __restrict__ S, long long ld, int conj_transpose,
                                    long long Nn, long long K, long long Kp, int hybrid, float* __restrict__ P) {
    const long long total = Nn * Kp;
    const long long plane = Nn * Kp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / Kp, k = i - n * Kp;
        double re = 0.0, im = 0.0;
        if (k < K) {
            const double2 v = conj_transpose ? S[k * ld + n] : S[n * ld + k];
            re = v.x;
            im = conj_transpose ? -v.y : v.y;
        }
        float h, l;
        split_from_double(re, h, l);
        P[i] = h;
        P[2 * plane + i] = hybrid ? pack_bf16_lo_hi(h, l) : l;
        split_from_double(im, h, l);
        P[plane + i] = h;
        P[3 * plane + i] = hybrid ? pack_bf16_lo_hi(h, l) : l;
    }
}

// d [F][Cd] complex128 -> per-column diagonal [F][M] (complex64 or complex128)
template <class T2>
__global__ void expand_diag_kernel(const double2* __restrict__ d, long long F, int Cd, int M, int per_order,
                                   T2* __restrict__ out) {
    const long long total = F * M;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long f = i / M;
        const int c = (int)(i - f * M);
        const double2 v = d[f * Cd + (per_order ? order_of_column(c) : c)];
        T2 o;
        o.x = v.x;
        o.y = v.y;
        out[i] = o;
    }
}

// G[g][l*Qp + q] = sum_{m in group g} Y_L[l,m] * conj(Y_Q[q,m])   (FP64 accumulate)
// YQt = conj(Y_Q)^T, [M][Q]: consecutive lanes (consecutive q) read consecutive 16-byte values; with
// Y_Q[q*M + m] every load touched 32 lines and the kernel was L1-wavefront-bound (42 us on cfg 3).
template <class T2>
__global__ void group_products_kernel(const double2* __restrict__ YL, const double2* __restrict__ YQt, long long L,
                                      long long Q, long long Qp, int M, int G, int per_order,
                                      T2* __restrict__ out) {
    const long long total = L * Qp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long l = i / Qp, q = i - l * Qp;
        for (int g = 0; g < G; ++g) {
            double sr = 0.0, si = 0.0;
            if (q < Q) {
                const int m0 = per_order ? g * g : g;
                const int m1 = per_order ? (g + 1) * (g + 1) : g + 1;
                for (int m = m0; m < m1 && m < M; ++m) {
                    const double2 a = YL[l * M + m];
                    const double2 b = YQt[(long long)m * Q + q];
                    sr += a.x * b.x - a.y * b.y;       // a * conj(Y_Q[q,m])
                    si += a.y * b.x + a.x * b.y;
                }
            }
            T2 o;
            o.x = sr;
            o.y = si;
            out[(long long)g * total + i] = o;
        }
    }
}

// Steering matrices of a frequency chunk, written pre-split for the tensor-core GEMM:
//   A[f][plane][l*Qp + q] = split( sum_g d[f][g] * G[g][l*Qp+q] )
// One thread per (l, q) keeps its NG group products in registers and walks the bins of the
// slab; every store instruction of a warp writes 128 contiguous bytes of one plane.
// HBM-store-bound: 16 B out per steering entry.
template <int NG>
__global__ void __launch_bounds__(128)
steer_build_split_kernel(const float2* __restrict__ G, const float2* __restrict__ d, int Cd, int ngroups,
                         long long f0, int nf, long long LQ, int hybrid, float* __restrict__ A) {
    // No shared memory and 128 threads: one of these CTAs fits on an SM beside the (persistent)
    // GEMM CTA.  Each thread owns TWO adjacent steering entries (LQ is a multiple of 4), so every
    // store instruction of a warp writes 256 contiguous bytes of one plane.
    const long long pairs = LQ >> 1;
    for (long long e2 = (long long)blockIdx.x * blockDim.x + threadIdx.x; e2 < pairs;
         e2 += (long long)gridDim.x * blockDim.x) {
        const long long e = e2 * 2;
        float4 g[NG];                                   // (re0, im0, re1, im1) of the two entries
#pragma unroll
        for (int c = 0; c < NG; ++c)
            g[c] = (c < ngroups) ? __ldg(reinterpret_cast<const float4*>(G + (long long)c * LQ + e))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        // blockIdx.y slices the bins of the chunk (more CTAs in flight: the kernel is store-latency bound)
        const int fs = (nf + (int)gridDim.y - 1) / (int)gridDim.y;
        const int fa = (int)blockIdx.y * fs, fb = (fa + fs < nf) ? fa + fs : nf;
        float* o = A + e + (long long)fa * 4 * LQ;
        for (int f = fa; f < fb; ++f) {
            const float2* df = d + (f0 + f) * Cd;           // warp-uniform address: broadcast through L1
            float ar0 = 0.f, ai0 = 0.f, ar1 = 0.f, ai1 = 0.f;
#pragma unroll
            for (int c = 0; c < NG; ++c) {
                if (c < ngroups) {
                    const float2 dd = __ldg(df + c);
                    ar0 = fmaf(dd.x, g[c].x, ar0);
                    ar0 = fmaf(-dd.y, g[c].y, ar0);
                    ai0 = fmaf(dd.x, g[c].y, ai0);
                    ai0 = fmaf(dd.y, g[c].x, ai0);
                    ar1 = fmaf(dd.x, g[c].z, ar1);
                    ar1 = fmaf(-dd.y, g[c].w, ar1);
                    ai1 = fmaf(dd.x, g[c].w, ai1);
                    ai1 = fmaf(dd.y, g[c].z, ai1);
                }
            }
            const float hr0 = __uint_as_float(tf32_rna(ar0)), hi0 = __uint_as_float(tf32_rna(ai0));
            const float hr1 = __uint_as_float(tf32_rna(ar1)), hi1 = __uint_as_float(tf32_rna(ai1));
            __stcs(reinterpret_cast<float2*>(o), make_float2(hr0, hr1));
            __stcs(reinterpret_cast<float2*>(o + LQ), make_float2(hi0, hi1));
            __stcs(reinterpret_cast<float2*>(o + 2 * LQ),
                   hybrid ? make_float2(pack_bf16_lo_hi(hr0, ar0 - hr0), pack_bf16_lo_hi(hr1, ar1 - hr1))
                          : make_float2(ar0 - hr0, ar1 - hr1));
            __stcs(reinterpret_cast<float2*>(o + 3 * LQ),
                   hybrid ? make_float2(pack_bf16_lo_hi(hi0, ai0 - hi0), pack_bf16_lo_hi(hi1, ai1 - hi1))
                          : make_float2(ai0 - hi0, ai1 - hi1));
            o += 4 * LQ;
        }
    }
}

// Faster variant for few groups (spherical arrays up to order 15: Cd = N + 1 <= 16).  The kernel is as
// much FP32-issue-bound as store-bound (36 FMAs per 16 bytes), so:
//  * E adjacent entries per thread (E = 4: every store is 16 bytes per lane, 512 bytes per warp);
//  * the diagonal of the CTA's bins sits in shared memory, pre-duplicated as (re, re, im, im) so that
//    one 16-byte broadcast load feeds two packed FP32x2 FMAs (fma.rn.f32x2) per entry and group:
//    P += re(d) * (g.re, g.im),  Q += im(d) * (g.re, g.im),  A = (P.x - Q.y, P.y + Q.x);
//  * G is read once per 16-64 bins instead of once per 8.
constexpr int kBuildMaxBins = 64;
template <int NG, int E>
__global__ void __launch_bounds__(128)
steer_build_fast_kernel(const float2* __restrict__ G, const float2* __restrict__ d, int Cd, int nf, int bins_per_cta,
                        long long LQ, int hybrid, float* __restrict__ A) {
    __shared__ float4 sd[kBuildMaxBins * NG];
    const int fa = (int)blockIdx.y * bins_per_cta;
    const int fb = (fa + bins_per_cta < nf) ? fa + bins_per_cta : nf;
    for (int i = threadIdx.x; i < (fb - fa) * NG; i += blockDim.x) {
        const int f = i / NG, c = i - f * NG;
        const float2 v = (c < Cd) ? __ldg(d + (long long)(fa + f) * Cd + c) : make_float2(0.f, 0.f);
        sd[i] = make_float4(v.x, v.x, v.y, v.y);
    }
    __syncthreads();
    const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * E;
    if (e >= LQ) return;
    float2 g[NG][E];
#pragma unroll
    for (int c = 0; c < NG; ++c) {
#pragma unroll
        for (int j = 0; j < E; j += 2) {
            const float4 v = (c < Cd) ? __ldg(reinterpret_cast<const float4*>(G + (long long)c * LQ + e + j))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            g[c][j] = make_float2(v.x, v.y);
            g[c][j + 1] = make_float2(v.z, v.w);
        }
    }
    float* o = A + (long long)fa * 4 * LQ + e;
    for (int f = 0; f < fb - fa; ++f) {
        float2 P[E], Qa[E];
#pragma unroll
        for (int j = 0; j < E; ++j) P[j] = Qa[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < NG; ++c) {
            const float4 dd = sd[f * NG + c];                  // same address for the whole warp: broadcast
#pragma unroll
            for (int j = 0; j < E; ++j) {
                P[j] = __ffma2_rn(make_float2(dd.x, dd.y), g[c][j], P[j]);
                Qa[j] = __ffma2_rn(make_float2(dd.z, dd.w), g[c][j], Qa[j]);
            }
        }
        float hr[E], hi[E], xr[E], xi[E];
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const float ar = P[j].x - Qa[j].y, ai = P[j].y + Qa[j].x;
            hr[j] = __uint_as_float(tf32_rna(ar));
            hi[j] = __uint_as_float(tf32_rna(ai));
            xr[j] = hybrid ? pack_bf16_lo_hi(hr[j], ar - hr[j]) : ar - hr[j];
            xi[j] = hybrid ? pack_bf16_lo_hi(hi[j], ai - hi[j]) : ai - hi[j];
        }
        if (E == 4) {
            __stcs(reinterpret_cast<float4*>(o), make_float4(hr[0], hr[1], hr[2], hr[3]));
            __stcs(reinterpret_cast<float4*>(o + LQ), make_float4(hi[0], hi[1], hi[2], hi[3]));
            __stcs(reinterpret_cast<float4*>(o + 2 * LQ), make_float4(xr[0], xr[1], xr[2], xr[3]));
            __stcs(reinterpret_cast<float4*>(o + 3 * LQ), make_float4(xi[0], xi[1], xi[2], xi[3]));
        } else {
            __stcs(reinterpret_cast<float2*>(o), make_float2(hr[0], hr[1]));
            __stcs(reinterpret_cast<float2*>(o + LQ), make_float2(hi[0], hi[1]));
            __stcs(reinterpret_cast<float2*>(o + 2 * LQ), make_float2(xr[0], xr[1]));
            __stcs(reinterpret_cast<float2*>(o + 3 * LQ), make_float2(xi[0], xi[1]));
        }
        o += 4 * LQ;
    }
}

__global__ void d_to_float_kernel(const double2* __restrict__ d, long long n, float2* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = make_float2((float)d[i].x, (float)d[i].y);
}

static int launch_steer_build(const float2* G, const float2* d, int Cd, long long f0, int nf, long long LQ,
                              int hybrid, float* A, cudaStream_t st) {
    const dim3 grid((unsigned)ceil_div(LQ / 2, 128), (unsigned)ceil_div(nf, 8));
    prof_aux_begin(st);
    if (Cd <= 16 && !options().old_build) {
        // ~8 CTAs per SM in flight; G is re-read once per `bpc` bins
        const int E = Cd <= 9 ? 4 : 2;
        const long long gx = ceil_div(LQ / E, 128);
        long long gy = ceil_div((long long)sm_count() * 8, gx);
        int bpc = (int)ceil_div(nf, gy < 1 ? 1 : gy);
        if (bpc < 8) bpc = nf < 8 ? nf : 8;
        if (bpc > kBuildMaxBins) bpc = kBuildMaxBins;
        const dim3 fgrid((unsigned)gx, (unsigned)ceil_div(nf, bpc));
        const float2* dd = d + f0 * Cd;
        if (Cd <= 9)
            steer_build_fast_kernel<9, 4><<<fgrid, 128, 0, st>>>(G, dd, Cd, nf, bpc, LQ, hybrid, A);
        else
            steer_build_fast_kernel<16, 2><<<fgrid, 128, 0, st>>>(G, dd, Cd, nf, bpc, LQ, hybrid, A);
    } else if (Cd <= 16)
        steer_build_split_kernel<16><<<grid, 128, 0, st>>>(G, d, Cd, Cd, f0, nf, LQ, hybrid, A);
    else if (Cd <= 32)
        steer_build_split_kernel<32><<<grid, 128, 0, st>>>(G, d, Cd, Cd, f0, nf, LQ, hybrid, A);
    else
        steer_build_split_kernel<64><<<grid, 128, 0, st>>>(G, d, Cd, Cd, f0, nf, LQ, hybrid, A);
    prof_aux_end(st);
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

// FP64 variant: A[f][l*Q + q] complex128
__global__ void steer_build_f64_kernel(const double2* __restrict__ G, const double2* __restrict__ d, int Cd,
                                       int ngroups, long long f0, int nf, long long LQ,
                                       double2* __restrict__ A) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < LQ * nf;
         e += (long long)gridDim.x * blockDim.x) {
        const long long f = e / LQ, i = e - f * LQ;
        double ar = 0.0, ai = 0.0;
        for (int c = 0; c < ngroups; ++c) {
            const double2 dv = d[(f0 + f) * Cd + c];
            const double2 g = G[(long long)c * LQ + i];
            ar += dv.x * g.x - dv.y * g.y;
            ai += dv.x * g.y + dv.y * g.x;
        }
        A[e] = make_double2(ar, ai);
    }
}

// ---------------------------------------------------------------------------
// FP64 batched complex GEMM on the CUDA cores (c128 path):
//   out[b,n,m] = sum_k P[b|shared][n][k] * R[b][k][m]  (* rowscale[b][n])
// 64 (m) x 64 (n) tile per CTA, 256 threads, 4x4 complex outputs per thread.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
zgemm_kernel(long long B, long long Nn, long long Mm, long long K, const double2* __restrict__ P, long long ldP,
             long long strideP, const double2* __restrict__ R, long long ldR, long long strideR,
             double2* __restrict__ out, long long ldo, long long strideO, const double2* __restrict__ rowscale,
             long long strideS) {
    constexpr int TM = 64, TN = 64, TK = 8;
    __shared__ double2 Rs[TK][TM];
    __shared__ double2 Ps[TN][TK + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long long m_tiles = (Mm + TM - 1) / TM, n_tiles = (Nn + TN - 1) / TN;
    const long long tiles = B * m_tiles * n_tiles;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long b = tile / (m_tiles * n_tiles);
        const long long rem = tile - b * m_tiles * n_tiles;
        const long long nt = rem / m_tiles, mt = rem - nt * m_tiles;
        const long long m0 = mt * TM, n0 = nt * TN;
        const double2* Pb = P + b * strideP;
        const double2* Rb = R + b * strideR;
        double accr[4][4], acci[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) accr[i][j] = acci[i][j] = 0.0;
        for (long long k0 = 0; k0 < K; k0 += TK) {
            for (int e = threadIdx.x; e < TK * TM; e += 256) {
                const int kk = e / TM, mm = e - kk * TM;
                const long long k = k0 + kk, m = m0 + mm;
                Rs[kk][mm] = (k < K && m < Mm) ? Rb[k * ldR + m] : make_double2(0.0, 0.0);
            }
            for (int e = threadIdx.x; e < TN * TK; e += 256) {
                const int nn = e / TK, kk = e - nn * TK;
                const long long k = k0 + kk, n = n0 + nn;
                Ps[nn][kk] = (k < K && n < Nn) ? Pb[n * ldP + k] : make_double2(0.0, 0.0);
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < TK; ++kk) {
                double2 r[4], p[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = Rs[kk][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = Ps[ty * 4 + i][kk];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        accr[i][j] = fma(p[i].x, r[j].x, accr[i][j]);
                        accr[i][j] = fma(-p[i].y, r[j].y, accr[i][j]);
                        acci[i][j] = fma(p[i].x, r[j].y, acci[i][j]);
                        acci[i][j] = fma(p[i].y, r[j].x, acci[i][j]);
                    }
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long n = n0 + ty * 4 + i;
            if (n >= Nn) continue;
            double2 s = make_double2(1.0, 0.0);
            if (rowscale) s = rowscale[b * strideS + n];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long m = m0 + tx * 4 + j;
                if (m >= Mm) continue;
                double xr = accr[i][j], xi = acci[i][j];
                if (rowscale) {
                    const double t = xr * s.x - xi * s.y;
                    xi = xr * s.y + xi * s.x;
                    xr = t;
                }
                out[b * strideO + n * ldo + m] = make_double2(xr, xi);
            }
        }
    }
}

// The same product on the FP64 tensor cores: mma.sync.m8n8k4.f64 (DMMA), four real products per complex
// one.  CTA tile 64 (n) x 64 (m), eight warps as 2 (n) x 4 (m), warp tile 32 x 16 = 4 x 2 DMMA tiles with (Re, Im)
// accumulators; k advances 16 per step.  Operands are staged in shared memory as separate Re / Im planes whose
// pitches (20 and 68 doubles) make every fragment load conflict-free (lane (g, t) of a half-warp hits bank
// 4 g + t); the global loads of k-tile i + 1 are in registers while tile i is multiplied.
// A fragment a0: row g = lane / 4, col t = lane % 4;  B fragment b0: row t, col g;  C fragment: row g, cols 2t, 2t+1.
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
zgemm_dmma_kernel(long long B, long long Nn, long long Mm, long long K, const double2* __restrict__ P, long long ldP,
                  long long strideP, const double2* __restrict__ R, long long ldR, long long strideR,
                  double2* __restrict__ out, long long ldo, long long strideO, const double2* __restrict__ rowscale,
                  long long strideS) {
    constexpr int TM = 64, TN = 64, TK = 16, PP = 20, PR = 68;
    __shared__ double Pre[TN * PP], Pim[TN * PP];          // [n][k]
    __shared__ double Rre[TK * PR], Rim[TK * PR];          // [k][m]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wn = (warp >> 2) * 32, wm = (warp & 3) * 16;  // warp tile origin inside the CTA tile
    const long long m_tiles = (Mm + TM - 1) / TM, n_tiles = (Nn + TN - 1) / TN;
    const long long tiles = B * m_tiles * n_tiles;
    // global -> register staging: thread loads 4 elements of the P tile ([64 n][16 k]) and 4 of the R tile ([16 k][64 m])
    const int pn = threadIdx.x >> 2, pk = (threadIdx.x & 3) * 4;         // P: row pn, k = pk .. pk + 3
    const int rk = threadIdx.x >> 4, rm = (threadIdx.x & 15) * 4;        // R: row rk, m = rm .. rm + 3
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long b = tile / (m_tiles * n_tiles);
        const long long rem = tile - b * m_tiles * n_tiles;
        const long long nt = rem / m_tiles, mt = rem - nt * m_tiles;
        const long long m0 = mt * TM, n0 = nt * TN;
        const double2* Pb = P + b * strideP;
        const double2* Rb = R + b * strideR;
        double acr[4][2][2], aci[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acr[i][j][0] = acr[i][j][1] = aci[i][j][0] = aci[i][j][1] = 0.0;
        double2 pv[4], rv[4];
        auto fetch = [&](long long k0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const long long n = n0 + pn, k = k0 + pk + e;
                pv[e] = (n < Nn && k < K) ? __ldg(Pb + n * ldP + k) : make_double2(0.0, 0.0);
                const long long kk = k0 + rk, m = m0 + rm + e;
                rv[e] = (kk < K && m < Mm) ? __ldg(Rb + kk * ldR + m) : make_double2(0.0, 0.0);
            }
        };
        fetch(0);
        for (long long k0 = 0; k0 < K; k0 += TK) {
            __syncthreads();                                // the previous tile's fragments have been read
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                Pre[pn * PP + pk + e] = pv[e].x;
                Pim[pn * PP + pk + e] = pv[e].y;
                Rre[rk * PR + rm + e] = rv[e].x;
                Rim[rk * PR + rm + e] = rv[e].y;
            }
            __syncthreads();
            if (k0 + TK < K) fetch(k0 + TK);                // in flight during the multiplications below
#pragma unroll
            for (int ks = 0; ks < TK; ks += 4) {
                double are[4], aim[4], nim[4], bre[2], bim[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    are[i] = Pre[(wn + 8 * i + g) * PP + ks + t];
                    aim[i] = Pim[(wn + 8 * i + g) * PP + ks + t];
                    nim[i] = -aim[i];
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    bre[j] = Rre[(ks + t) * PR + wm + 8 * j + g];
                    bim[j] = Rim[(ks + t) * PR + wm + 8 * j + g];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        dmma(acr[i][j], are[i], bre[j]);
                        dmma(acr[i][j], nim[i], bim[j]);
                        dmma(aci[i][j], are[i], bim[j]);
                        dmma(aci[i][j], aim[i], bre[j]);
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long n = n0 + wn + 8 * i + g;
            if (n >= Nn) continue;
            double2 sc = make_double2(1.0, 0.0);
            if (rowscale) sc = rowscale[b * strideS + n];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const long long m = m0 + wm + 8 * j + 2 * t + e;
                    if (m >= Mm) continue;
                    double xr = acr[i][j][e], xi = aci[i][j][e];
                    if (rowscale) {
                        const double tr = xr * sc.x - xi * sc.y;
                        xi = xr * sc.y + xi * sc.x;
                        xr = tr;
                    }
                    out[b * strideO + n * ldo + m] = make_double2(xr, xi);
                }
            }
        }
    }
}

__global__ void conj_transpose_kernel(const double2* __restrict__ S, long long rows, long long cols,
                                      double2* __restrict__ out) {   // out[c][r] = conj(S[r][c])
    const long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long c = i / rows, r = i - c * rows;
        const double2 v = S[r * cols + c];
        out[i] = make_double2(v.x, -v.y);
    }
}

static unsigned grid1d(long long items, int threads = 256) {
    long long b = ceil_div(items, threads);
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline long long round_up(long long x, long long a) { return (x + a - 1) / a * a; }
static long long gcdll(long long a, long long b) { return b ? gcdll(b, a % b) : a; }

struct Plan {
    int form;
    int fused;                  // form 1 through the fused persistent kernel (no steering slabs)
    size_t off_G2, off_Xs;
    long long Qp, Mp, Tld, Old; // padded k extents / leading dims (Z rows, out rows)
    int ngroups;
    long long fc;               // bins per chunk
    size_t off_YQt, off_G, off_A, off_dF, off_P1, off_P2, off_dcols, off_Z, off_scratch, scratch_bytes, total, a_slab;
};

static inline bool is_c64(const sfa_pwd_shape* s) {
    return s->precision == SFA_PREC_C64_3XTF32 || s->precision == SFA_PREC_C64_TF32_BF16;
}
static inline int is_hybrid(const sfa_pwd_shape* s) { return s->precision == SFA_PREC_C64_TF32_BF16 ? 1 : 0; }

static double gemm_cost(double Nn, double Mm, double K, double batches, bool p_shared) {
    // Rough seconds per chunk of `batches` bins: MMA work in padded tile units plus output traffic.
    if (use_kloop((long long)K, p_shared ? 1 : 0)) {
        // in-kernel K loop (cgemm3x_kloop_sm100.cu): folded (batch, m) rows, 128-row n tiles, one output pass
        const double mma = ceil(batches * Mm / 128.0) * 128.0 * ceil(Nn / 128.0) * 128.0 * ceil(K / 32.0) * 32.0 * 8.0 * 3.0;
        return mma / 600e12 + batches * Nn * Mm * 8.0 / 5.0e12;
    }
    // resident-K kernel (cgemm3x_sm100.cu): one launch and one read-modify-write of the output per 64-wide K chunk
    const double mma = ceil(Mm / 128.0) * 128.0 * ceil(Nn / 64.0) * 64.0 * ceil(K / 8.0) * 8.0 * 8.0 * 3.0;
    const double chunks = ceil(K / 64.0);
    const double bytes = Nn * Mm * 8.0 * (1.0 + 2.0 * (chunks - 1.0));
    return batches * (mma / 700e12 + bytes / 6.0e12);
}

static int make_plan(const sfa_pwd_shape* s, Plan* p) {
    if (!s || s->F < 0 || s->L <= 0 || s->Q <= 0 || s->T <= 0 || s->M <= 0 || s->Cd <= 0) return SFA_ERR_INVALID;
    if (!is_c64(s) && s->precision != SFA_PREC_C128_FP64) return SFA_ERR_INVALID;
    if (s->d_is_per_order) {
        const int N1 = s->Cd;
        if (N1 * N1 != s->M) return SFA_ERR_INVALID;
    } else if (s->Cd != s->M) return SFA_ERR_INVALID;
    const bool c64 = is_c64(s);
    p->Qp = round_up(s->Q, 4);
    p->Mp = round_up(s->M, 4);
    // Row pitches that are multiples of 32 frames (256 B) keep every 128-byte line of a tile row inside
    // one CTA's tile; with an unaligned pitch (T = 938 -> 7504 B) the lines at tile edges are shared by
    // two CTAs and HBM absorbs the partial-line writes at 3.3 instead of 4.9 TB/s (tools/store_probe2.cu).
    p->Tld = (s->T > 128) ? round_up(s->T, 32) : s->T;
    p->Old = s->out_ld > 0 ? s->out_ld : s->T;
    if (p->Old < s->T) return SFA_ERR_INVALID;
    p->ngroups = s->Cd;
    int form = s->form;
    if (form == 0) {
        const double nb = 64.0;      // compare per 64 bins (the folded K-loop tiles span several bins)
        double c1 = gemm_cost((double)s->L, (double)s->T, (double)s->Q, nb, false) +
                    nb * (double)s->L * p->Qp * (16.0 * 2.0 / 6.0e12 + 8.0 * s->Cd / 40e12);
        if (c64 && pwd_fused_supported(s->L, s->Q, s->T, s->Cd, p->Old, nullptr)) {
            // fused persistent kernel: no steering slabs; 128-row tiles (pairs: an odd tile count costs a ghost
            // tile, a tail of <= 16 rows is free), 64-frame steps, the build on the CUDA cores beside the MMAs
            long long lt = (s->L % 128 != 0 && s->L % 128 <= 16 && s->L > 128) ? s->L / 128 : ceil_div(s->L, 128);
            if (lt >= 2) lt = round_up(lt, 2);
            const double rows = (double)lt * 128.0;
            const double mma = rows * ceil(s->T / 64.0) * 64.0 * ceil(s->Q / 8.0) * 8.0 * 8.0 * 3.0 / 600e12;
            // the builders stream rows x Qp x Cd group products (8 B each) out of L2 per bin: with many groups and
            // few frames (cfg 2: 33 groups, T = 256) that stream, not the MMAs, sets the pace (measured 0.60 ms
            // against 0.31 ms for the factored form)
            const double build_fma = rows * p->Qp * s->Cd * 8.0 / 30e12;
            // (spherical-harmonic operators with per-order filters are built by the Legendre recurrence: no stream)
            // (circular-harmonic operators get the rotation build, which takes the table stream away as well -- but on
            // BASELINE config 2 (33 columns, 360 look directions = 3 tiles + a ghost, T = 256: four short tasks per
            // bin) the fused kernel still measures 0.54 ms against 0.24 ms for the factored form, so the term stays
            // for them and keeps the plan on the factored form there)
            const bool leg = s->d_is_per_order && s->Cd >= 2 && s->Cd <= 16;
            const double build_l2 = leg ? 0.0 : rows * p->Qp * s->Cd * 8.0 / 2.5e12;
            const double build = build_fma > build_l2 ? build_fma : build_l2;
            const double bytes = (double)s->L * s->T * 8.0 / 4.5e12;
            const double t = (mma > bytes ? mma : bytes);
            c1 = nb * (t > build ? t : build);
        }
        const double c2 = gemm_cost((double)s->M, (double)s->T, (double)s->Q, nb, true) +
                          gemm_cost((double)s->L, (double)s->T, (double)s->M, nb, true);
        form = (c1 <= c2 && s->Cd <= 64) ? 1 : 2;
    }
    if (form == 1 && s->Cd > 64) return SFA_ERR_UNSUPPORTED;
    p->form = form;
    p->fused = (form == 1 && c64 && pwd_fused_supported(s->L, s->Q, s->T, s->Cd, p->Old, nullptr)) ? 1 : 0;
    p->off_G2 = p->off_Xs = 0;
    // bins per chunk: keep the per-chunk scratch around `budget` bytes and the task count a
    // multiple of the SM count (tasks = bins * ceil(T/128)) so persistent CTAs finish together.
    const double budget = options().pwd_chunk_mb * 1024 * 1024;   // per slab
    const double per_bin = (form == 1) ? (double)s->L * p->Qp * 16.0 : (double)s->M * p->Tld * (c64 ? 8.0 : 16.0);
    const long long m_tiles = ceil_div(s->T, 128);
    const long long base = sm_count() / gcdll(sm_count(), m_tiles);
    long long fc = (long long)(budget / per_bin);
    if (fc >= base) fc = fc / base * base;
    if (fc < 1) fc = 1;
    if (fc > s->F) fc = s->F > 0 ? s->F : 1;
    if (p->fused) fc = s->F > 0 ? s->F : 1;     // one launch covers every bin
    p->fc = fc;
    size_t off = 0;
    p->off_G = p->off_A = p->off_dF = p->off_P1 = p->off_P2 = p->off_dcols = p->off_Z = p->a_slab = 0;
    p->off_scratch = p->scratch_bytes = 0;
    const size_t esz = c64 ? 8 : 16;
    p->off_YQt = 0;
    if (p->fused) {
        p->off_YQt = off;
        off += align_up((size_t)s->M * s->Q * 16);
        p->off_G2 = off;                                      // group products in the builders' layout
        off += align_up(pwd_fused_g2_bytes(s->L, s->Q, s->Cd));
        p->off_dF = off;
        off += align_up((size_t)fc * s->Cd * 8);
        p->off_Xs = off;                                      // pre-split spectra [fc][4][T][Kp]
        off += align_up(pwd_fused_xs_bytes(fc, s->L, s->Q, s->T));
    } else if (form == 1) {
        p->off_YQt = off;                                     // conj(Y_Q)^T for the group products
        off += align_up((size_t)s->M * s->Q * 16);
        p->off_G = off;
        off += align_up((size_t)s->Cd * s->L * p->Qp * esz);
        p->off_A = off;                                       // two slabs: build(i+1) overlaps gemm(i)
        p->a_slab = align_up((size_t)fc * s->L * p->Qp * 16); // 4 fp32 planes == one complex128
        off += 2 * p->a_slab;
        p->off_dF = off;                                      // d as complex64 for the build kernel
        off += align_up((size_t)(s->F > fc ? s->F : fc) * s->Cd * 8);
    } else {
        p->off_P1 = off;
        off += align_up((size_t)s->M * p->Qp * 16);
        p->off_P2 = off;
        off += align_up((size_t)s->L * p->Mp * 16);
        p->off_dcols = off;
        off += align_up((size_t)fc * s->M * esz);
        p->off_Z = off;
        off += align_up((size_t)fc * s->M * p->Tld * esz);
        if (c64) {                                           // partial-sum slots of the in-kernel K loop
            size_t a = use_kloop(s->Q, 1) ? cgemm3x_kloop_scratch_bytes(s->Q, is_hybrid(s)) : 0;
            size_t b = use_kloop(s->M, 1) ? cgemm3x_kloop_scratch_bytes(s->M, is_hybrid(s)) : 0;
            p->scratch_bytes = a > b ? a : b;
            p->off_scratch = off;
            off += align_up(p->scratch_bytes);
        }
    }
    p->total = off + 256;
    return SFA_OK;
}

static int cgemm_splitk(long long B, long long Nn, long long Mm, long long K, const float* Pp, long long Kp,
                        int p_shared, const float2* R, long long ldR, long long strideR, float2* out, long long ldo,
                        long long strideO, const float2* rowscale, long long strideS, int hybrid, cudaStream_t st,
                        void* scratch = nullptr, size_t scratch_bytes = 0) {
    // long inner dimension with a shared left operand (factored form): one launch, K loop in the kernel
    if (use_kloop(K, p_shared))
        return launch_cgemm3x_kloop(B, Nn, Mm, K, Pp, Kp, R, ldR, strideR, out, ldo, strideO, rowscale, strideS, hybrid,
                                    scratch, scratch_bytes, st);
    for (long long k0 = 0; k0 < K; k0 += 64) {
        const int kc = (int)((K - k0 < 64) ? (K - k0) : 64);
        // the first pass of a split-K product is marked too, so that every pass uses the same (unfolded) tiling
        const int mode = (k0 == 0) ? (K > 64 ? kOutFirstOfMany : kOutC64) : kOutAccum;
        int rc = launch_cgemm3x(B, Nn, Mm, kc, Pp + k0, Kp, p_shared, R + k0 * ldR, ldR, strideR,
                                reinterpret_cast<float*>(out), ldo, strideO, mode, rowscale, strideS, hybrid, st);
        if (rc != SFA_OK) return rc;
    }
    return SFA_OK;
}

// One-time operator preparation (independent of d and X).
static int pwd_prepare(const sfa_pwd_shape* s, const Plan& p, const double2* yl, const double2* yq,
                       unsigned char* w, cudaStream_t st) {
    const long long L = s->L, Q = s->Q;
    const int M = s->M, Cd = s->Cd;
    const bool c64 = is_c64(s);
    const long long LQ = L * p.Qp;
    if (p.fused)
        return launch_group_products_g2(yl, yq, L, Q, M, Cd, s->d_is_per_order, reinterpret_cast<float*>(w + p.off_G2), st);
    if (p.form == 1) {
        double2* yqt = reinterpret_cast<double2*>(w + p.off_YQt);
        conj_transpose_kernel<<<grid1d(Q * M), 256, 0, st>>>(yq, Q, M, yqt);
        SFA_LAUNCH_CHECK();
        if (c64)
            group_products_kernel<float2><<<grid1d(LQ), 256, 0, st>>>(yl, yqt, L, Q, p.Qp, M, Cd, s->d_is_per_order,
                                                                     reinterpret_cast<float2*>(w + p.off_G));
        else
            group_products_kernel<double2><<<grid1d(LQ), 256, 0, st>>>(yl, yqt, L, Q, p.Qp, M, Cd, s->d_is_per_order,
                                                                      reinterpret_cast<double2*>(w + p.off_G));
        SFA_LAUNCH_CHECK();
    } else if (c64) {
        split_planar_kernel<<<grid1d((long long)M * p.Qp), 256, 0, st>>>(yq, M, 1, M, Q, p.Qp, is_hybrid(s),
                                                                        reinterpret_cast<float*>(w + p.off_P1));
        SFA_LAUNCH_CHECK();
        split_planar_kernel<<<grid1d(L * p.Mp), 256, 0, st>>>(yl, M, 0, L, M, p.Mp, is_hybrid(s),
                                                             reinterpret_cast<float*>(w + p.off_P2));
        SFA_LAUNCH_CHECK();
    } else {
        conj_transpose_kernel<<<grid1d(Q * M), 256, 0, st>>>(yq, Q, M, reinterpret_cast<double2*>(w + p.off_P1));
        SFA_LAUNCH_CHECK();
    }
    return SFA_OK;
}

static unsigned zgrid(long long tiles);
// FP64 complex GEMM launch: DMMA kernel unless switched off ("no_dmma": the CUDA-core FMA tile kernel)
static int launch_zgemm(long long B, long long Nn, long long Mm, long long K, const double2* P, long long ldP,
                        long long strideP, const double2* R, long long ldR, long long strideR, double2* out,
                        long long ldo, long long strideO, const double2* rowscale, long long strideS, cudaStream_t st) {
    const unsigned grid = zgrid(B * ceil_div(Mm, 64) * ceil_div(Nn, 64));
    if (options().no_dmma)
        zgemm_kernel<<<grid, 256, 0, st>>>(B, Nn, Mm, K, P, ldP, strideP, R, ldR, strideR, out, ldo, strideO, rowscale, strideS);
    else
        zgemm_dmma_kernel<<<grid, 256, 0, st>>>(B, Nn, Mm, K, P, ldP, strideP, R, ldR, strideR, out, ldo, strideO, rowscale,
                                               strideS);
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}
static unsigned zgrid(long long tiles) { return (unsigned)(tiles < 65535LL * 16 ? tiles : 65535LL * 16); }

// Bins [0, nf) of d / X / out (pointers already offset to the first bin); nf <= p.fc.
static int pwd_run_bins(const sfa_pwd_shape* s, const Plan& p, long long nf, const double2* yl, const double2* dd,
                        const void* X, void* out, unsigned char* w, cudaStream_t st, float* power = nullptr) {
    const long long L = s->L, Q = s->Q, T = s->T;
    const int M = s->M, Cd = s->Cd;
    const bool c64 = is_c64(s);
    const long long LQ = L * p.Qp;
    int rc;
    if (c64) {
        const float2* x = reinterpret_cast<const float2*>(X);
        float2* o = reinterpret_cast<float2*>(out);
        if (p.fused) {
            if (((uintptr_t)out) & 15) return SFA_ERR_INVALID;              // (a null `out` = power map only)
            return launch_pwd_fused(nf, L, Q, T, Cd, reinterpret_cast<const float*>(w + p.off_G2), dd, x,
                                    reinterpret_cast<float*>(w + p.off_Xs), o, p.Old, power, is_hybrid(s), st,
                                    s->out_pad_writable);
        }
        if (p.form == 1) {
            const float2* G = reinterpret_cast<const float2*>(w + p.off_G);
            float* A = reinterpret_cast<float*>(w + p.off_A);
            float2* dF = reinterpret_cast<float2*>(w + p.off_dF);
            d_to_float_kernel<<<grid1d(nf * Cd), 256, 0, st>>>(dd, nf * Cd, dF);
            SFA_LAUNCH_CHECK();
            rc = launch_steer_build(G, dF, Cd, 0, (int)nf, LQ, is_hybrid(s), A, st);
            if (rc != SFA_OK) return rc;
            return cgemm_splitk(nf, L, T, Q, A, p.Qp, 0, x, T, Q * T, o, p.Old, L * p.Old, nullptr, 0, is_hybrid(s), st);
        }
        const float* P1 = reinterpret_cast<const float*>(w + p.off_P1);
        const float* P2 = reinterpret_cast<const float*>(w + p.off_P2);
        float2* dcols = reinterpret_cast<float2*>(w + p.off_dcols);
        float2* Z = reinterpret_cast<float2*>(w + p.off_Z);
        expand_diag_kernel<float2><<<grid1d(nf * M), 256, 0, st>>>(dd, nf, Cd, M, s->d_is_per_order, dcols);
        SFA_LAUNCH_CHECK();
        rc = cgemm_splitk(nf, M, T, Q, P1, p.Qp, 1, x, T, Q * T, Z, p.Tld, (long long)M * p.Tld, dcols, M, is_hybrid(s), st,
                          w + p.off_scratch, p.scratch_bytes);
        if (rc != SFA_OK) return rc;
        return cgemm_splitk(nf, L, T, M, P2, p.Mp, 1, Z, p.Tld, (long long)M * p.Tld, o, p.Old, L * p.Old, nullptr, 0, is_hybrid(s), st,
                            w + p.off_scratch, p.scratch_bytes);
    }
    const double2* x = reinterpret_cast<const double2*>(X);
    double2* o = reinterpret_cast<double2*>(out);
    if (p.form == 1) {
        const double2* G = reinterpret_cast<const double2*>(w + p.off_G);
        double2* A = reinterpret_cast<double2*>(w + p.off_A);
        steer_build_f64_kernel<<<grid1d(LQ * nf), 256, 0, st>>>(G, dd, Cd, Cd, 0, (int)nf, LQ, A);
        SFA_LAUNCH_CHECK();
        return launch_zgemm(nf, L, T, Q, A, p.Qp, LQ, x, T, Q * T, o, p.Old, L * p.Old, nullptr, 0, st);
    }
    const double2* P1 = reinterpret_cast<const double2*>(w + p.off_P1);      // conj(Y_Q)^T [M][Q]
    double2* dcols = reinterpret_cast<double2*>(w + p.off_dcols);
    double2* Z = reinterpret_cast<double2*>(w + p.off_Z);
    expand_diag_kernel<double2><<<grid1d(nf * M), 256, 0, st>>>(dd, nf, Cd, M, s->d_is_per_order, dcols);
    SFA_LAUNCH_CHECK();
    rc = launch_zgemm(nf, M, T, Q, P1, Q, 0, x, T, Q * T, Z, p.Tld, (long long)M * p.Tld, dcols, M, st);
    if (rc != SFA_OK) return rc;
    return launch_zgemm(nf, L, T, M, yl, M, 0, Z, p.Tld, (long long)M * p.Tld, o, p.Old, L * p.Old, nullptr, 0, st);
}

// Form 1, several chunks: the (HBM-store-bound) steering-matrix build of chunk i+1 is issued on a
// side stream as soon as its slab is free (GEMM i-1 done), the GEMM of chunk i on the caller's stream.
// Measured with the launch timeline (sfa_prof_timeline, tools/prof_pwd.py): the block scheduler never
// places a build CTA on an SM that hosts a persistent GEMM CTA, whatever the build CTA's size (tried
// 32-128 threads x 62 registers, no shared memory, matching carve-out, stream priorities), so the
// build runs in the ~48 us window between two GEMM launches, overlapping only their ramp-down/ramp-up.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, built[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
};
static SideStream* side_for_current_device() {
    static SideStream side[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SideStream& s = side[dev];
    if (!s.stream) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
        for (int i = 0; i < 2; ++i) {
            cudaEventCreateWithFlags(&s.built[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&s.done[i], cudaEventDisableTiming);
        }
    }
    return &s;
}

static int pwd_steering_pipelined(const sfa_pwd_shape* s, const Plan& p, const double2* dd, const float2* x,
                                  float2* o, unsigned char* w, cudaStream_t st) {
    // The side stream and its events are per device: calls from several host threads enqueue one after
    // the other (the event waits are bound at enqueue time, so the GPU-side dependencies stay per call).
    static std::mutex enqueue_mutex;
    std::lock_guard<std::mutex> lock(enqueue_mutex);
    SideStream* ss = side_for_current_device();
    if (!ss) return SFA_ERR_CUDA;
    const long long L = s->L, Q = s->Q, T = s->T, F = s->F;
    const int Cd = s->Cd;
    const long long LQ = L * p.Qp;
    const float2* G = reinterpret_cast<const float2*>(w + p.off_G);
    float* Aslab[2] = {reinterpret_cast<float*>(w + p.off_A), reinterpret_cast<float*>(w + p.off_A + p.a_slab)};
    float2* dF = reinterpret_cast<float2*>(w + p.off_dF);
    d_to_float_kernel<<<grid1d(F * Cd), 256, 0, st>>>(dd, F * Cd, dF);
    SFA_LAUNCH_CHECK();
    SFA_CUDA_CHECK(cudaEventRecord(ss->fork, st));               // G, dF are ready / earlier work ordered
    SFA_CUDA_CHECK(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
    int slot = 0;
    long long it = 0;
    for (long long f0 = 0; f0 < F; f0 += p.fc, ++it, slot ^= 1) {
        const long long nf = (F - f0 < p.fc) ? (F - f0) : p.fc;
        if (it >= 2) SFA_CUDA_CHECK(cudaStreamWaitEvent(ss->stream, ss->done[slot], 0));
        {
            int rc = launch_steer_build(G, dF + f0 * Cd, Cd, 0, (int)nf, LQ, is_hybrid(s), Aslab[slot], ss->stream);
            if (rc != SFA_OK) return rc;
        }
        SFA_CUDA_CHECK(cudaEventRecord(ss->built[slot], ss->stream));
        SFA_CUDA_CHECK(cudaStreamWaitEvent(st, ss->built[slot], 0));
        int rc = cgemm_splitk(nf, L, T, Q, Aslab[slot], p.Qp, 0, x + f0 * Q * T, T, Q * T, o + f0 * L * p.Old, p.Old, L * p.Old,
                              nullptr, 0, is_hybrid(s), st);
        if (rc != SFA_OK) return rc;
        SFA_CUDA_CHECK(cudaEventRecord(ss->done[slot], st));
    }
    return SFA_OK;
}

static inline size_t elem_bytes(const sfa_pwd_shape* s) { return is_c64(s) ? 8 : 16; }

}  // namespace sfa

using namespace sfa;

extern "C" size_t sfa_pwd_work_bytes(const sfa_pwd_shape* s) {
    Plan p;
    if (make_plan(s, &p) != SFA_OK) return 0;
    return p.total;
}

// Operator preparation (everything that depends on Y_L / Y_Q only): group products or split operand planes,
// written into `work`.  Valid until Y_L, Y_Q or the shape change; sfa_pwd_run reads it.
extern "C" int sfa_pwd_prepare(const sfa_pwd_shape* s, const sfa_c128* Y_L, const sfa_c128* Y_Q, void* work,
                               size_t work_bytes, void* stream) {
    Plan p;
    int rc = make_plan(s, &p);
    if (rc != SFA_OK) return rc;
    if (!Y_L || !Y_Q || !work) return SFA_ERR_INVALID;
    if (work_bytes < p.total) return SFA_ERR_WORKSPACE;
    unsigned char* w = reinterpret_cast<unsigned char*>(((uintptr_t)work + 255) & ~(uintptr_t)255);
    return pwd_prepare(s, p, reinterpret_cast<const double2*>(Y_L), reinterpret_cast<const double2*>(Y_Q), w,
                       (cudaStream_t)stream);
}

// 1 when the plan for this shape is the fused persistent steering kernel (power map in the epilogue, map-only runs)
extern "C" int sfa_pwd_is_fused(const sfa_pwd_shape* s) {
    Plan p;
    if (make_plan(s, &p) != SFA_OK) return 0;
    return p.fused;
}

extern "C" int sfa_pwd_run(const sfa_pwd_shape* s, const sfa_c128* Y_L, const sfa_c128* d, const void* X, void* out,
                           float* power, void* work, size_t work_bytes, void* stream) {
    Plan p;
    int rc = make_plan(s, &p);
    if (rc != SFA_OK) return rc;
    if (s->F == 0) return SFA_OK;
    if (!Y_L || !d || !X || !work) return SFA_ERR_INVALID;
    if (power && !is_c64(s)) return SFA_ERR_UNSUPPORTED;
    // out == NULL: only the beam power map is wanted.  The fused steering kernel then skips its stores (the beams
    // exist in tensor memory only); no other path can do that.
    if (!out && !(power && p.fused)) return SFA_ERR_UNSUPPORTED;
    if (work_bytes < p.total) return SFA_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* w = reinterpret_cast<unsigned char*>(((uintptr_t)work + 255) & ~(uintptr_t)255);
    const double2* yl = reinterpret_cast<const double2*>(Y_L);
    const double2* dd = reinterpret_cast<const double2*>(d);
    const size_t eb = elem_bytes(s);
    if (is_c64(s) && p.form == 1 && !p.fused && s->F > p.fc) {
        rc = pwd_steering_pipelined(s, p, dd, reinterpret_cast<const float2*>(X), reinterpret_cast<float2*>(out), w, st);
    } else {
        for (long long f0 = 0; f0 < s->F && rc == SFA_OK; f0 += p.fc) {
            const long long nf = (s->F - f0 < p.fc) ? (s->F - f0) : p.fc;
            rc = pwd_run_bins(s, p, nf, yl, dd + f0 * s->Cd,
                              reinterpret_cast<const unsigned char*>(X) + (size_t)f0 * s->Q * s->T * eb,
                              out ? reinterpret_cast<unsigned char*>(out) + (size_t)f0 * s->L * p.Old * eb : nullptr, w, st,
                              (power && p.fused) ? power + f0 * s->L : nullptr);
        }
    }
    if (rc != SFA_OK) return rc;
    // every path but the fused kernel (which accumulates the map in its epilogue) takes one more pass over the beams
    if (power && !p.fused)
        return sfa_power_map(s->F * s->L, s->T, p.Old, reinterpret_cast<const sfa_c64*>(out), power, stream);
    return SFA_OK;
}

extern "C" int sfa_pwd_apply(const sfa_pwd_shape* s, const sfa_c128* Y_L, const sfa_c128* Y_Q,
                             const sfa_c128* d, const void* X, void* out, void* work, size_t work_bytes,
                             void* stream) {
    if (s && s->F == 0) {
        Plan p;
        return make_plan(s, &p);
    }
    int rc = sfa_pwd_prepare(s, Y_L, Y_Q, work, work_bytes, stream);
    if (rc != SFA_OK) return rc;
    return sfa_pwd_run(s, Y_L, d, X, out, nullptr, work, work_bytes, stream);
}

// ---------------------------------------------------------------------------
// Host-buffer path: chunked H2D -> compute -> D2H pipeline on three streams.
// ---------------------------------------------------------------------------
struct sfa_pwd_host_ctx {
    sfa_pwd_shape shape;
    Plan plan;
    int device;
    long long chunk;                 // bins per pipeline slot (multiple of plan.fc)
    cudaStream_t s_in, s_cmp, s_out;
    cudaEvent_t in_done[2], cmp_done[2], out_done[2];
    unsigned char *d_YL, *d_YQ, *d_d, *d_work, *d_X[2], *d_O[2];
    float* d_P;                      // power map [F][L], allocated on first use
    bool operators_set;
};

extern "C" int sfa_pwd_host_create(sfa_pwd_host_ctx** out_ctx, const sfa_pwd_shape* s, int device,
                                   int64_t chunk_bins) {
    if (!out_ctx || !s) return SFA_ERR_INVALID;
    SFA_CUDA_CHECK(cudaSetDevice(device));
    sfa_pwd_host_ctx* c = new sfa_pwd_host_ctx();
    c->shape = *s;
    c->shape.out_ld = (s->T > 128) ? round_up(s->T, 32) : s->T;     // device-side row pitch of the beams
    c->shape.out_pad_writable = 1;                                  // the device slabs are the library's own
    s = &c->shape;
    int rc = make_plan(s, &c->plan);
    if (rc != SFA_OK) {
        delete c;
        return rc;
    }
    c->device = device;
    // fused steering plans cover every bin with one launch: the host pipeline still moves 74-bin chunks so
    // that H2D, compute and D2H overlap
    long long chunk = chunk_bins > 0 ? chunk_bins : (c->plan.fused ? 74 : c->plan.fc);
    if (chunk > s->F) chunk = s->F > 0 ? s->F : 1;
    c->chunk = chunk;
    c->operators_set = false;
    const size_t eb = elem_bytes(s);
    // every handle starts out null (value-initialised ctx), so a failure half-way releases exactly
    // what exists: sfa_pwd_host_destroy skips null handles
    rc = [&]() -> int {
        SFA_CUDA_CHECK(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
        SFA_CUDA_CHECK(cudaStreamCreateWithFlags(&c->s_cmp, cudaStreamNonBlocking));
        SFA_CUDA_CHECK(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            SFA_CUDA_CHECK(cudaEventCreateWithFlags(&c->in_done[i], cudaEventDisableTiming));
            SFA_CUDA_CHECK(cudaEventCreateWithFlags(&c->cmp_done[i], cudaEventDisableTiming));
            SFA_CUDA_CHECK(cudaEventCreateWithFlags(&c->out_done[i], cudaEventDisableTiming));
            SFA_CUDA_CHECK(cudaMalloc(&c->d_X[i], (size_t)chunk * s->Q * s->T * eb));
            SFA_CUDA_CHECK(cudaMalloc(&c->d_O[i], (size_t)chunk * s->L * s->out_ld * eb));
        }
        SFA_CUDA_CHECK(cudaMalloc(&c->d_YL, (size_t)s->L * s->M * 16));
        SFA_CUDA_CHECK(cudaMalloc(&c->d_YQ, (size_t)s->Q * s->M * 16));
        SFA_CUDA_CHECK(cudaMalloc(&c->d_d, (size_t)(s->F > 0 ? s->F : 1) * s->Cd * 16));
        SFA_CUDA_CHECK(cudaMalloc(&c->d_work, c->plan.total));
        return SFA_OK;
    }();
    if (rc != SFA_OK) {
        sfa_pwd_host_destroy(c);
        cudaGetLastError();             // the error is reported through rc / sfa_last_cuda_error
        return rc;
    }
    *out_ctx = c;
    return SFA_OK;
}

extern "C" int sfa_pwd_host_set_operators(sfa_pwd_host_ctx* c, const sfa_c128* Y_L, const sfa_c128* Y_Q,
                                          const sfa_c128* d) {
    if (!c || !Y_L || !Y_Q || !d) return SFA_ERR_INVALID;
    const sfa_pwd_shape* s = &c->shape;
    if (is_c64(s)) {
        // the complex64 paths narrow the filters to float: |1/b_n| of an unregularised open array near k = 0 would
        // turn into Inf and the beams into NaN without a word (the reference computes that case in complex128)
        const double* dv = reinterpret_cast<const double*>(d);
        for (long long i = 0; i < 2 * s->F * s->Cd; ++i)
            if (!(fabs(dv[i]) <= 1e30)) return SFA_ERR_RANGE;
    }
    SFA_CUDA_CHECK(cudaSetDevice(c->device));
    SFA_CUDA_CHECK(cudaMemcpyAsync(c->d_YL, Y_L, (size_t)s->L * s->M * 16, cudaMemcpyHostToDevice, c->s_cmp));
    SFA_CUDA_CHECK(cudaMemcpyAsync(c->d_YQ, Y_Q, (size_t)s->Q * s->M * 16, cudaMemcpyHostToDevice, c->s_cmp));
    SFA_CUDA_CHECK(cudaMemcpyAsync(c->d_d, d, (size_t)s->F * s->Cd * 16, cudaMemcpyHostToDevice, c->s_cmp));
    unsigned char* w = reinterpret_cast<unsigned char*>(((uintptr_t)c->d_work + 255) & ~(uintptr_t)255);
    int rc = pwd_prepare(s, c->plan, reinterpret_cast<const double2*>(c->d_YL),
                         reinterpret_cast<const double2*>(c->d_YQ), w, c->s_cmp);
    if (rc != SFA_OK) return rc;
    SFA_CUDA_CHECK(cudaStreamSynchronize(c->s_cmp));
    c->operators_set = true;
    return SFA_OK;
}

extern "C" int sfa_pwd_host_run(sfa_pwd_host_ctx* c, const void* X_host, void* out_host) {
    return sfa_pwd_host_run_power(c, X_host, out_host, nullptr);
}

extern "C" int sfa_pwd_host_set_operators_dev(sfa_pwd_host_ctx* c, const sfa_c128* Y_L, const sfa_c128* Y_Q,
                                              const sfa_c128* d, void* stream) {
    if (!c || !Y_L || !Y_Q || !d) return SFA_ERR_INVALID;
    const sfa_pwd_shape* s = &c->shape;
    SFA_CUDA_CHECK(cudaSetDevice(c->device));
    cudaStream_t src = (cudaStream_t)stream;
    // order after the producer of the operators (stages 1-2 on the caller's stream), then copy device to device
    cudaEvent_t ev;
    SFA_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    SFA_CUDA_CHECK(cudaEventRecord(ev, src));
    SFA_CUDA_CHECK(cudaStreamWaitEvent(c->s_cmp, ev, 0));
    SFA_CUDA_CHECK(cudaEventDestroy(ev));
    SFA_CUDA_CHECK(cudaMemcpyAsync(c->d_YL, Y_L, (size_t)s->L * s->M * 16, cudaMemcpyDeviceToDevice, c->s_cmp));
    SFA_CUDA_CHECK(cudaMemcpyAsync(c->d_YQ, Y_Q, (size_t)s->Q * s->M * 16, cudaMemcpyDeviceToDevice, c->s_cmp));
    SFA_CUDA_CHECK(cudaMemcpyAsync(c->d_d, d, (size_t)s->F * s->Cd * 16, cudaMemcpyDeviceToDevice, c->s_cmp));
    unsigned char* w = reinterpret_cast<unsigned char*>(((uintptr_t)c->d_work + 255) & ~(uintptr_t)255);
    int rc = pwd_prepare(s, c->plan, reinterpret_cast<const double2*>(c->d_YL),
                         reinterpret_cast<const double2*>(c->d_YQ), w, c->s_cmp);
    if (rc != SFA_OK) return rc;
    SFA_CUDA_CHECK(cudaStreamSynchronize(c->s_cmp));      // the caller may overwrite its operator buffers now
    c->operators_set = true;
    return SFA_OK;
}

extern "C" int sfa_pwd_host_run_power(sfa_pwd_host_ctx* c, const void* X_host, void* out_host, float* power_host) {
    if (!c || !X_host || (!out_host && !power_host) || !c->operators_set) return SFA_ERR_INVALID;
    const sfa_pwd_shape* s = &c->shape;
    if (power_host && !is_c64(s)) return SFA_ERR_UNSUPPORTED;
    SFA_CUDA_CHECK(cudaSetDevice(c->device));
    if (power_host && !c->d_P) SFA_CUDA_CHECK(cudaMalloc(&c->d_P, (size_t)(s->F > 0 ? s->F : 1) * s->L * sizeof(float)));
    const size_t eb = elem_bytes(s);
    unsigned char* w = reinterpret_cast<unsigned char*>(((uintptr_t)c->d_work + 255) & ~(uintptr_t)255);
    const size_t x_bin = (size_t)s->Q * s->T * eb, o_bin = (size_t)s->L * s->T * eb;
    const size_t o_bin_dev = (size_t)s->L * s->out_ld * eb;        // padded rows on the device
    int slot = 0;
    long long it = 0;
    for (long long f0 = 0; f0 < s->F; f0 += c->chunk, ++it, slot ^= 1) {
        const long long nf = (s->F - f0 < c->chunk) ? (s->F - f0) : c->chunk;
        // the input slot is free once the compute that read it (two chunks ago) has finished
        if (it >= 2) SFA_CUDA_CHECK(cudaStreamWaitEvent(c->s_in, c->cmp_done[slot], 0));
        SFA_CUDA_CHECK(cudaMemcpyAsync(c->d_X[slot], reinterpret_cast<const unsigned char*>(X_host) + f0 * x_bin,
                                       nf * x_bin, cudaMemcpyHostToDevice, c->s_in));
        SFA_CUDA_CHECK(cudaEventRecord(c->in_done[slot], c->s_in));
        SFA_CUDA_CHECK(cudaStreamWaitEvent(c->s_cmp, c->in_done[slot], 0));
        // the output slot is free once its previous contents have been copied out
        if (it >= 2) SFA_CUDA_CHECK(cudaStreamWaitEvent(c->s_cmp, c->out_done[slot], 0));
        for (long long g0 = 0; g0 < nf; g0 += c->plan.fc) {
            const long long ng = (nf - g0 < c->plan.fc) ? (nf - g0) : c->plan.fc;
            float* pw = power_host ? c->d_P + (f0 + g0) * s->L : nullptr;
            // only the map is wanted and the plan is the fused kernel: the beams never leave tensor memory
            const bool map_only = !out_host && pw && c->plan.fused;
            int rc = pwd_run_bins(s, c->plan, ng, reinterpret_cast<const double2*>(c->d_YL),
                                  reinterpret_cast<const double2*>(c->d_d) + (f0 + g0) * s->Cd,
                                  c->d_X[slot] + g0 * x_bin, map_only ? nullptr : c->d_O[slot] + g0 * o_bin_dev, w, c->s_cmp,
                                  c->plan.fused ? pw : nullptr);
            if (rc != SFA_OK) return rc;
            if (pw && !c->plan.fused) {
                rc = sfa_power_map(ng * s->L, s->T, s->out_ld, reinterpret_cast<const sfa_c64*>(c->d_O[slot] + g0 * o_bin_dev),
                                   pw, c->s_cmp);
                if (rc != SFA_OK) return rc;
            }
        }
        SFA_CUDA_CHECK(cudaEventRecord(c->cmp_done[slot], c->s_cmp));
        SFA_CUDA_CHECK(cudaStreamWaitEvent(c->s_out, c->cmp_done[slot], 0));
        // rows are compacted by the copy engine (device pitch out_ld -> host pitch T)
        if (out_host)
            SFA_CUDA_CHECK(cudaMemcpy2DAsync(reinterpret_cast<unsigned char*>(out_host) + f0 * o_bin, (size_t)s->T * eb,
                                             c->d_O[slot], (size_t)s->out_ld * eb, (size_t)s->T * eb,
                                             (size_t)nf * s->L, cudaMemcpyDeviceToHost, c->s_out));
        SFA_CUDA_CHECK(cudaEventRecord(c->out_done[slot], c->s_out));
    }
    if (power_host)
        SFA_CUDA_CHECK(cudaMemcpyAsync(power_host, c->d_P, (size_t)s->F * s->L * sizeof(float), cudaMemcpyDeviceToHost,
                                       c->s_cmp));
    SFA_CUDA_CHECK(cudaStreamSynchronize(c->s_out));
    SFA_CUDA_CHECK(cudaStreamSynchronize(c->s_cmp));
    return SFA_OK;
}

extern "C" int sfa_pwd_host_destroy(sfa_pwd_host_ctx* c) {
    if (!c) return SFA_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; ++i) {
        if (c->d_X[i]) cudaFree(c->d_X[i]);
        if (c->d_O[i]) cudaFree(c->d_O[i]);
        if (c->in_done[i]) cudaEventDestroy(c->in_done[i]);
        if (c->cmp_done[i]) cudaEventDestroy(c->cmp_done[i]);
        if (c->out_done[i]) cudaEventDestroy(c->out_done[i]);
    }
    if (c->d_YL) cudaFree(c->d_YL);
    if (c->d_YQ) cudaFree(c->d_YQ);
    if (c->d_d) cudaFree(c->d_d);
    if (c->d_work) cudaFree(c->d_work);
    if (c->d_P) cudaFree(c->d_P);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_cmp) cudaStreamDestroy(c->s_cmp);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    delete c;
    return SFA_OK;
}

// Test/bench entry: raw complex64 operands; P is split into planes in `scratch`-free fashion by
// allocating a temporary (stream-ordered) buffer.
extern "C" int sfa_cgemm3x(int64_t B, int64_t Nn, int64_t Mm, int64_t K, const sfa_c64* P, int64_t ldP,
                           int64_t strideP, const sfa_c64* R, int64_t ldR, int64_t strideR, sfa_c64* out,
                           int64_t ldo, int64_t strideO, const sfa_c64* rowscale, int64_t strideS,
                           int hybrid, void* stream);

namespace sfa {
__global__ void split_planar_c64_kernel(const float2* __restrict__ S, long long ld, long long strideB,
                                        long long batches, long long Nn, long long K, long long Kp, int hybrid,
                                        float* __restrict__ P) {
    const long long plane = Nn * Kp;
    const long long total = batches * plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / plane, r = i - b * plane;
        const long long n = r / Kp, k = r - n * Kp;
        float2 v = make_float2(0.f, 0.f);
        if (k < K) v = S[b * strideB + n * ld + k];
        float* o = P + b * 4 * plane + r;
        const float hr = __uint_as_float(tf32_rna(v.x)), hi = __uint_as_float(tf32_rna(v.y));
        o[0] = hr;
        o[plane] = hi;
        o[2 * plane] = hybrid ? pack_bf16_lo_hi(hr, v.x - hr) : v.x - hr;
        o[3 * plane] = hybrid ? pack_bf16_lo_hi(hi, v.y - hi) : v.y - hi;
    }
}
}  // namespace sfa

extern "C" int sfa_cgemm3x(int64_t B, int64_t Nn, int64_t Mm, int64_t K, const sfa_c64* P, int64_t ldP,
                           int64_t strideP, const sfa_c64* R, int64_t ldR, int64_t strideR, sfa_c64* out,
                           int64_t ldo, int64_t strideO, const sfa_c64* rowscale, int64_t strideS,
                           int hybrid, void* stream) {
    if (B < 0 || Nn < 0 || Mm < 0 || K <= 0) return SFA_ERR_INVALID;
    if (B == 0 || Nn == 0 || Mm == 0) return SFA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long Kp = round_up(K, 4);
    const long long pb = strideP == 0 ? 1 : B;
    float* Pp = nullptr;
    SFA_CUDA_CHECK(cudaMallocAsync(&Pp, (size_t)pb * 4 * Nn * Kp * sizeof(float), st));
    split_planar_c64_kernel<<<grid1d(pb * Nn * Kp), 256, 0, st>>>(reinterpret_cast<const float2*>(P), ldP, strideP,
                                                                   pb, Nn, K, Kp, hybrid, Pp);
    int rc = SFA_OK;
    if (cudaGetLastError() != cudaSuccess) rc = SFA_ERR_CUDA;
    void* scratch = nullptr;
    const size_t sb = use_kloop(K, strideP == 0) ? cgemm3x_kloop_scratch_bytes(K, hybrid) : 0;
    if (sb) SFA_CUDA_CHECK(cudaMallocAsync(&scratch, sb, st));
    if (rc == SFA_OK)
        rc = cgemm_splitk(B, Nn, Mm, K, Pp, Kp, strideP == 0, reinterpret_cast<const float2*>(R), ldR, strideR,
                          reinterpret_cast<float2*>(out), ldo, strideO, reinterpret_cast<const float2*>(rowscale),
                          strideS, hybrid, st, scratch, sb);
    if (scratch) cudaFreeAsync(scratch, st);
    cudaFreeAsync(Pp, st);
    return rc;
}
