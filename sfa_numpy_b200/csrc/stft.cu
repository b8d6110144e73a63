// Time <-> frequency ends of the path.
//
// Every example script of the reference ends with
//     q_t = np.fft.fftshift(np.fft.irfft(q, axis=0), axes=0)
// (doc/examples/modal_beamforming_rigid_spherical_array.py:40, ..._open_circular_array.py:40, ...): an inverse
// real FFT along the FREQUENCY axis -- the outermost axis of the beam tensor q[F, L(, T)] -- followed by a
// half-length rotation.  `sfa_irfft_axis0` does exactly that on the device; `sfa_stft` is the matching front
// end that turns time signals into spectra X[F, Q, T] directly in the layout the steering kernels read.
//
// Hand-written FFT (no cuFFT on the path):
//   * Stockham autosort radix-4 (+ one radix-2 pass for odd log2) in shared memory, ping-pong buffers,
//     twiddles from a table computed once in FP64;
//   * real transforms two at a time: two adjacent columns a, b share one complex FFT
//     (z = x_a + i x_b  <->  Z[k] = X_a[k] + i X_b[k] with the Hermitian halves), which also makes the
//     global accesses 16..64 contiguous bytes per frequency row;
//   * lengths that are not powers of two (the examples use n = 2 (100 - 1) = 198) go through Bluestein's
//     chirp-z identity on top of the power-of-two kernel (m >= 2n - 1), chirp angles reduced exactly
//     (t^2 mod 2n in integers) so that large n loses no accuracy;
//   * float for complex64 data, double for complex128 (1e-10 class parity with np.fft).
#include <type_traits>
#include <math.h>
#include "sfa_common.cuh"

namespace sfa {
namespace fft {

template <class R>
struct cx {
    R x, y;
};
template <class R>
__device__ __forceinline__ cx<R> mk(R a, R b) {
    cx<R> r;
    r.x = a;
    r.y = b;
    return r;
}
template <class R>
__device__ __forceinline__ cx<R> cmul(cx<R> a, cx<R> b) {
    return mk<R>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <class R>
__device__ __forceinline__ cx<R> cadd(cx<R> a, cx<R> b) {
    return mk<R>(a.x + b.x, a.y + b.y);
}
template <class R>
__device__ __forceinline__ cx<R> csub(cx<R> a, cx<R> b) {
    return mk<R>(a.x - b.x, a.y - b.y);
}

constexpr int kThreads = 512;

// Forward (e^{-2 pi i jk/m}) twiddle table T[i] = exp(-2 pi i * i / m), i < m, computed in FP64.
template <class R>
__global__ void twiddle_kernel(int m, cx<R>* __restrict__ T) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        double s, c;
        sincospi(-2.0 * (double)i / (double)m, &s, &c);
        T[i] = mk<R>((R)c, (R)s);
    }
}

// In-place-looking Stockham FFT of `nseq` sequences of length m (power of two) held in shared memory:
// data in `a` ([nseq][m]), scratch `b`; returns the buffer that holds the result.  kInverse conjugates the
// (forward) twiddles, i.e. computes sum_k z[k] e^{+2 pi i jk/m} (unscaled).  All threads of the CTA call it.
template <class R, bool kInverse>
__device__ cx<R>* fft_shared(cx<R>* a, cx<R>* b, const cx<R>* __restrict__ tw, int m, int log2m, int nseq) {
    int Ns = 1;
    if (log2m & 1) {
        // one radix-2 pass (Ns = 1: no twiddles)
        const int half = m >> 1;
        for (int w = threadIdx.x; w < nseq * half; w += blockDim.x) {
            const int s = w / half, j = w - s * half;
            const cx<R>* in = a + (size_t)s * m;
            cx<R>* out = b + (size_t)s * m;
            const cx<R> v0 = in[j], v1 = in[j + half];
            out[2 * j] = cadd(v0, v1);
            out[2 * j + 1] = csub(v0, v1);
        }
        __syncthreads();
        cx<R>* t = a;
        a = b;
        b = t;
        Ns = 2;
    }
    const int quarter = m >> 2;
    for (; Ns < m; Ns <<= 2) {
        const int tstep = m / (Ns * 4);                    // table stride: angle(k) = k * tstep
        for (int w = threadIdx.x; w < nseq * quarter; w += blockDim.x) {
            const int s = w / quarter, j = w - s * quarter;
            const cx<R>* in = a + (size_t)s * m;
            cx<R>* out = b + (size_t)s * m;
            const int k = j & (Ns - 1);
            cx<R> v0 = in[j], v1 = in[j + quarter], v2 = in[j + 2 * quarter], v3 = in[j + 3 * quarter];
            if (k) {
                cx<R> w1 = tw[k * tstep], w2 = tw[2 * k * tstep], w3 = tw[3 * k * tstep];
                if (kInverse) {
                    w1.y = -w1.y;
                    w2.y = -w2.y;
                    w3.y = -w3.y;
                }
                v1 = cmul(v1, w1);
                v2 = cmul(v2, w2);
                v3 = cmul(v3, w3);
            }
            const cx<R> p = cadd(v0, v2), q = csub(v0, v2), r = cadd(v1, v3), u = csub(v1, v3);
            // forward: d = -i u ; inverse: d = +i u
            const cx<R> d = kInverse ? mk<R>(-u.y, u.x) : mk<R>(u.y, -u.x);
            const int j0 = ((j - k) << 2) + k;             // (j / Ns) * Ns * 4 + k
            out[j0] = cadd(p, r);
            out[j0 + Ns] = cadd(q, d);
            out[j0 + 2 * Ns] = csub(p, r);
            out[j0 + 3 * Ns] = csub(q, d);
        }
        __syncthreads();
        cx<R>* t = a;
        a = b;
        b = t;
    }
    return a;
}

// Bluestein tables for length n (sign +: inverse DFT): chirp[t] = exp(+i pi t^2 / n), t < n, and
// Bf = FFT_m(bwrap), bwrap[t] = conj(chirp[|t|]) wrapped on m points.  One CTA.
template <class R>
__global__ void __launch_bounds__(kThreads)
bluestein_setup_kernel(int n, int m, int log2m, int sign, const cx<R>* __restrict__ tw, cx<R>* __restrict__ chirp,
                       cx<R>* __restrict__ Bf) {
    extern __shared__ unsigned char smem_raw[];
    cx<R>* a = reinterpret_cast<cx<R>*>(smem_raw);
    cx<R>* b = a + m;
    for (int t = threadIdx.x; t < m; t += blockDim.x) a[t] = mk<R>((R)0, (R)0);
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const long long t2 = ((long long)t * t) % (2LL * n);          // exact angle reduction
        double s, c;
        sincospi((double)sign * (double)t2 / (double)n, &s, &c);
        chirp[t] = mk<R>((R)c, (R)s);
        const cx<R> bc = mk<R>((R)c, (R)(-s));                         // conj(chirp)
        a[t] = bc;
        if (t) a[m - t] = bc;
    }
    __syncthreads();
    cx<R>* res = fft_shared<R, false>(a, b, tw, m, log2m, 1);
    for (int t = threadIdx.x; t < m; t += blockDim.x) Bf[t] = res[t];
}

struct IrfftParams {
    long long F, C, n;            // bins available, columns, output samples
    long long ld, ldo;            // row pitches (elements) of q and out
    int m, log2m;                 // power-of-two FFT length (m == n: direct, else Bluestein)
    int nseq;                     // column pairs per CTA tile
    int shift;                    // fftshift rotation (n / 2 or 0)
};

// out[(j + shift) % n, c] = (1/n) sum_k Xh[k, c] e^{+2 pi i jk/n},  Xh = Hermitian extension of q[0 .. n/2, c]
// (imaginary parts of the DC and -- for even n -- Nyquist bins ignored, bins >= F taken as zero: np.fft.irfft).
template <class R>
__global__ void __launch_bounds__(kThreads)
irfft_axis0_kernel(const IrfftParams prm, const cx<R>* __restrict__ q, R* __restrict__ out,
                   const cx<R>* __restrict__ tw_g, const cx<R>* __restrict__ chirp, const cx<R>* __restrict__ Bf) {
    extern __shared__ unsigned char smem_raw[];
    const int m = prm.m, nseq = prm.nseq;
    const long long n = prm.n;
    cx<R>* tw = reinterpret_cast<cx<R>*>(smem_raw);
    cx<R>* a = tw + m;
    cx<R>* b = a + (size_t)nseq * m;
    for (int i = threadIdx.x; i < m; i += blockDim.x) tw[i] = tw_g[i];
    const bool direct = (m == n);
    const int cols = 2 * nseq;
    const long long tiles = (prm.C + cols - 1) / cols;
    const long long kmax = n / 2;                       // highest bin index used
    const R scale = (R)(1.0 / (double)n);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long c0 = tile * cols;
        __syncthreads();                                 // previous tile's readers of a/b are done; tw is loaded
        // ---- zero, then load + Hermitian pack: Z[k] = Xa[k] + i Xb[k], Z[n-k] = conj(Xa[k]) + i conj(Xb[k])
        for (int i = threadIdx.x; i < nseq * m; i += blockDim.x) a[i] = mk<R>((R)0, (R)0);
        __syncthreads();
        for (long long w = threadIdx.x; w < (kmax + 1) * nseq; w += blockDim.x) {
            const long long k = w / nseq;
            const int s = (int)(w - k * nseq);
            const long long ca = c0 + 2 * s, cb = ca + 1;
            cx<R> xa = mk<R>((R)0, (R)0), xb = xa;
            if (k < prm.F) {
                if (ca < prm.C) xa = q[k * prm.ld + ca];
                if (cb < prm.C) xb = q[k * prm.ld + cb];
            }
            const bool self_conj = (k == 0) || (2 * k == n);
            if (self_conj) {
                xa.y = (R)0;
                xb.y = (R)0;
            }
            cx<R> zk = mk<R>(xa.x - xb.y, xa.y + xb.x);                   // Xa + i Xb
            cx<R> zn = mk<R>(xa.x + xb.y, xb.x - xa.y);                   // conj(Xa) + i conj(Xb)
            cx<R>* seq = a + (size_t)s * m;
            if (!direct) {                                                 // Bluestein: a[k] = Z[k] chirp[k]
                zk = cmul(zk, chirp[k]);
                if (!self_conj) zn = cmul(zn, chirp[n - k]);
            }
            seq[k] = zk;
            if (!self_conj) seq[n - k] = zn;
        }
        __syncthreads();
        cx<R>* res;
        if (direct) {
            res = fft_shared<R, true>(a, b, tw, m, prm.log2m, nseq);
        } else {
            cx<R>* fa = fft_shared<R, false>(a, b, tw, m, prm.log2m, nseq);
            cx<R>* other = (fa == a) ? b : a;
            for (int i = threadIdx.x; i < nseq * m; i += blockDim.x) fa[i] = cmul(fa[i], Bf[i & (m - 1)]);
            __syncthreads();
            res = fft_shared<R, true>(fa, other, tw, m, prm.log2m, nseq);
        }
        // ---- unpack: x_a[j] = Re z[j] / n, x_b[j] = Im z[j] / n (Bluestein: times chirp[j] / m first)
        const R bscale = direct ? scale : (R)(1.0 / ((double)n * (double)m));
        for (long long w = threadIdx.x; w < n * nseq; w += blockDim.x) {
            const long long j = w / nseq;
            const int s = (int)(w - j * nseq);
            cx<R> z = res[(size_t)s * m + j];
            if (!direct) z = cmul(z, chirp[j]);
            long long jo = j + prm.shift;
            if (jo >= n) jo -= n;
            const long long ca = c0 + 2 * s;
            R* o = out + jo * prm.ldo + ca;
            if (ca < prm.C) o[0] = z.x * bscale;
            if (ca + 1 < prm.C) o[1] = z.y * bscale;
        }
    }
}

struct StftParams {
    long long Q, S, T;            // channels, samples per channel, frames
    int nfft, hop, log2n;
    long long Fk;                 // bins kept (<= nfft/2 + 1)
    int nseq;                     // frame pairs per CTA tile
};

// X[f, q, t] = sum_j win[j] x[q, t*hop + j] e^{-2 pi i f j / nfft}   (np.fft.rfft of the windowed frame)
template <class R>
__global__ void __launch_bounds__(kThreads)
stft_kernel(const StftParams prm, const R* __restrict__ x, const R* __restrict__ win, cx<R>* __restrict__ X,
            const cx<R>* __restrict__ tw_g) {
    extern __shared__ unsigned char smem_raw[];
    const int n = prm.nfft, nseq = prm.nseq;
    cx<R>* tw = reinterpret_cast<cx<R>*>(smem_raw);
    cx<R>* a = tw + n;
    cx<R>* b = a + (size_t)nseq * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tw[i] = tw_g[i];
    const int frames = 2 * nseq;
    const long long ttiles = (prm.T + frames - 1) / frames;
    const long long tiles = prm.Q * ttiles;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long qi = tile / ttiles;
        const long long t0 = (tile - qi * ttiles) * frames;
        __syncthreads();
        const R* xq = x + qi * prm.S;
        for (int w = threadIdx.x; w < nseq * n; w += blockDim.x) {
            const int s = w / n, j = w - s * n;
            const long long ta = t0 + 2 * s, tb = ta + 1;
            const R wj = win ? win[j] : (R)1;
            const R va = (ta < prm.T) ? xq[ta * prm.hop + j] * wj : (R)0;
            const R vb = (tb < prm.T) ? xq[tb * prm.hop + j] * wj : (R)0;
            a[w] = mk<R>(va, vb);
        }
        __syncthreads();
        cx<R>* res = fft_shared<R, false>(a, b, tw, n, prm.log2n, nseq);
        // Xa[k] = (Z[k] + conj(Z[n-k])) / 2,  Xb[k] = (Z[k] - conj(Z[n-k])) / (2i)
        for (long long w = threadIdx.x; w < prm.Fk * nseq; w += blockDim.x) {
            const long long k = w / nseq;
            const int s = (int)(w - k * nseq);
            const cx<R>* z = res + (size_t)s * n;
            const cx<R> zk = z[k], zn = z[(n - k) & (n - 1)];
            const cx<R> xa = mk<R>((R)0.5 * (zk.x + zn.x), (R)0.5 * (zk.y - zn.y));
            const cx<R> xb = mk<R>((R)0.5 * (zk.y + zn.y), (R)0.5 * (zn.x - zk.x));
            const long long ta = t0 + 2 * s;
            cx<R>* o = X + (k * prm.Q + qi) * prm.T + ta;
            if (ta < prm.T) o[0] = xa;
            if (ta + 1 < prm.T) o[1] = xb;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Fast path (float, power-of-two length 256 .. 4096): radix-16 passes with the butterflies in REGISTERS.
//   * the sequences of a tile live in ONE shared-memory buffer (no ping-pong): a pass reads the 16 (or 2/4/8)
//     inputs of a butterfly into registers, the CTA synchronises, and the results go back in Stockham order --
//     2048 points are 16 x 16 x 8: three passes instead of the six radix-4/2 passes of the generic kernel, a third
//     of the barriers and of the shared-memory traffic;
//   * index i of a sequence sits at i + i / 16, and the sequences are 2 elements further apart than that:
//     the stride-16 writes of the first pass, the strided reads of the later ones and the pair-interleaved loads
//     and stores of the global <-> shared phases are all free of bank conflicts;
//   * half the footprint per column lets a tile hold 8 column pairs: every frequency row of the tile is one
//     128-byte line of q (64 bytes of the real output).
// ------------------------------------------------------------------------------------------------------------
namespace fast {

// 256 threads: the radix-16 butterflies with their composed twiddles want ~240 registers; with 512 threads (128
// registers) the passes spilled 260 bytes per thread.  256 threads without spills measure 6-15 % faster on every
// length (n = 2048: 2.31 -> 2.16 ms on the headline shape, n = 256: 6.08 -> 5.18 ms, STFT 0.83 -> 0.77 ms); 384
// threads (168 registers, 104 bytes of spills, uneven rounds) are slower than both.
#ifndef SFA_FFT_THREADS
#define SFA_FFT_THREADS 256
#endif
#ifndef SFA_FFT_PAIRS
#define SFA_FFT_PAIRS 8
#endif
#ifndef SFA_FFT_CTAS
#define SFA_FFT_CTAS 1
#endif
#ifndef SFA_FFT_LOAD_BATCH
#define SFA_FFT_LOAD_BATCH 16
#endif
constexpr int kLoadBatch = SFA_FFT_LOAD_BATCH;   // independent 16-byte loads a thread has in flight in the load phase
constexpr int kT = SFA_FFT_THREADS;
constexpr int kCtasPerSm = SFA_FFT_CTAS;
__host__ __device__ __forceinline__ int pad_idx(int i) { return i + (i >> 4); }
__host__ __device__ __forceinline__ int seq_stride(int m) { return m + (m >> 4) + 2; }

typedef cx<float> cf;

// cos / sin of pi t / 8 for t = 1 .. 7 (compile-time t after unrolling: these fold to immediates)
__device__ __forceinline__ constexpr float cos_pi8(int t) {
    return t == 1 ? 0.92387953251128674f : t == 2 ? 0.70710678118654752f : t == 3 ? 0.38268343236508977f
         : t == 5 ? -0.38268343236508977f : t == 6 ? -0.70710678118654752f : t == 7 ? -0.92387953251128674f : 0.f;
}
__device__ __forceinline__ constexpr float sin_pi8(int t) {
    return t == 1 ? 0.38268343236508977f : t == 2 ? 0.70710678118654752f : t == 3 ? 0.92387953251128674f
         : t == 5 ? 0.92387953251128674f : t == 6 ? 0.70710678118654752f : t == 7 ? 0.38268343236508977f : 1.f;
}

template <int R>
__device__ __forceinline__ int brev(int i) {
    int r = 0;
#pragma unroll
    for (int b = 1; b < R; b <<= 1) {
        r = (r << 1) | (i & 1);
        i >>= 1;
    }
    return r;
}

// In-register DFT of R points (decimation in frequency): afterwards v[i] holds output brev(i).
// kInv: e^{+2 pi i jk/R}.  All loops unroll; the twiddles are compile-time constants.
template <int R, bool kInv>
__device__ __forceinline__ void dft_reg(cf (&v)[R]) {
#pragma unroll
    for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int base = 0; base < R; base += 2 * half) {
#pragma unroll
            for (int jj = 0; jj < half; ++jj) {
                const cf a = v[base + jj], b = v[base + jj + half];
                v[base + jj] = cadd(a, b);
                const cf d = csub(a, b);
                const int t = jj * (8 / half);                  // angle = pi t / 8
                if (t == 0) {
                    v[base + jj + half] = d;
                } else if (t == 4) {                            // -i (forward) / +i (inverse)
                    v[base + jj + half] = kInv ? mk<float>(-d.y, d.x) : mk<float>(d.y, -d.x);
                } else {
                    const float c = cos_pi8(t), sn = kInv ? sin_pi8(t) : -sin_pi8(t);
                    v[base + jj + half] = mk<float>(d.x * c - d.y * sn, d.x * sn + d.y * c);
                }
            }
        }
    }
}

// One Stockham pass of radix R over `nseq` sequences of length m (in place, see above).  Ns = product of the
// radices of the earlier passes (compile time).  blockDim.x = kT, 16 <= m / R <= kT.  Ends with a barrier.
// All 2 R shared-memory accesses of a butterfly are base + r * stride with strides that do not depend on the
// thread: inputs j + r m/R sit at pad(j) + r (m/R + m/(16 R)), outputs j0 + c Ns at pad(j0) + c (Ns + Ns/16)
// (Ns = 1: the 16 outputs of the first pass are adjacent) -- no per-access padding arithmetic.
template <int R, bool kInv, int Ns>
__device__ __forceinline__ void pass(cf* data, int sstride, const cf* __restrict__ tw, int m, int nseq) {
    const int per_seq = m / R;
    const int G = kT / per_seq;                                 // sequences per round
    const int lps = 31 - __clz(per_seq);                        // per_seq is a power of two
    const int s_in = (int)threadIdx.x >> lps, j = (int)threadIdx.x & (per_seq - 1);
    const int k = j & (Ns - 1);
    const int tstep = m / (Ns * R);
    const int j0 = (j - k) * R + k;
    const int rbase = pad_idx(j), rstride = per_seq + (per_seq >> 4);
    const int wbase = pad_idx(j0);
    constexpr int wstride = Ns >= 16 ? Ns + Ns / 16 : Ns;
    for (int s0 = 0; s0 < nseq; s0 += G) {
        const int s = s0 + s_in;
        const bool act = s < nseq;
        cf v[R];
        cf* const a = data + s * sstride;
        if (act) {
            const cf* rd = a + rbase;
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = rd[r * rstride];
            if (Ns > 1 && k) {
                // twiddles w^(r k tstep), r = 1 .. R-1: the powers of two come from the table (their addresses
                // r k tstep are far apart -- up to 16-way bank conflicts if all R-1 were loaded), the others
                // are products of them (error ~2e-7, well below the FP32 transform's own)
                cf w1 = tw[k * tstep];
                if (kInv) w1.y = -w1.y;
                v[1] = cmul(v[1], w1);
                if (R > 2) {
                    cf w2 = tw[2 * k * tstep];
                    if (kInv) w2.y = -w2.y;
                    const cf w3 = cmul(w2, w1);
                    v[2] = cmul(v[2], w2);
                    v[3] = cmul(v[3], w3);
                    if (R > 4) {
                        cf w4 = tw[4 * k * tstep];
                        if (kInv) w4.y = -w4.y;
                        v[4] = cmul(v[4], w4);
                        v[5] = cmul(v[5], cmul(w4, w1));
                        v[6] = cmul(v[6], cmul(w4, w2));
                        v[7] = cmul(v[7], cmul(w4, w3));
                        if (R > 8) {
                            cf w8 = tw[8 * k * tstep];
                            if (kInv) w8.y = -w8.y;
                            const cf w12 = cmul(w8, w4);
                            v[8 % R] = cmul(v[8 % R], w8);
                            v[9 % R] = cmul(v[9 % R], cmul(w8, w1));
                            v[10 % R] = cmul(v[10 % R], cmul(w8, w2));
                            v[11 % R] = cmul(v[11 % R], cmul(w8, w3));
                            v[12 % R] = cmul(v[12 % R], w12);
                            v[13 % R] = cmul(v[13 % R], cmul(w12, w1));
                            v[14 % R] = cmul(v[14 % R], cmul(w12, w2));
                            v[15 % R] = cmul(v[15 % R], cmul(w12, w3));
                        }
                    }
                }
            }
            dft_reg<R, kInv>(v);
        }
        __syncthreads();                                        // every butterfly of the round has its inputs
        if (act) {
            cf* wr = a + wbase;
#pragma unroll
            for (int r = 0; r < R; ++r) wr[brev<R>(r) * wstride] = v[r];
        }
    }
    __syncthreads();
}

// all passes for length m = 2^log2m (8 <= log2m <= 12): 16 x 16 x {1, 2, 4, 8, 16}
template <bool kInv>
__device__ __forceinline__ void fft_inplace(cf* data, int sstride, const cf* tw, int m, int log2m, int nseq) {
    pass<16, kInv, 1>(data, sstride, tw, m, nseq);
    pass<16, kInv, 16>(data, sstride, tw, m, nseq);
    switch (log2m) {
        case 9: pass<2, kInv, 256>(data, sstride, tw, m, nseq); break;
        case 10: pass<4, kInv, 256>(data, sstride, tw, m, nseq); break;
        case 11: pass<8, kInv, 256>(data, sstride, tw, m, nseq); break;
        case 12: pass<16, kInv, 256>(data, sstride, tw, m, nseq); break;
        default: break;
    }
}

__global__ void __launch_bounds__(kT, kCtasPerSm)
irfft_fast_kernel(const IrfftParams prm, const cf* __restrict__ q, float* __restrict__ out, const cf* __restrict__ tw_g) {
    extern __shared__ unsigned char smem_raw[];
    const int m = prm.m, nseq = prm.nseq;                        // nseq is a power of two
    const int lg = 31 - __clz(nseq);
    const int n = (int)prm.n;                                   // == m
    const int sstride = seq_stride(m);
    cf* tw = reinterpret_cast<cf*>(smem_raw);
    cf* data = tw + m;
    for (int i = threadIdx.x; i < m; i += kT) tw[i] = tw_g[i];
    const int cols = 2 * nseq;
    const long long tiles = (prm.C + cols - 1) / cols;
    const int kmax = n / 2;
    const float scale = (float)(1.0 / (double)n);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long c0 = tile * cols;
        __syncthreads();                                        // previous tile's stores have read `data`; tw is loaded
        // load + Hermitian pack (every index 0 .. n-1 of every sequence is written: no zero fill needed).
        // Eight independent 16-byte loads per thread are in flight before the first one is used: with one load
        // per loop iteration this phase was a chain of ~16 DRAM latencies per tile.
        const int items = (kmax + 1) * nseq;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(q) & 15) == 0) && ((prm.ld & 1) == 0);
        auto load_pack = [&](auto batch_tag) {
        constexpr int kB = decltype(batch_tag)::value;
        for (int w0 = threadIdx.x; w0 < items; w0 += kT * kB) {
            float4 v[kB];
#pragma unroll
            for (int u = 0; u < kB; ++u) {
                const int w = w0 + u * kT;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (w < items) {
                    const int k = w >> lg, s = w & (nseq - 1);
                    const long long ca = c0 + 2 * s;
                    if (k < prm.F && ca < prm.C) {
                        const cf* row = q + (long long)k * prm.ld + ca;
                        if (vec_ok && ca + 1 < prm.C) {
                            v[u] = __ldg(reinterpret_cast<const float4*>(row));
                        } else {
                            const float2 a = __ldg(reinterpret_cast<const float2*>(row));
                            v[u].x = a.x;
                            v[u].y = a.y;
                            if (ca + 1 < prm.C) {
                                const float2 b = __ldg(reinterpret_cast<const float2*>(row + 1));
                                v[u].z = b.x;
                                v[u].w = b.y;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kB; ++u) {
                const int w = w0 + u * kT;
                if (w < items) {
                    const int k = w >> lg, s = w & (nseq - 1);
                    cf xa = mk<float>(v[u].x, v[u].y), xb = mk<float>(v[u].z, v[u].w);
                    const bool self_conj = (k == 0) || (2 * k == n);
                    if (self_conj) {
                        xa.y = 0.f;
                        xb.y = 0.f;
                    }
                    cf* seq = data + (size_t)s * sstride;
                    seq[pad_idx(k)] = mk<float>(xa.x - xb.y, xa.y + xb.x);                       // Xa + i Xb
                    if (!self_conj) seq[pad_idx(n - k)] = mk<float>(xa.x + xb.y, xb.x - xa.y);   // conj(Xa) + i conj(Xb)
                }
            }
        }
        };
        // long sequences: 16 loads in flight per thread (n = 2048: 2.19 -> 2.09 ms); short ones have too few rows per
        // tile for that (n = 256: 5.2 ms with 8, 5.5 ms with 16)
        if (m >= 1024) load_pack(std::integral_constant<int, kLoadBatch>{});
        else load_pack(std::integral_constant<int, 8>{});
        __syncthreads();
        fft_inplace<true>(data, sstride, tw, m, prm.log2m, nseq);
        // unpack: x_a[j] = Re z[j] / n, x_b[j] = Im z[j] / n
        // (whole tiles whose rows are 16-byte aligned: two sequences = four real columns per store)
        if (nseq >= 2 && c0 + cols <= prm.C && ((prm.ldo | c0) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            const int hs = nseq >> 1, lgh = lg - 1;
#pragma unroll 4
            for (int w = threadIdx.x; w < n * hs; w += kT) {
                const int j = w >> lgh, s = (w & (hs - 1)) << 1;
                const cf z0 = data[(size_t)s * sstride + pad_idx(j)];
                const cf z1 = data[(size_t)(s + 1) * sstride + pad_idx(j)];
                int jo = j + prm.shift;
                if (jo >= n) jo -= n;
                *reinterpret_cast<float4*>(out + (long long)jo * prm.ldo + c0 + 2 * s) =
                    make_float4(z0.x * scale, z0.y * scale, z1.x * scale, z1.y * scale);
            }
        } else
#pragma unroll 4
        for (int w = threadIdx.x; w < n * nseq; w += kT) {
            const int j = w >> lg, s = w & (nseq - 1);
            const cf z = data[(size_t)s * sstride + pad_idx(j)];
            int jo = j + prm.shift;
            if (jo >= n) jo -= n;
            const long long ca = c0 + 2 * s;
            float* o = out + (long long)jo * prm.ldo + ca;
            if (ca + 1 < prm.C && ((reinterpret_cast<uintptr_t>(o) & 7) == 0)) {
                *reinterpret_cast<float2*>(o) = make_float2(z.x * scale, z.y * scale);
            } else {
                if (ca < prm.C) o[0] = z.x * scale;
                if (ca + 1 < prm.C) o[1] = z.y * scale;
            }
        }
    }
}

__global__ void __launch_bounds__(kT, kCtasPerSm)
stft_fast_kernel(const StftParams prm, const float* __restrict__ x, const float* __restrict__ win, cf* __restrict__ X,
                 const cf* __restrict__ tw_g) {
    extern __shared__ unsigned char smem_raw[];
    const int n = prm.nfft, nseq = prm.nseq;                     // both powers of two
    const int lg = 31 - __clz(nseq);
    const int sstride = seq_stride(n);
    cf* tw = reinterpret_cast<cf*>(smem_raw);
    cf* data = tw + n;
    for (int i = threadIdx.x; i < n; i += kT) tw[i] = tw_g[i];
    const int frames = 2 * nseq;
    const long long ttiles = (prm.T + frames - 1) / frames;
    const long long tiles = prm.Q * ttiles;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long qi = tile / ttiles;
        const long long t0 = (tile - qi * ttiles) * frames;
        __syncthreads();
        const float* xq = x + qi * prm.S;
        // four consecutive samples per load when everything is 16-byte aligned (hop and signal length multiples of 4):
        // a quarter of the load instructions of the scalar loop below
        const bool vec4 = ((prm.hop & 3) == 0) && ((prm.S & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                          (!win || (reinterpret_cast<uintptr_t>(win) & 15) == 0);
        if (vec4) {
            const int lq = prm.log2n - 2;
#pragma unroll 4
            for (int w = threadIdx.x; w < nseq * (n >> 2); w += kT) {
                const int s = w >> lq, j = (w & ((n >> 2) - 1)) << 2;
                const long long ta = t0 + 2 * s, tb = ta + 1;
                const float4 wv = win ? __ldg(reinterpret_cast<const float4*>(win + j)) : make_float4(1.f, 1.f, 1.f, 1.f);
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 va = (ta < prm.T) ? __ldg(reinterpret_cast<const float4*>(xq + ta * prm.hop + j)) : z4;
                const float4 vb = (tb < prm.T) ? __ldg(reinterpret_cast<const float4*>(xq + tb * prm.hop + j)) : z4;
                cf* d = data + (size_t)s * sstride + pad_idx(j);             // j .. j + 3 lie in one 16-entry block
                d[0] = mk<float>(va.x * wv.x, vb.x * wv.x);
                d[1] = mk<float>(va.y * wv.y, vb.y * wv.y);
                d[2] = mk<float>(va.z * wv.z, vb.z * wv.z);
                d[3] = mk<float>(va.w * wv.w, vb.w * wv.w);
            }
        } else
#pragma unroll 8
        for (int w = threadIdx.x; w < nseq * n; w += kT) {
            const int s = w >> prm.log2n, j = w & (n - 1);      // j fastest: coalesced along the signal
            const long long ta = t0 + 2 * s, tb = ta + 1;
            const float wj = win ? __ldg(win + j) : 1.f;
            const float va = (ta < prm.T) ? __ldg(xq + ta * prm.hop + j) * wj : 0.f;
            const float vb = (tb < prm.T) ? __ldg(xq + tb * prm.hop + j) * wj : 0.f;
            data[(size_t)s * sstride + pad_idx(j)] = mk<float>(va, vb);
        }
        __syncthreads();
        fft_inplace<false>(data, sstride, tw, n, prm.log2n, nseq);
        // Xa[k] = (Z[k] + conj(Z[n-k])) / 2,  Xb[k] = (Z[k] - conj(Z[n-k])) / (2i); s fastest: 16 bytes per pair
        for (long long w = threadIdx.x; w < prm.Fk * nseq; w += kT) {
            const int k = (int)(w >> lg), s = (int)(w & (nseq - 1));
            const cf* z = data + (size_t)s * sstride;
            const cf zk = z[pad_idx(k)], zn = z[pad_idx((n - k) & (n - 1))];
            const cf xa = mk<float>(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
            const cf xb = mk<float>(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
            const long long ta = t0 + 2 * s;
            cf* o = X + ((long long)k * prm.Q + qi) * prm.T + ta;
            if (ta + 1 < prm.T && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                *reinterpret_cast<float4*>(o) = make_float4(xa.x, xa.y, xb.x, xb.y);     // frames ta, ta + 1: 16 bytes
            } else {
                if (ta < prm.T) o[0] = xa;
                if (ta + 1 < prm.T) o[1] = xb;
            }
        }
    }
}

static inline bool supported(long long m) { return m >= 256 && m <= 4096 && (m & (m - 1)) == 0; }
static inline int pairs_per_tile(long long m) { return m <= 2048 ? SFA_FFT_PAIRS : SFA_FFT_PAIRS / 2; }
static inline size_t smem_bytes(long long m, int nseq) { return ((size_t)m + (size_t)nseq * seq_stride((int)m)) * sizeof(cf); }

}  // namespace fast

static int ilog2(long long v) {
    int l = 0;
    while ((1LL << l) < v) ++l;
    return l;
}

template <class R>
static int run_irfft(long long F, long long C, long long n, const void* q, long long ld, void* out, long long ldo,
                     int shift, void* work, size_t work_bytes, cudaStream_t st) {
    const bool pow2 = (n & (n - 1)) == 0;
    const long long m = pow2 ? n : (1LL << ilog2(2 * n - 1));
    const int log2m = ilog2(m);
    const size_t esz = sizeof(cx<R>);
    if (m > 16384 || m < 2) return SFA_ERR_UNSUPPORTED;
    const size_t need = (size_t)m * esz + (pow2 ? 0 : (size_t)(n + m) * esz);
    if (work_bytes < need || !work) return SFA_ERR_WORKSPACE;
    cx<R>* tw = reinterpret_cast<cx<R>*>(work);
    cx<R>* chirp = tw + m;
    cx<R>* Bf = chirp + n;
    twiddle_kernel<R><<<(unsigned)ceil_div(m, 256), 256, 0, st>>>((int)m, tw);
    SFA_LAUNCH_CHECK();
    if (!pow2) {
        const size_t smem = 2 * (size_t)m * esz;
        SFA_ENSURE_DYN_SMEM(bluestein_setup_kernel<R>, smem);
        bluestein_setup_kernel<R><<<1, kThreads, smem, st>>>((int)n, (int)m, log2m, +1, tw, chirp, Bf);
        SFA_LAUNCH_CHECK();
    }
    if (sizeof(R) == 4 && pow2 && fast::supported(m) && !options().fft_generic) {
        long long nseq = fast::pairs_per_tile(m);
        while (nseq > 1 && nseq > (C + 1) / 2) nseq >>= 1;      // a power of two (index math by shifts)
        IrfftParams prm;
        prm.F = F;
        prm.C = C;
        prm.n = n;
        prm.ld = ld;
        prm.ldo = ldo;
        prm.m = (int)m;
        prm.log2m = log2m;
        prm.nseq = (int)nseq;
        prm.shift = shift ? (int)(n / 2) : 0;
        const size_t smem = fast::smem_bytes(m, (int)nseq);
        SFA_ENSURE_DYN_SMEM(fast::irfft_fast_kernel, smem);
        long long blocks = ceil_div(C, 2 * nseq);
        if (blocks > (long long)sm_count() * fast::kCtasPerSm) blocks = (long long)sm_count() * fast::kCtasPerSm;
        prof_begin(st);
        fast::irfft_fast_kernel<<<(unsigned)blocks, fast::kT, smem, st>>>(prm, reinterpret_cast<const fast::cf*>(q),
                                                                          reinterpret_cast<float*>(out),
                                                                          reinterpret_cast<const fast::cf*>(tw));
        prof_end(st);
        SFA_LAUNCH_CHECK();
        return SFA_OK;
    }
    // column pairs per tile: as many as fit in ~200 KB next to the twiddle table (capped at 8 = 16 columns)
    const size_t budget = 200 * 1024;
    long long nseq = ((long long)budget - (long long)(m * esz)) / (long long)(2 * m * esz);
    if (nseq < 1) return SFA_ERR_UNSUPPORTED;
    if (nseq > 8) nseq = 8;
    if (nseq > (C + 1) / 2) nseq = (C + 1) / 2;
    // two resident CTAs per SM hide the barrier-heavy passes better than one fat one, when they fit
    if (nseq >= 4 && (size_t)m * esz + (size_t)nseq * m * esz <= 100 * 1024) nseq /= 2;
    IrfftParams prm;
    prm.F = F;
    prm.C = C;
    prm.n = n;
    prm.ld = ld;
    prm.ldo = ldo;
    prm.m = (int)m;
    prm.log2m = log2m;
    prm.nseq = (int)nseq;
    prm.shift = shift ? (int)(n / 2) : 0;
    const size_t smem = (size_t)m * esz + 2 * (size_t)nseq * m * esz;
    SFA_ENSURE_DYN_SMEM(irfft_axis0_kernel<R>, smem);
    const long long tiles = ceil_div(C, 2 * nseq);
    const long long per_sm = (smem <= 100 * 1024) ? 2 : 1;
    long long blocks = tiles;
    if (blocks > sm_count() * per_sm) blocks = sm_count() * per_sm;
    prof_begin(st);
    irfft_axis0_kernel<R><<<(unsigned)blocks, kThreads, smem, st>>>(prm, reinterpret_cast<const cx<R>*>(q),
                                                                    reinterpret_cast<R*>(out), tw, chirp, Bf);
    prof_end(st);
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

template <class R>
static int run_stft(long long Q, long long S, int nfft, int hop, long long Fk, const void* x, const void* win,
                    void* X, void* work, size_t work_bytes, cudaStream_t st) {
    const size_t esz = sizeof(cx<R>);
    if (work_bytes < (size_t)nfft * esz || !work) return SFA_ERR_WORKSPACE;
    const long long T = 1 + (S - nfft) / hop;
    cx<R>* tw = reinterpret_cast<cx<R>*>(work);
    twiddle_kernel<R><<<(unsigned)ceil_div(nfft, 256), 256, 0, st>>>(nfft, tw);
    SFA_LAUNCH_CHECK();
    if (sizeof(R) == 4 && fast::supported(nfft) && !options().fft_generic) {
        long long nseq = fast::pairs_per_tile(nfft);
        while (nseq > 1 && nseq > (T + 1) / 2) nseq >>= 1;
        StftParams prm;
        prm.Q = Q;
        prm.S = S;
        prm.T = T;
        prm.nfft = nfft;
        prm.hop = hop;
        prm.log2n = ilog2(nfft);
        prm.Fk = Fk;
        prm.nseq = (int)nseq;
        const size_t smem = fast::smem_bytes(nfft, (int)nseq);
        SFA_ENSURE_DYN_SMEM(fast::stft_fast_kernel, smem);
        long long blocks = Q * ceil_div(T, 2 * nseq);
        if (blocks > (long long)sm_count() * fast::kCtasPerSm) blocks = (long long)sm_count() * fast::kCtasPerSm;
        fast::stft_fast_kernel<<<(unsigned)blocks, fast::kT, smem, st>>>(prm, reinterpret_cast<const float*>(x),
                                                                         reinterpret_cast<const float*>(win),
                                                                         reinterpret_cast<fast::cf*>(X),
                                                                         reinterpret_cast<const fast::cf*>(tw));
        SFA_LAUNCH_CHECK();
        return SFA_OK;
    }
    const size_t budget = 200 * 1024;
    long long nseq = ((long long)budget - (long long)(nfft * esz)) / (long long)(2 * nfft * esz);
    if (nseq < 1) return SFA_ERR_UNSUPPORTED;
    if (nseq > 8) nseq = 8;
    if (nseq > (T + 1) / 2) nseq = (T + 1) / 2;
    StftParams prm;
    prm.Q = Q;
    prm.S = S;
    prm.T = T;
    prm.nfft = nfft;
    prm.hop = hop;
    prm.log2n = ilog2(nfft);
    prm.Fk = Fk;
    prm.nseq = (int)nseq;
    const size_t smem = (size_t)nfft * esz + 2 * (size_t)nseq * nfft * esz;
    SFA_ENSURE_DYN_SMEM(stft_kernel<R>, smem);
    const long long tiles = Q * ceil_div(T, 2 * nseq);
    long long blocks = tiles;
    const long long per_sm = (smem <= 100 * 1024) ? 2 : 1;
    if (blocks > sm_count() * per_sm) blocks = sm_count() * per_sm;
    stft_kernel<R><<<(unsigned)blocks, kThreads, smem, st>>>(prm, reinterpret_cast<const R*>(x),
                                                             reinterpret_cast<const R*>(win),
                                                             reinterpret_cast<cx<R>*>(X), tw);
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

}  // namespace fft
}  // namespace sfa

using namespace sfa;

extern "C" size_t sfa_irfft_work_bytes(int64_t n, int is_double) {
    if (n < 1) return 0;
    const bool pow2 = (n & (n - 1)) == 0;
    const long long m = pow2 ? n : (1LL << fft::ilog2(2 * n - 1));
    const size_t esz = is_double ? 16 : 8;
    return (size_t)m * esz + (pow2 ? 0 : (size_t)(n + m) * esz) + 256;
}

extern "C" int sfa_irfft_axis0(int64_t F, int64_t C, int64_t n, const void* q, int64_t ld, void* out, int64_t ldo,
                               int is_double, int shift, void* work, size_t work_bytes, void* stream) {
    if (F < 0 || C < 0 || n < 0 || ld < C || ldo < C) return SFA_ERR_INVALID;
    if (n == 0) n = 2 * (F - 1);                         // np.fft.irfft default
    if (n < 1) return SFA_ERR_INVALID;                   // numpy: "Invalid number of FFT data points"
    if (C == 0) return SFA_OK;
    if (!q && F > 0) return SFA_ERR_INVALID;
    if (!out) return SFA_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 1) n = 1;
    return is_double ? fft::run_irfft<double>(F, C, n, q, ld, out, ldo, shift, work, work_bytes, st)
                     : fft::run_irfft<float>(F, C, n, q, ld, out, ldo, shift, work, work_bytes, st);
}

extern "C" size_t sfa_stft_work_bytes(int nfft, int is_double) {
    return nfft > 0 ? (size_t)nfft * (is_double ? 16 : 8) + 256 : 0;
}

extern "C" int sfa_stft(int64_t Q, int64_t S, int nfft, int hop, int64_t F_keep, const void* x, const void* window,
                        void* X, int is_double, void* work, size_t work_bytes, void* stream) {
    if (Q < 0 || nfft < 2 || (nfft & (nfft - 1)) || nfft > 16384 || hop < 1 || S < nfft) return SFA_ERR_INVALID;
    if (F_keep <= 0) F_keep = nfft / 2 + 1;
    if (F_keep > nfft / 2 + 1) return SFA_ERR_INVALID;
    if (Q == 0) return SFA_OK;
    if (!x || !X) return SFA_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    return is_double ? fft::run_stft<double>(Q, S, nfft, hop, F_keep, x, window, X, work, work_bytes, st)
                     : fft::run_stft<float>(Q, S, nfft, hop, F_keep, x, window, X, work, work_bytes, st);
}
