// FP64 special-function and complex arithmetic shared by the stage-1/2 kernels.
//
// Everything here is `SFA_HD` (host + device) so that the exact same source is
// (a) compiled by nvcc into the sm_100a kernels and (b) compiled by g++ into
// tests/hostmath (a TEST harness, never loaded by the product path) so that the
// recurrences can be checked against the reference's golden vectors on the CPU
// build box, which has no GPU.
//
// Reference semantics being reproduced (paths relative to /root/reference):
//   micarray/modal/radial.py:114-162   weights()            j_n, j_n', h_n^(2), h_n^(2)'
//   micarray/modal/radial.py:393-441   circ_radial_weights  J_n, J_n', H_n^(2), H_n^(2)'
//   micarray/modal/radial.py:216-266   regularize()
// The reference delegates j_n/y_n/J_n/Y_n to scipy.special (not vendored); the
// algorithms below are the textbook ones: Miller backward recurrence for the
// minimal solutions (j_n, J_n), forward recurrence for the dominant ones
// (y_n, Y_n), Neumann series (A&S 9.1.88/9.1.89) for Y_0 and Y_1.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SFA_HD __host__ __device__ __forceinline__
#else
#define SFA_HD inline
#endif

namespace sfa {

struct c128 {
    double re, im;
};

SFA_HD c128 mk(double re, double im) {
    c128 z;
    z.re = re;
    z.im = im;
    return z;
}
// NumPy's complex multiply: (ac - bd, ad + bc), no Annex-G recovery.
SFA_HD c128 cmul(c128 a, c128 b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
SFA_HD c128 cadd(c128 a, c128 b) { return mk(a.re + b.re, a.im + b.im); }
SFA_HD c128 csub(c128 a, c128 b) { return mk(a.re - b.re, a.im - b.im); }
SFA_HD c128 cscale(c128 a, double s) { return mk(a.re * s, a.im * s); }
SFA_HD double cabs(c128 a) { return hypot(a.re, a.im); }
// NumPy's complex divide (Smith's algorithm, npymath/loops: nc_quot).
SFA_HD c128 cdiv(c128 a, c128 b) {
    const double abr = fabs(b.re), abi = fabs(b.im);
    if (abr >= abi) {
        if (abr == 0.0 && abi == 0.0) return mk(a.re / abr, a.im / abr);
        const double rat = b.im / b.re;
        const double scl = 1.0 / (b.re + b.im * rat);
        return mk((a.re + a.im * rat) * scl, (a.im - a.re * rat) * scl);
    }
    const double rat = b.re / b.im;
    const double scl = 1.0 / (b.im + b.re * rat);
    return mk((a.re * rat + a.im) * scl, (a.im * rat - a.re) * scl);
}
// |z| < t with np.abs semantics (hypot), without paying for hypot unless z is borderline.
SFA_HD bool cabs_less(c128 a, double t) {
    const double m = fmax(fabs(a.re), fabs(a.im));
    if (!(m < t)) return false;                  // hypot >= max component (also catches NaN)
    if (m < 0.70710678 * t) return true;         // hypot <= sqrt(2) * max component
    return hypot(a.re, a.im) < t;
}
// exact powers of i (radial.py:67,348: (1j)**n on an integer array), any sign of n
SFA_HD c128 mul_i_pow(c128 z, int n) {
    switch (((n % 4) + 4) % 4) {
        case 0: return z;
        case 1: return mk(-z.im, z.re);
        case 2: return mk(-z.re, -z.im);
        default: return mk(z.im, -z.re);
    }
}

// ---------------------------------------------------------------------------
// Strided per-thread array view: element n of this thread lives at p[n*stride].
// On the GPU `stride` is blockDim.x and `p` points into shared memory at the
// thread's column, which makes every access conflict-free.
// ---------------------------------------------------------------------------
struct dvec {
    double* p;
    int stride;
    SFA_HD double& operator[](int n) const { return p[(long)n * stride]; }
};

constexpr double kRescaleHi = 1.0e150;
constexpr double kRescaleLo = 1.0e-150;

// Start order for Miller's algorithm.  Above the turning point n ~ x the minimal
// solution decays like exp(-n (2 d)^{3/2} / 3) with d = (n - x)/x, so reaching
// 1e-16 needs a margin of ~11.5 x^{1/3} orders beyond max(nmax, x); measured
// against mpmath: margin 68 is enough at x = 100, 100 at x = 857, 150 at x = 3000.
SFA_HD int miller_start(int nmax, double x) {
    const double top = fmax((double)nmax, x);
    return (int)top + 20 + (int)(12.0 * cbrt(top));
}

// ---------------------------------------------------------------------------
// Spherical Bessel j_0..j_nmax (x >= 0).
//   x == 0          : delta_{n0}
//   x >= nmax + 1   : forward recurrence from sin/cos (stable while n < x; this
//                     is also what SciPy does for n < x)
//   otherwise       : Miller backward recurrence with running rescale,
//                     normalised on the larger of j_0, j_1 (closed forms).
// ---------------------------------------------------------------------------
SFA_HD void sph_jn_array(int nmax, double x, dvec j) {
    if (x == 0.0) {
        j[0] = 1.0;
        for (int n = 1; n <= nmax; ++n) j[n] = 0.0;
        return;
    }
    double s, c;
#if defined(__CUDA_ARCH__)
    sincos(x, &s, &c);
#else
    s = sin(x);
    c = cos(x);
#endif
    const double ix = 1.0 / x;
    const double j0 = s * ix;
    const double j1 = (s * ix - c) * ix;
    if (x >= (double)nmax + 1.0) {
        j[0] = j0;
        if (nmax >= 1) j[1] = j1;
        double a = j0, b = j1;
        for (int n = 1; n < nmax; ++n) {
            const double t = (2 * n + 1) * ix * b - a;
            a = b;
            b = t;
            j[n + 1] = t;
        }
        return;
    }
    // backward: f_{n-1} = (2n+1)/x f_n - f_{n+1};  c = (2n+1)/x is kept as a running double
    int n = miller_start(nmax < 1 ? 1 : nmax, x);
    double fp = 0.0, f = 1.0e-30;
    double cf = (2 * n + 1) * ix;
    const double dcf = 2.0 * ix;
    // four steps per overflow check: one step grows the pair by at most (2n+1)/x, so from 1e150
    // four steps stay finite as long as x is not absurdly small (then: one check per step)
    if (x > 1e-30) {
        for (; n > nmax + 4; n -= 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double t = cf * f - fp;
                cf -= dcf;
                fp = f;
                f = t;
            }
            if (fabs(f) > kRescaleHi) {
                f *= kRescaleLo;
                fp *= kRescaleLo;
            }
        }
    }
    for (; n > nmax + 1; --n) {
        const double t = cf * f - fp;
        cf -= dcf;
        fp = f;
        f = t;
        if (fabs(f) > kRescaleHi) {
            f *= kRescaleLo;
            fp *= kRescaleLo;
        }
    }
    // now f = f_{nmax+1}, fp = f_{nmax+2}; continue storing
    // Stored values may need rescaling later: keep them relative to the running
    // scale by rescaling the already stored part when we rescale (rare: only
    // for tiny x), which keeps the code branch-light in the common case.
    for (; n >= 1; --n) {
        const double t = (2 * n + 1) * ix * f - fp;   // f_{n-1}  (exact coefficient in the stored range)
        fp = f;
        f = t;
        if (n - 1 <= nmax) j[n - 1] = f;
        if (fabs(f) > kRescaleHi) {
            f *= kRescaleLo;
            fp *= kRescaleLo;
            for (int q = n - 1; q <= nmax; ++q) j[q] *= kRescaleLo;
        }
    }
    // f == computed f_0, fp == computed f_1
    const double norm = (fabs(j0) >= fabs(j1)) ? j0 / f : j1 / fp;
    for (int q = 0; q <= nmax; ++q) j[q] *= norm;
}

// y_n forward recurrence state (dominant solution, always stable upward).
struct sph_yn_iter {
    double ym1, y, ix;   // y_{n-1}, y_n
    int n;
    SFA_HD void init(double x) {
        double s, c;
#if defined(__CUDA_ARCH__)
        sincos(x, &s, &c);
#else
        s = sin(x);
        c = cos(x);
#endif
        ix = 1.0 / x;
        y = -c * ix;                     // y_0
        ym1 = s * ix;                    // y_{-1} = j_0  (so that the recurrence yields y_1)
        n = 0;
    }
    SFA_HD void step() {                 // -> y_{n+1}
        const double t = (2 * n + 1) * ix * y - ym1;
        ym1 = y;
        y = t;
        ++n;
    }
};

// ---------------------------------------------------------------------------
// Cylinder Bessel J_0..J_nmax plus Y_0, Y_1 from one Miller pass (x > 0).
// Normalisation: J_0 + 2 sum_{k>=1} J_{2k} = 1.
// Y_0 = (2/pi)(ln(x/2)+gamma) J_0 - (4/pi) sum_{k>=1} (-1)^k J_{2k}/k          (A&S 9.1.89)
// Y_1 = -(2/(pi x)) J_0 + (2/pi)(ln(x/2)+gamma-1) J_1
//       - (2/pi) sum_{k>=1} (-1)^k (2k+1)/(k(k+1)) J_{2k+1}                      (A&S 9.1.88, n=1)
// ---------------------------------------------------------------------------
// `rk_tab` (optional): table of 1/k for k < rk_len, so the Miller loop carries no FP64 division.
SFA_HD void cyl_jn_array(int nmax, double x, dvec J, double* Y0, double* Y1, const double* rk_tab = nullptr,
                         int rk_len = 0) {
    if (x == 0.0) {
        J[0] = 1.0;
        for (int n = 1; n <= nmax; ++n) J[n] = 0.0;
        *Y0 = -INFINITY;
        *Y1 = -INFINITY;
        return;
    }
    const double ix2 = 2.0 / x;
    int n = miller_start(nmax < 1 ? 1 : nmax, x);
    if (n & 1) ++n;                      // start on an even order 2K
    // Two orders per iteration: fe = f_{2k} (even), fo = f_{2k+1} (the odd order above it).
    // The Y_1 weight (2k+1)/(k(k+1)) = 1/k + 1/(k+1), so ONE division (1/k) serves both series.
    double fe = 1.0e-30, fo = 0.0;
    double sum_even = 0.0;               // sum_{k>=1} f_{2k}
    double sum_y0 = 0.0;                 // sum (-1)^k f_{2k}/k
    double sum_y1 = 0.0;                 // sum (-1)^k (1/k + 1/(k+1)) f_{2k+1}
    double rk1 = 0.0;                    // 1/(k+1)
    for (int k = n >> 1; k >= 1; --k) {
        const double rk = (k < rk_len) ? rk_tab[k] : 1.0 / (double)k;
        const double sg = (k & 1) ? -1.0 : 1.0;
        sum_y1 += sg * fo * (rk + rk1);
        sum_even += fe;
        sum_y0 += sg * fe * rk;
        const double f_odd = (double)(2 * k) * ix2 * fe - fo;             // f_{2k-1}
        const double f_even = (double)(2 * k - 1) * ix2 * f_odd - fe;     // f_{2k-2}
        fo = f_odd;
        fe = f_even;
        rk1 = rk;
        if (2 * k - 1 <= nmax) J[2 * k - 1] = fo;
        if (2 * k - 2 <= nmax) J[2 * k - 2] = fe;
        if (fabs(fe) > kRescaleHi || fabs(fo) > kRescaleHi) {
            fe *= kRescaleLo;
            fo *= kRescaleLo;
            sum_even *= kRescaleLo;
            sum_y0 *= kRescaleLo;
            sum_y1 *= kRescaleLo;
            for (int q = 2 * k - 2; q <= nmax; ++q) J[q] *= kRescaleLo;
        }
    }
    const double f = fe, fp = fo;        // f_0, f_1
    const double norm = 1.0 / (f + 2.0 * sum_even);
    for (int q = 0; q <= nmax; ++q) J[q] *= norm;
    const double j0 = f * norm, j1 = fp * norm;
    const double two_over_pi = 0.63661977236758134308;
    const double euler_gamma = 0.57721566490153286061;
    const double lg = log(0.5 * x) + euler_gamma;
    *Y0 = two_over_pi * (lg * j0 - 2.0 * (sum_y0 * norm));
    *Y1 = two_over_pi * (-j0 / x + (lg - 1.0) * j1 - sum_y1 * norm);
}

// ---------------------------------------------------------------------------
// regularize() -- radial.py:216-266.  `alpha` for Tikh is precomputed on the
// host exactly as the reference does (a0 -> sqrt(a0/2), :255-256).
// hn is complex (ones_like(dn)) for methods 0..2 and real for 3..5; the product
// dn*hn is NumPy's full complex multiply in both cases (a real hn is upcast to
// hn+0j), which is what produces NaN for inf*0.
// ---------------------------------------------------------------------------

SFA_HD double tikh_alpha(double a0) {
    const double a = sqrt(a0 / 2.0);
    const double s = sqrt(1.0 - 1.0 / (a * a));
    return (1.0 - s) / (1.0 + s);
}

enum RegMethod { REG_NONE = 0, REG_DISCARD = 1, REG_HARDCLIP = 2, REG_SOFTCLIP = 3, REG_TIKH = 4, REG_WNG = 5 };

SFA_HD void regularize_one(c128 dn, double a0, int method, double alpha, c128* dn_out, double* hn_out);

// regularize(1/b) in one step.  In the range where |b|^2 = b.re^2 + b.im^2 can neither overflow nor
// lose bits to underflow the closed forms below need ONE division per value:
//   none : d = conj(b)/|b|^2,              h = 1
//   Tikh : d h = conj(b)/(|b|^2 + alpha^2), h = |b|^2/(|b|^2 + alpha^2)     (= 1/(1 + alpha^2 |1/b|^2))
//   wng  : d h = conj(b),                   h = |b|^2                         (= 1/|1/b|^2)
// (identical to the reference up to rounding).  Everything else -- the clipping methods, huge or
// tiny |b|, Inf/NaN -- takes NumPy's Smith division + hypot so that the reference's special-value
// behaviour is reproduced exactly.
SFA_HD void regularize_inverse(c128 b, double a0, int method, double alpha, c128* dn_out, double* hn_out) {
    const double m = fmax(fabs(b.re), fabs(b.im));
    if (m > 1e-140 && m < 1e140 && (method == REG_TIKH || method == REG_WNG || method == REG_NONE)) {
        const double m2 = b.re * b.re + b.im * b.im;
        if (method == REG_WNG) {
            *hn_out = m2;
            *dn_out = mk(b.re, -b.im);
            return;
        }
        const double inv = 1.0 / (method == REG_TIKH ? m2 + alpha * alpha : m2);
        *hn_out = (method == REG_TIKH) ? m2 * inv : 1.0;
        *dn_out = mk(b.re * inv, -b.im * inv);
        return;
    }
    regularize_one(cdiv(mk(1.0, 0.0), b), a0, method, alpha, dn_out, hn_out);
}

SFA_HD void regularize_one(c128 dn, double a0, int method, double alpha, c128* dn_out, double* hn_out) {
    const double mag = cabs(dn);
    double h = 1.0;
    switch (method) {
        case REG_NONE: break;
        case REG_DISCARD: if (mag > a0) h = 0.0; break;
        case REG_HARDCLIP: if (mag > a0) h = a0 / mag; break;
        case REG_SOFTCLIP: {
            const double pi = 3.141592653589793;
            h = 2.0 / pi * atan(pi / 2.0 * (a0 / mag));
            break;
        }
        case REG_TIKH: h = 1.0 / (1.0 + (alpha * alpha) * (mag * mag)); break;
        default: h = 1.0 / (mag * mag); break;
    }
    *hn_out = h;
    *dn_out = cmul(dn, mk(h, 0.0));
}

// ---------------------------------------------------------------------------
// One row (one wavenumber) of a radial filter.
//   family 0: sphere (weights / spherical_pw / spherical_ps), C = N+1
//   family 1: circle (circ_radial_weights / circular_pw / circular_ls), C = 2N+1
//   scale  0: b_n            1: plane wave (4 pi i^n b_n | i^n b_n)
//          2: point/line source (radial.py:109 | :388-390)
// jx / js are scratch arrays of length >= N+3 (js only for scale 2).
// out[c] receives column c of the row (reference column order).
// ---------------------------------------------------------------------------
enum Setup { SETUP_OPEN = 0, SETUP_CARD = 1, SETUP_RIGID = 2 };

template <class Out>
SFA_HD void sph_radial_row(int N, double k, double r, double rs, int setup, int scale,
                           dvec jx, dvec js, Out out) {
    const double x = k * r;
    const int top = N + 1;               // j_0' = -j_1 needs order 1 even for N = 0
    sph_jn_array(top, x, jx);
    sph_yn_iter yx;
    const bool need_y = (setup == SETUP_RIGID) && (x != 0.0);
    if (need_y) yx.init(x);
    const double ix = (x != 0.0) ? 1.0 / x : 0.0;
    const double xs = k * rs;
    sph_yn_iter ys;
    if (scale == 2) {
        sph_jn_array(top, xs, js);
        if (xs != 0.0) ys.init(xs);
    }
    const double four_pi_a = 4.0;        // the reference multiplies by 4, then by pi (:67, :109)
    const double pi = 3.141592653589793;
    for (int n = 0; n <= N; ++n) {
        const double jn = jx[n];
        c128 b;
        if (setup == SETUP_OPEN) {
            b = mk(jn, 0.0);
        } else {
            double jd;                   // scipy: j_n' = j_{n-1} - (n+1)/x j_n ; n=0: -j_1 ; x=0: 1/3 at n=1
            if (x == 0.0) jd = (n == 1) ? 1.0 / 3.0 : 0.0;
            else if (n == 0) jd = -jx[1];
            else jd = jx[n - 1] - (n + 1) * ix * jn;
            if (setup == SETUP_CARD) {
                b = mk(jn, -jd);
            } else if (x == 0.0) {
                b = mk(jn, 0.0);         // radial.py:151-153
            } else {
                const double yn = yx.y;
                double yd;
                if (n == 0) {
                    // y_0' = -y_1
                    const double y1 = (2 * 0 + 1) * yx.ix * yx.y - yx.ym1;
                    yd = -y1;
                } else {
                    yd = yx.ym1 - (n + 1) * ix * yn;
                }
                // j_n - (j_n'/h_n') h_n = -i W[j_n, y_n] / h_n' = -i / (x^2 h_n'), h = j - i y (Wronskian
                // 1/x^2): one real division instead of a complex one, and no cancellation.  Outside the
                // range where |h_n'|^2 is safe the reference's expression is evaluated literally.
                const double mag = fmax(fabs(jd), fabs(yd));
                if (mag > 1e-140 && mag < 1e140) {
                    const double inv = 1.0 / ((x * x) * (jd * jd + yd * yd));
                    b = mk(yd * inv, -jd * inv);
                } else {
                    const c128 hn = mk(jn, -yn);
                    const c128 hd = mk(jd, -yd);
                    const c128 q = cdiv(mk(jd, 0.0), hd);
                    b = csub(mk(jn, 0.0), cmul(q, hn));
                }
            }
        }
        if (need_y) yx.step();
        if (scale == 1) {
            // 4*np.pi * (1j)**n * bn  ==  ((4*pi) * i^n) * bn
            b = mul_i_pow(cscale(b, four_pi_a * pi), n);
        } else if (scale == 2) {
            if (xs == 0.0) {
                b = mk(NAN, NAN);        // y_n(0) = -inf -> NaN row (survey A9)
            } else {
                const c128 hs = mk(js[n], -ys.y);
                // bn * 4*pi * (-1j) * hn * k   (radial.py:109, left to right)
                c128 t = cscale(cscale(b, four_pi_a), pi);
                t = cmul(t, mk(0.0, -1.0));
                t = cmul(t, hs);
                b = cscale(t, k);
                ys.step();
            }
        }
        out(n, b);
    }
}

template <class Out>
SFA_HD void circ_radial_row(int N, double k, double r, double rs, int setup, int scale,
                            dvec jx, dvec js, Out out, const double* rk_tab = nullptr, int rk_len = 0) {
    const double x = k * r;
    const int top = N + 1;               // J_N' = (J_{N-1} - J_{N+1})/2
    double Y0 = 0.0, Y1 = 0.0;
    cyl_jn_array(top, x, jx, &Y0, &Y1, rk_tab, rk_len);
    const double ix2 = (x != 0.0) ? 2.0 / x : 0.0;
    // forward recurrence window for Y: (ym1, y, yp1) = (Y_{n-1}, Y_n, Y_{n+1})
    double ym1 = -Y1, y = Y0, yp1 = Y1;  // Y_{-1} = -Y_1
    const double xs = k * rs;
    double S0 = 0.0, S1 = 0.0;
    double sm1 = 0.0, s = 0.0;
    if (scale == 2) {
        cyl_jn_array(top, xs, js, &S0, &S1, rk_tab, rk_len);
        sm1 = -S1;
        s = S0;
    }
    const double ixs2 = (xs != 0.0) ? 2.0 / xs : 0.0;
    const int C = 2 * N + 1;
    for (int n = 0; n <= N; ++n) {
        const double Jn = jx[n];
        c128 b;
        if (setup == SETUP_OPEN) {
            b = mk(Jn, 0.0);
        } else {
            const double Jm1 = (n == 0) ? -jx[1] : jx[n - 1];
            const double Jd = 0.5 * (Jm1 - jx[n + 1]);      // scipy jvp: (J_{n-1} - J_{n+1})/2
            if (setup == SETUP_CARD) {
                b = mk(Jn, -Jd);
            } else if (x == 0.0) {
                b = mk(Jn, 0.0);                             // radial.py:429-431
            } else {
                const double Yd = 0.5 * (ym1 - yp1);
                // J_n - (J_n'/H_n') H_n = -i W[J_n, Y_n] / H_n' = -2i / (pi x H_n')  (Wronskian 2/(pi x))
                const double mag = fmax(fabs(Jd), fabs(Yd));
                if (mag > 1e-140 && mag < 1e140) {
                    const double inv = 0.63661977236758134308 / (x * (Jd * Jd + Yd * Yd));
                    b = mk(Yd * inv, -Jd * inv);
                } else {
                    const c128 Hn = mk(Jn, -y);
                    const c128 Hd = mk(Jd, -Yd);
                    const c128 q = cdiv(mk(Jd, 0.0), Hd);
                    b = csub(mk(Jn, 0.0), cmul(q, Hn));
                }
            }
        }
        if (setup == SETUP_RIGID && x != 0.0) {
            // advance window: Y_{n+2} = 2(n+1)/x Y_{n+1} - Y_n
            const double t = (n + 1) * ix2 * yp1 - y;
            ym1 = y;
            y = yp1;
            yp1 = t;
        }
        c128 bpos = b, bneg = (n & 1) ? mk(-b.re, -b.im) : b;   // radial.py:440 mirror (-1)^n
        if (scale == 1) {
            bpos = mul_i_pow(bpos, n);                       // 1j**n, n = [0..N, -N..-1]
            bneg = mul_i_pow(bneg, -n);
        } else if (scale == 2) {
            // bn * H_n^(2)(k rs), then * (-1j/4); H_{-n} = (-1)^n H_n
            c128 Hs = mk(js[n], -s);
            if (xs == 0.0) Hs = mk(NAN, NAN);
            const c128 Hsn = (n & 1) ? mk(-Hs.re, -Hs.im) : Hs;
            bpos = cmul(mk(-0.0, -0.25), cmul(bpos, Hs));
            bneg = cmul(mk(-0.0, -0.25), cmul(bneg, Hsn));
            const double t = (n == 0) ? S1 : n * ixs2 * s - sm1;
            sm1 = s;
            s = t;
        }
        out(n, bpos);
        if (n > 0) out(C - n, bneg);
    }
}

// ---------------------------------------------------------------------------
// Spherical harmonics (angular.py:12-70).  Fully normalised associated
// Legendre functions with Condon-Shortley phase:
//   Pb_m^m   = C_m sin^m(theta),  C_0 = sqrt(1/4pi), C_m = -C_{m-1} sqrt((2m+1)/(2m))
//   Pb_n^m   = a(n,m) (cos(theta) Pb_{n-1}^m - b(n,m) Pb_{n-2}^m),   Pb_{m-1}^m = 0
//   a(n,m)   = sqrt((4n^2-1)/(n^2-m^2)),  b(n,m) = sqrt(((n-1)^2-m^2)/(4(n-1)^2-1))
//   Y_n^m    = Pb_n^m e^{i m phi},  Y_n^{-m} = (-1)^m conj(Y_n^m)
// ---------------------------------------------------------------------------
SFA_HD double sh_a(int n, int m) {
    return sqrt((4.0 * n * n - 1.0) / ((double)n * n - (double)m * m));
}
SFA_HD double sh_b(int n, int m) {
    const double n1 = n - 1.0;
    return sqrt((n1 * n1 - (double)m * m) / (4.0 * n1 * n1 - 1.0));
}
SFA_HD double sh_seed(int m) {           // C_m
    double c = 0.28209479177387814347;   // sqrt(1/(4 pi))
    for (int i = 1; i <= m; ++i) c = -c * sqrt((2.0 * i + 1.0) / (2.0 * i));
    return c;
}
SFA_HD double ipow(double s, int m) {    // s^m by binary powering, m >= 0
    double r = 1.0, b = s;
    while (m) {
        if (m & 1) r *= b;
        b *= b;
        m >>= 1;
    }
    return r;
}
// scalar restatement used by the host-side unit test; the kernel evaluates the
// same recurrence with one lane per order m.
inline void sh_row_scalar(int N, double azi, double colat, double w, c128* out) {
    const double ct = cos(colat), st = sin(colat);
    for (int m = 0; m <= N; ++m) {
        const double cm = cos(m * azi), sm = sin(m * azi);
        double p2 = 0.0, p1 = w * sh_seed(m) * ipow(st, m);
        for (int n = m; n <= N; ++n) {
            double p;
            if (n == m) p = p1;
            else {
                p = sh_a(n, m) * (ct * p1 - sh_b(n, m) * p2);
                p2 = p1;
                p1 = p;
            }
            const int base = n * n + n;
            out[base + m] = mk(p * cm, p * sm);
            if (m > 0) {
                const double sg = (m & 1) ? -1.0 : 1.0;
                out[base - m] = mk(sg * p * cm, -sg * p * sm);
            }
        }
    }
}

}  // namespace sfa
