// Stage 2: radial mode strengths b_n(kr), zero replacement, regularised inverse.
//
//   sfa_radial_filter  <- radial.weights / spherical_pw / spherical_ps            radial.py:36-162
//                         radial.circ_radial_weights / circular_pw / circular_ls  radial.py:317-441
//                         + replace_zeros_of_radial_function (decorator)          radial.py:9-33,165-213
//                         + regularize(1/bn, a0, method)                          radial.py:216-266
//   sfa_regularize, sfa_replace_zeros, sfa_repeat_n_m, sfa_diagonal_stack         radial.py:216-314,444-465
//
// radial_filter_kernel: one thread per wavenumber runs the FP64 recurrences of
// sfa_math.cuh (Miller backward for j_n/J_n, forward for y_n/Y_n) with its
// scratch arrays strided through shared memory (conflict-free), writes its row
// into a shared tile, and the CTA then streams the tile out with coalesced
// 16-byte stores -- b_n, and in the same pass d_n = (1/b_n) h_n and h_n, so the
// inverse never makes a separate trip through HBM.  Algorithmic traffic: 8 B in
// per bin, 16 B per b_n out (+16 +8/16 B per (d_n, h_n)).
#include "sfa_common.cuh"
#include "sfa_math.cuh"

namespace sfa {

constexpr double kZeroThreshold = 1e-300;    // radial.py:198
constexpr int kRadialThreads = 128;
constexpr int kRkTable = 256;
constexpr long long kWorklistCap = 1 << 20;  // zero entries that can be repaired per call

struct TileOut {
    double2* row;
    int ncols;      // columns kept in the tile (circle: only the orders n >= 0, the mirror is rebuilt on the way out)
    __device__ __forceinline__ void operator()(int c, c128 v) const {
        if (c < ncols) row[c] = make_double2(v.re, v.im);
    }
};

__device__ __forceinline__ bool is_zero_entry(double2 v) { return cabs_less(mk(v.x, v.y), kZeroThreshold); }

// status[0] zero count, [1] unrepairable column, [2] worklist overflow, [3] worklist length
// FAMILY / SETUP / SCALE are compile-time so the per-order loop carries no mode branches.
template <int FAMILY, int SETUP, int SCALE>
__global__ void __launch_bounds__(kRadialThreads)
radial_filter_kernel(int N, long long F, const double* __restrict__ k, double r, double rs,
                     int want_zero_list, int method, double a0, double alpha,
                     double2* __restrict__ bn, double2* __restrict__ dn, double* __restrict__ hn_real,
                     double2* __restrict__ hn_cplx, int* __restrict__ status,
                     long long* __restrict__ worklist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = FAMILY ? 2 * N + 1 : N + 1;
    // Circle: column C-n is (-1)^n column n for the bare weights and EQUAL to column n once the i^n /
    // H_n factors are applied ((-1)^n i^-n = i^n, H_-n = (-1)^n H_n; radial.py:345-348, 388-390, 440), so
    // the tile keeps the N+1 non-negative orders only: half the shared memory per wavenumber (twice the
    // resident warps for these latency-bound FP64 recurrences) and one regularised inverse per pair.
    const int Ct = N + 1;                                   // tile columns
    const int Cp = Ct | 1;                                  // odd row pitch: conflict-free 16 B rows
    const int T = blockDim.x;
    double2* tile = reinterpret_cast<double2*>(smem_raw);   // [T][Cp]
    double* scratch = reinterpret_cast<double*>(tile + (size_t)T * Cp);   // 1 or 2 arrays [(N+3)][T]
    const int tid = threadIdx.x;
    __shared__ double rk_tab[kRkTable];                     // 1/k for the cylinder Miller loop
    if (FAMILY == 1) {
        for (int i = tid; i < kRkTable; i += T) rk_tab[i] = i ? 1.0 / (double)i : 0.0;
        __syncthreads();
    }
    const long long nblk = (F + T - 1) / T;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long f0 = blk * T;
        const long long f = f0 + tid;
        if (f < F) {
            dvec jx{scratch + tid, T};
            dvec js{scratch + (size_t)(N + 3) * T + tid, T};
            TileOut out{tile + (size_t)tid * Cp, Ct};
            if (FAMILY == 0) sph_radial_row(N, k[f], r, rs, SETUP, SCALE, jx, js, out);
            else circ_radial_row(N, k[f], r, rs, SETUP, SCALE, jx, js, out, rk_tab, kRkTable);
        }
        __syncthreads();
        const int rows = (int)((F - f0 < T) ? (F - f0) : T);
        const int total = rows * Ct;
        int local_zero = 0;
        // (row, col) of tile element e = tid, tid + T, ... tracked incrementally (no division per element)
        int rr = tid / Ct, cc = tid - rr * Ct;
        const int drr = T / Ct, dcc = T - drr * Ct;
        for (int e = tid; e < total; e += T, rr += drr, cc += dcc) {
            if (cc >= Ct) {
                cc -= Ct;
                ++rr;
            }
            const double2 b = tile[(size_t)rr * Cp + cc];
            const long long g = (f0 + rr) * C + cc;
            const bool mirror = FAMILY == 1 && cc > 0;          // also emit column C - n
            const long long gm = (f0 + rr) * C + (C - cc);
            const double sgn = (SCALE == 0 && (cc & 1)) ? -1.0 : 1.0;
            bn[g] = b;
            if (mirror) bn[gm] = make_double2(sgn * b.x, sgn * b.y);
            if (is_zero_entry(b)) {
                local_zero += mirror ? 2 : 1;
                if (want_zero_list) {
                    // every slot below the cap is written (a mirrored pair may straddle it); find/apply
                    // do nothing once the overflow flag is up, so no unwritten slot is ever read
                    const int slot = atomicAdd(&status[3], mirror ? 2 : 1);
                    if (slot < kWorklistCap) worklist[slot] = g;
                    if (mirror && slot + 1 < kWorklistCap) worklist[slot + 1] = gm;
                    if (slot + (mirror ? 1 : 0) >= kWorklistCap) status[2] = 1;
                }
            }
            if (method >= 0) {
                c128 d;
                double h;
                regularize_inverse(mk(b.x, b.y), a0, method, alpha, &d, &h);
                dn[g] = make_double2(d.re, d.im);
                if (hn_real) hn_real[g] = h;
                else hn_cplx[g] = make_double2(h, 0.0);
                if (mirror) {                                    // 1/(-b) = -(1/b), same |.| and h
                    dn[gm] = make_double2(sgn * d.re, sgn * d.im);
                    if (hn_real) hn_real[gm] = h;
                    else hn_cplx[gm] = make_double2(h, 0.0);
                }
            }
        }
        if (local_zero) atomicAdd(&status[0], local_zero);
        __syncthreads();
    }
}

// Phase 1 of the repair: for every listed zero entry find the row of the nearest
// kr whose entry in the same column is non-zero (ties -> smaller kr, which is what
// argmin over the kr-sorted rows gives, radial.py:200-206).  One warp per entry.
// `order` maps sorted position -> row (NULL: rows are already sorted by kr).
__global__ void __launch_bounds__(256)
find_replacement_kernel(long long F, int C, const double* __restrict__ kr,
                        const double2* __restrict__ A, int* __restrict__ status,
                        long long* __restrict__ worklist) {
    // One CTA per listed entry (the normal case is a handful of entries: the k = 0 row): 256 threads
    // scan the F rows of the column, then a warp-shuffle + shared-memory argmin.
    __shared__ double s_best[8], s_bestk[8];
    __shared__ long long s_besti[8];
    if (status[2]) return;                                  // worklist overflow: reported by the host layer
    const int nlist = min((long long)status[3], kWorklistCap);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (long long e = blockIdx.x; e < nlist; e += gridDim.x) {
        const long long g = worklist[e];
        const long long i = g / C;
        const int j = (int)(g - i * C);
        const double ki = kr[i];
        double best = INFINITY, bestk = INFINITY;
        long long besti = -1;
        for (long long l = threadIdx.x; l < F; l += blockDim.x) {
            if (is_zero_entry(A[l * C + j])) continue;
            const double kl = kr[l];
            const double dist = fabs(kl - ki);
            if (dist < best || (dist == best && (kl < bestk || (kl == bestk && l < besti)))) {
                best = dist;
                bestk = kl;
                besti = l;
            }
        }
        auto merge = [&](double ob, double ok, long long oi) {
            const bool take = (oi >= 0) && (besti < 0 || ob < best ||
                              (ob == best && (ok < bestk || (ok == bestk && oi < besti))));
            if (take) {
                best = ob;
                bestk = ok;
                besti = oi;
            }
        };
        for (int off = 16; off; off >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, off);
            const double ok = __shfl_xor_sync(0xffffffffu, bestk, off);
            const long long oi = __shfl_xor_sync(0xffffffffu, besti, off);
            merge(ob, ok, oi);
        }
        if (lane == 0) {
            s_best[wib] = best;
            s_bestk[wib] = bestk;
            s_besti[wib] = besti;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) merge(s_best[w], s_bestk[w], s_besti[w]);
            if (besti < 0) status[1] = 1;
            worklist[kWorklistCap + e] = besti;
        }
        __syncthreads();
    }
}

// Phase 2: copy the replacement values and redo the inverse at those entries.
__global__ void apply_replacement_kernel(int C, int method, double a0, double alpha,
                                         double2* __restrict__ A, double2* __restrict__ dn,
                                         double* __restrict__ hn_real, double2* __restrict__ hn_cplx,
                                         const int* __restrict__ status,
                                         const long long* __restrict__ worklist) {
    const int nlist = min((long long)status[3], kWorklistCap);
    if (status[1] || status[2]) return;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nlist;
         e += (long long)gridDim.x * blockDim.x) {
        const long long g = worklist[e];
        const long long src = worklist[kWorklistCap + e];
        if (src < 0) continue;
        const int j = (int)(g % C);
        const double2 b = A[src * C + j];
        A[g] = b;
        if (method >= 0) {
            c128 d;
            double h;
            regularize_one(cdiv(mk(1.0, 0.0), mk(b.x, b.y)), a0, method, alpha, &d, &h);
            dn[g] = make_double2(d.re, d.im);
            if (hn_real) hn_real[g] = h;
            else hn_cplx[g] = make_double2(h, 0.0);
        }
    }
}

// Lists the zero entries of an existing matrix (standalone replace_zeros path).
__global__ void list_zeros_kernel(long long total, const double2* __restrict__ A,
                                  int* __restrict__ status, long long* __restrict__ worklist) {
    int local = 0;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        if (is_zero_entry(A[g])) {
            ++local;
            const int slot = atomicAdd(&status[3], 1);
            if (slot < kWorklistCap) worklist[slot] = g;
            else status[2] = 1;
        }
    }
    if (local) atomicAdd(&status[0], local);
}

__global__ void regularize_kernel(long long n, const double2* __restrict__ in, double a0, int method,
                                  double alpha, double2* __restrict__ out, double* __restrict__ hn_real,
                                  double2* __restrict__ hn_cplx) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double2 v = in[i];
        c128 d;
        double h;
        regularize_one(mk(v.x, v.y), a0, method, alpha, &d, &h);
        out[i] = make_double2(d.re, d.im);
        if (hn_real) hn_real[i] = h;
        else if (hn_cplx) hn_cplx[i] = make_double2(h, 0.0);
    }
}

__global__ void repeat_n_m_kernel(long long F, int N, const double2* __restrict__ v,
                                  double2* __restrict__ out) {
    const int M = (N + 1) * (N + 1);
    const long long total = F * M;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long f = g / M;
        const int i = (int)(g - f * M);
        int n = (int)sqrtf((float)i);
        while ((n + 1) * (n + 1) <= i) ++n;
        while (n * n > i) --n;
        out[g] = v[f * (N + 1) + n];
    }
}

__global__ void diagonal_stack_kernel(long long F, int M, const double2* __restrict__ d,
                                      double2* __restrict__ D) {
    const long long per = (long long)M * M;
    const long long total = F * per;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long f = g / per;
        const long long rem = g - f * per;
        const int i = (int)(rem / M), j = (int)(rem - (long long)i * M);
        D[g] = (i == j) ? d[f * M + i] : make_double2(0.0, 0.0);
    }
}

static unsigned grid_for(long long items, int threads) {
    long long blocks = ceil_div(items, threads);
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

static void split_hn(int method, void* hn, double** hr, double2** hc) {
    *hr = nullptr;
    *hc = nullptr;
    if (method < 0 || !hn) return;
    if (method <= REG_HARDCLIP) *hc = reinterpret_cast<double2*>(hn);
    else *hr = reinterpret_cast<double*>(hn);
}

}  // namespace sfa

using namespace sfa;

extern "C" size_t sfa_radial_work_bytes(int, int, int64_t) {
    return (size_t)(2 * kWorklistCap) * sizeof(long long);
}

extern "C" int sfa_radial_filter(int kind, int N, int64_t F, const double* k, double r, double rs,
                                 int setup, int replace_zeros, int method, double a0,
                                 sfa_c128* bn, sfa_c128* dn, void* hn, int32_t* status, void* work,
                                 void* stream) {
    if (kind < 0 || kind > SFA_CIRC_LS || N < 0 || F < 0 || setup < 0 || setup > SFA_RIGID ||
        method > SFA_REG_WNG || !status)
        return SFA_ERR_INVALID;
    if (F > 0 && (!k || !bn)) return SFA_ERR_INVALID;
    if (method >= 0 && (!dn || !hn)) return SFA_ERR_INVALID;
    if (replace_zeros && !work) return SFA_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    SFA_CUDA_CHECK(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st));
    if (F == 0) return SFA_OK;
    const int family = kind >= SFA_CIRC_WEIGHTS;
    const int scale = family ? kind - SFA_CIRC_WEIGHTS : kind;
    const int C = family ? 2 * N + 1 : N + 1;
    const int Cp = (N + 1) | 1;                 // tile columns: the non-negative orders (see the kernel)
    (void)C;
    // rows per CTA: as many as fit in shared memory (large orders, e.g. the examples' N = 50 circle
    // with 101 columns, run with narrower CTAs)
    int T = kRadialThreads;
    size_t smem = 0;
    for (;; T >>= 1) {
        smem = (size_t)T * Cp * sizeof(double2) + (size_t)(scale == 2 ? 2 : 1) * (N + 3) * T * sizeof(double);
        if (smem <= 200 * 1024 || T == 32) break;
    }
    if (smem > 220 * 1024) return SFA_ERR_UNSUPPORTED;
    const double alpha = (method == SFA_REG_TIKH) ? tikh_alpha(a0) : 0.0;
    double* hr;
    double2* hc;
    split_hn(method, hn, &hr, &hc);
    long long blocks = ceil_div(F, T);
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    long long* wl = reinterpret_cast<long long*>(work);
    int rc = SFA_ERR_INVALID;
#define SFA_RADIAL_CASE(FAM, SET, SCL)                                                                         \
    if (family == FAM && setup == SET && scale == SCL) {                                                       \
        auto kern = radial_filter_kernel<FAM, SET, SCL>;                                                       \
        SFA_ENSURE_DYN_SMEM(kern, smem);                                                                       \
        kern<<<(unsigned)blocks, T, smem, st>>>(N, F, k, r, rs, replace_zeros ? 1 : 0, method, a0, alpha,      \
                                                 reinterpret_cast<double2*>(bn), reinterpret_cast<double2*>(dn), \
                                                 hr, hc, status, wl);                                          \
        rc = SFA_OK;                                                                                           \
    }
#define SFA_RADIAL_SCALES(FAM, SET) SFA_RADIAL_CASE(FAM, SET, 0) SFA_RADIAL_CASE(FAM, SET, 1) SFA_RADIAL_CASE(FAM, SET, 2)
#define SFA_RADIAL_SETUPS(FAM) SFA_RADIAL_SCALES(FAM, 0) SFA_RADIAL_SCALES(FAM, 1) SFA_RADIAL_SCALES(FAM, 2)
    SFA_RADIAL_SETUPS(0)
    SFA_RADIAL_SETUPS(1)
#undef SFA_RADIAL_SETUPS
#undef SFA_RADIAL_SCALES
#undef SFA_RADIAL_CASE
    if (rc != SFA_OK) return rc;
    SFA_LAUNCH_CHECK();
    if (replace_zeros) {
        // Both kernels return at once when the zero list is empty (the normal case).
        find_replacement_kernel<<<64, 256, 0, st>>>(F, C, k, reinterpret_cast<double2*>(bn), status, wl);
        SFA_LAUNCH_CHECK();
        apply_replacement_kernel<<<32, 256, 0, st>>>(C, method, a0, alpha, reinterpret_cast<double2*>(bn),
                                                      reinterpret_cast<double2*>(dn), hr, hc, status, wl);
        SFA_LAUNCH_CHECK();
    }
    return SFA_OK;
}

extern "C" int sfa_replace_zeros(int64_t F, int C, const double* kr_sorted, sfa_c128* A,
                                 int32_t* status, void* work, size_t work_bytes, void* stream) {
    if (F < 0 || C < 0 || !status || !work) return SFA_ERR_INVALID;
    if (work_bytes < sfa_radial_work_bytes(0, 0, F)) return SFA_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    SFA_CUDA_CHECK(cudaMemsetAsync(status, 0, 4 * sizeof(int32_t), st));
    if (F == 0 || C == 0) return SFA_OK;
    if (!kr_sorted || !A) return SFA_ERR_INVALID;
    long long* wl = reinterpret_cast<long long*>(work);
    double2* a = reinterpret_cast<double2*>(A);
    list_zeros_kernel<<<grid_for((long long)F * C, 256), 256, 0, st>>>((long long)F * C, a, status, wl);
    SFA_LAUNCH_CHECK();
    find_replacement_kernel<<<64, 256, 0, st>>>(F, C, kr_sorted, a, status, wl);
    SFA_LAUNCH_CHECK();
    apply_replacement_kernel<<<32, 256, 0, st>>>(C, -1, 0.0, 0.0, a, nullptr, nullptr, nullptr, status, wl);
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

extern "C" int sfa_regularize(int64_t n, const sfa_c128* dn_in, double a0, int method,
                              sfa_c128* dn_out, void* hn, void* stream) {
    if (n < 0 || method < 0 || method > SFA_REG_WNG) return SFA_ERR_INVALID;
    if (n == 0) return SFA_OK;
    if (!dn_in || !dn_out) return SFA_ERR_INVALID;
    double* hr;
    double2* hc;
    split_hn(method, hn, &hr, &hc);
    const double alpha = (method == SFA_REG_TIKH) ? tikh_alpha(a0) : 0.0;
    regularize_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        n, reinterpret_cast<const double2*>(dn_in), a0, method, alpha,
        reinterpret_cast<double2*>(dn_out), hr, hc);
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

extern "C" int sfa_repeat_n_m(int64_t F, int N, const sfa_c128* v, sfa_c128* out, void* stream) {
    if (F < 0 || N < 0) return SFA_ERR_INVALID;
    if (F == 0) return SFA_OK;
    if (!v || !out) return SFA_ERR_INVALID;
    const long long total = (long long)F * (N + 1) * (N + 1);
    repeat_n_m_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        F, N, reinterpret_cast<const double2*>(v), reinterpret_cast<double2*>(out));
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

extern "C" int sfa_diagonal_stack(int64_t F, int M, const sfa_c128* d, sfa_c128* D, void* stream) {
    if (F < 0 || M < 0) return SFA_ERR_INVALID;
    if (F == 0 || M == 0) return SFA_OK;
    if (!d || !D) return SFA_ERR_INVALID;
    const long long total = (long long)F * M * M;
    diagonal_stack_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        F, M, reinterpret_cast<const double2*>(d), reinterpret_cast<double2*>(D));
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}
