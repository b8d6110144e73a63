/* sfa_b200.h -- C ABI of libsfa_b200.so (sm_100a kernels for the sfa-numpy
 * `micarray.modal` hot path).
 *
 * The reference (spatialaudio/sfa-numpy) is pure Python and has no FFI of its
 * own; its boundary is the Python signatures of `micarray.modal.angular` and
 * `micarray.modal.radial`.  Each entry point below names the reference
 * function it replaces (paths relative to the reference tree).  INTEGRATION.md
 * shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *  - plain C types only; complex128 = struct {double re, im}, complex64 =
 *    struct {float re, im} (layout-compatible with NumPy / torch / C99).
 *  - pointers are DEVICE pointers unless the function name ends in `_host`.
 *  - row-major, contiguous, caller-allocated and caller-owned outputs.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *    device-pointer functions are asynchronous with respect to the host.
 *  - return value: SFA_OK (0) or a negative sfa_status; `sfa_strerror` gives
 *    the text.  Semantic errors that the reference raises as ValueError are
 *    raised by the Python host layer before any launch.
 */
#ifndef SFA_B200_H
#define SFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } sfa_c128;
typedef struct { float re, im; } sfa_c64;

enum sfa_status {
    SFA_OK = 0,
    SFA_ERR_INVALID = -1,      /* bad argument (shape, enum, null pointer)           */
    SFA_ERR_CUDA = -2,         /* a CUDA runtime/driver call failed (see last error) */
    SFA_ERR_UNSUPPORTED = -3,  /* outside the supported range (e.g. order too large) */
    SFA_ERR_WORKSPACE = -4,    /* caller workspace too small                         */
    SFA_ERR_NO_DEVICE = -5,    /* no sm_100 device                                   */
    SFA_ERR_RANGE = -6         /* radial filters beyond the float32 range of the complex64 paths (|d| > 1e30 or
                                  non-finite): regularise them or use SFA_PREC_C128_FP64 (host-buffer entry points) */
};

enum sfa_setup { SFA_OPEN = 0, SFA_CARD = 1, SFA_RIGID = 2 };       /* radial.py:146-160 */
enum sfa_method {                                                   /* radial.py:242-264 */
    SFA_REG_NONE = 0, SFA_REG_DISCARD = 1, SFA_REG_HARDCLIP = 2,
    SFA_REG_SOFTCLIP = 3, SFA_REG_TIKH = 4, SFA_REG_WNG = 5
};
enum sfa_radial_kind {
    SFA_SPH_WEIGHTS = 0,   /* radial.weights             radial.py:114-162 */
    SFA_SPH_PW = 1,        /* radial.spherical_pw        radial.py:36-67   */
    SFA_SPH_PS = 2,        /* radial.spherical_ps        radial.py:70-111  */
    SFA_CIRC_WEIGHTS = 3,  /* radial.circ_radial_weights radial.py:393-441 */
    SFA_CIRC_PW = 4,       /* radial.circular_pw         radial.py:317-348 */
    SFA_CIRC_LS = 5        /* radial.circular_ls         radial.py:351-390 */
};
enum sfa_precision {
    SFA_PREC_C64_3XTF32 = 0,  /* complex64 in/out, tcgen05 kind::tf32 with hi/lo split (3 passes), FP32 accumulate */
    SFA_PREC_C128_FP64 = 1,   /* complex128 in/out, FP64 FMA on the CUDA cores                                    */
    SFA_PREC_C64_TF32_BF16 = 2 /* complex64 in/out, hi*hi as kind::tf32 + (hi*lo + lo*hi) as one bf16 pass: 2/3 of
                                  the tensor work of 3xTF32, cross terms rounded to bf16 (~2^-19 relative)        */
};

const char* sfa_strerror(int status);
const char* sfa_last_cuda_error(void);
int sfa_version(void);
/* number of SMs / compute capability of the current device (0 on success) */
int sfa_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Library-wide tuning knobs (process-global, initialised once at load time from the SFA_<KEY> environment
 * variables; no launch path calls getenv).  Keys: "pwd_chunk_mb" (scratch budget per frequency slab),
 * "no_fused", "no_legendre", "kloop_no_hint", "fused_prepass", "fft_generic", "no_dmma", "no_kloop", "no_cluster", "no_fold", "no_tma_store", "old_build" (0/1: switch a kernel
 * variant off -- test and tuning use), "kloop_block", "issuers", "fused_D", "fused_ring" (0 = automatic), "fused_pair" (1: the CTA pairs of the fused kernel on one
 * tcgen05.mma.cta_group::2 stream -- measured slower, off by default),
 * "reserve_sms" (SMs the persistent kernels leave free, e.g. for NCCL copy kernels running under them).  Unknown key: SFA_ERR_INVALID / NaN. */
int sfa_set_option(const char* key, double value);
double sfa_get_option(const char* key);

/* Instrumentation used by bench.py: kernels launched by this library so far; optional
 * CUDA-event timing of the dominant kernel (the tcgen05 GEMM) on its launching stream. */
long long sfa_launch_count(void);
int sfa_prof_enable(int on);
int sfa_prof_read(double* total_ms, long long* count);
int sfa_prof_timeline(double* rows /* [max_rows][3] = kind, start_ms, end_ms */, int max_rows);
int sfa_prof_span(double* span_ms);   /* first launch start -> last launch end (gaps = span - total) */

/* ---- stage 1: angular matrices ------------------------------------------------ */

/* angular.sht_matrix (angular.py:12-70).  Y[q, n*n+n+m] = w[q] * Y_n^m(azi[q], colat[q]).
 * colat has Q entries, or 1 entry when colat_is_scalar != 0 (angular.py:68 broadcast).
 * weights may be NULL (ones, angular.py:62-63).  Y: [Q, (N+1)^2]. */
int sfa_sht_matrix(int N, int64_t Q, const double* azi, const double* colat, int colat_is_scalar,
                   const double* weights, sfa_c128* Y, void* stream);

/* angular.cht_matrix (angular.py:106-150).  Psi[q, i] = w[q] * exp(1j * order[i] * pol[q]),
 * order = [0, 1..N, -N..-1] (angular.py:147).  Psi: [Q, 2N+1]. */
int sfa_cht_matrix(int N, int64_t Q, const double* pol, const double* weights, sfa_c128* Psi,
                   void* stream);

/* ---- stage 2: radial filters -------------------------------------------------- */

/* Bytes of scratch `work` that sfa_radial_filter needs (zero-replacement worklist). */
size_t sfa_radial_work_bytes(int kind, int N, int64_t F);

/* One call = radial function (+ zero replacement) (+ regularised inverse), i.e. the
 * reference sequence  bn = spherical_pw(N, k, r, setup); dn, hn = regularize(1/bn, a0, method)
 * (doc/examples/modal_beamforming_rigid_spherical_array.py:34-35).
 *   k   : [F] wavenumbers (for *_WEIGHTS kinds k is kr itself and r must be 1.0)
 *   bn  : [F, C] out, C = N+1 (sphere) or 2N+1 (circle)
 *   replace_zeros != 0 applies radial.replace_zeros_of_radial_function (radial.py:165-213)
 *   method < 0: no inverse requested (dn/hn ignored, may be NULL); otherwise
 *   dn  : [F, C] out = (1/bn) * hn,  hn: [F, C] out, double for SOFTCLIP/TIKH/WNG,
 *         sfa_c128 for NONE/DISCARD/HARDCLIP (the reference's dtypes, radial.py:243-262)
 *   status: device int32[4] out: [0] number of entries with |bn| < 1e-300 (before
 *         replacement), [1] != 0 if a column had no non-zero entry (reference raises
 *         ValueError, radial.py:209), [2] != 0 if the worklist overflowed.
 *   work: device scratch of sfa_radial_work_bytes() bytes. */
int sfa_radial_filter(int kind, int N, int64_t F, const double* k, double r, double rs, int setup,
                      int replace_zeros, int method, double a0,
                      sfa_c128* bn, sfa_c128* dn, void* hn, int32_t* status, void* work,
                      void* stream);

/* radial.replace_zeros_of_radial_function on an arbitrary [F, C] matrix whose rows are
 * already ordered by ascending, distinct kr (the host layer sorts/deduplicates as
 * radial.py:196 does).  status/work as above. */
int sfa_replace_zeros(int64_t F, int C, const double* kr_sorted, sfa_c128* A, int32_t* status,
                      void* work, size_t work_bytes, void* stream);

/* radial.regularize (radial.py:216-266) on n elements. hn as in sfa_radial_filter. */
int sfa_regularize(int64_t n, const sfa_c128* dn_in, double a0, int method, sfa_c128* dn_out,
                   void* hn, void* stream);

/* radial.repeat_n_m (radial.py:294-314): [F, N+1] -> [F, (N+1)^2]. */
int sfa_repeat_n_m(int64_t F, int N, const sfa_c128* v, sfa_c128* out, void* stream);

/* radial.diagonal_mode_mat / circ_diagonal_mode_mat without the repeat step
 * (radial.py:285-290, 459-464): d [F, M] -> D [F, M, M] dense, zeros off the diagonal. */
int sfa_diagonal_stack(int64_t F, int M, const sfa_c128* d, sfa_c128* D, void* stream);

/* ---- "next" rows: remaining angular / util surface ---------------------- */

/* angular.Legendre_matrix (angular.py:73-103): L[n, q] = (2n+1)/(4 pi) P_n(ctheta[q]), [(N+1), Q]. */
int sfa_legendre_matrix(int N, int64_t Q, const double* ctheta, sfa_c128* L, void* stream);
/* util.matdiagmul (util.py:65-93): C[k, m, n] = A[m, n] * b[k, n]. */
int sfa_matdiagmul(int64_t K, int64_t M, int64_t N, const sfa_c128* A, const sfa_c128* b, sfa_c128* C,
                   void* stream);
/* util.norm_of_columns (util.py:5-21): p-norm of each column of A [M, N]. */
int sfa_norm_of_columns(int64_t M, int64_t N, const sfa_c128* A, double p, double* out, void* stream);
/* util.coherence_of_columns (util.py:24-45): max off-diagonal |Gram| of the column-normalised A.
 * norms: device scratch [N]; result: device double. */
int sfa_coherence_of_columns(int64_t M, int64_t N, const sfa_c128* A, double* norms, double* result,
                             void* stream);

/* angular.grid_equal_angle (kind 0, angular.py:153-185), grid_gauss (kind 1, angular.py:188-214),
 * grid_equal_polar_angle (kind 2, angular.py:217-235) on the device.  sfa_grid_points gives the number of
 * sampling points ((2n+2)^2, (n+1)(2n+2), 2n+1; -1 for bad arguments); azi / colat / weights are device
 * arrays of that many doubles in the reference's order (colat is not written for kind 2). */
int64_t sfa_grid_points(int kind, int n);
int sfa_grid(int kind, int n, double* azi, double* colat, double* weights, void* stream);

/* ---- stage 3: plane-wave decomposition / steering ------------------------------ */

/* q_f = Y_L * diag(d_f) * Y_Q^H * X_f   for every bin f  -- the example idiom
 *   A = matmul(matmul(Y_q, D), conj(Y_p.T)); q = matmul(A, p)
 * (doc/examples/modal_beamforming_rigid_spherical_array.py:38-39, ..._open_circular_array.py:38-39)
 * with the diagonal given as the per-column vector d_cols [F, M] (M = (N+1)^2 or 2N+1;
 * = repeat_n_m(dn) for spheres, radial.py:284) or, when d_is_per_order != 0, as dn [F, N+1]
 * which is expanded on the fly (column i -> order floor(sqrt(i))).
 *   Y_L [L, M], Y_Q [Q, M] complex128 (Y_Q already carries the quadrature weights)
 *   X   [F, Q, T], out [F, L, T]: sfa_c64 for SFA_PREC_C64_3XTF32, sfa_c128 for SFA_PREC_C128_FP64
 * `work` is device scratch of sfa_pwd_work_bytes() bytes. */
typedef struct {
    int64_t F, L, Q, T;
    int M;                 /* columns of Y_L / Y_Q                                  */
    int Cd;                /* columns of d: M, or N+1 when d_is_per_order           */
    int d_is_per_order;
    int precision;         /* enum sfa_precision                                    */
    int form;              /* 0 = auto, 1 = steering matrix A_f then A_f X_f, 2 = factored (two GEMMs) */
    int64_t out_ld;        /* row pitch of `out` in elements (0 = T, i.e. contiguous [F, L, T]).  A pitch that
                              is a multiple of 32 frames (256 B) is worth ~1.4x of HBM write bandwidth: every
                              128-byte line of a row tile then belongs to one CTA                            */
    int out_pad_writable;  /* nonzero: the caller owns the row padding of `out` (out_ld >= T rounded up to 16) and
                              the kernels may overwrite the frames [T, roundup(T, 16)) of every row with zeros, so
                              that every row ends on a whole 128-byte line (a partial line at each row end costs the
                              store stream 15-20 % of its HBM bandwidth).  0: nothing beyond frame T is touched    */
} sfa_pwd_shape;

size_t sfa_pwd_work_bytes(const sfa_pwd_shape* s);
int sfa_pwd_apply(const sfa_pwd_shape* s, const sfa_c128* Y_L, const sfa_c128* Y_Q,
                  const sfa_c128* d, const void* X, void* out, void* work, size_t work_bytes,
                  void* stream);

/* The same in two steps, for operators that stay resident while spectra stream through:
 *   sfa_pwd_prepare  everything that depends on Y_L / Y_Q only (group products G_g = sum_{m in g} Y_L[:,m] Y_Q[:,m]^H
 *                    or the split operand planes), written into `work`;
 *   sfa_pwd_run      q = ... for the current d and X, reading the prepared `work` (same shape, same work buffer).
 * `power` (nullable, complex64 precisions only): additionally writes the beam power map mean_t |q[f,l,t]|^2 as
 * float32 [F, L] (see sfa_power_map).  The fused steering kernel accumulates it in its epilogue -- a CTA owns
 * whole rows of q -- so it costs no extra pass over the beams; other paths append one sfa_power_map pass.
 * With out == NULL and a `power` pointer only the map is produced: the fused kernel skips its stores and the beams
 * never leave tensor memory (SFA_ERR_UNSUPPORTED for plans that are not the fused kernel; sfa_pwd_is_fused tells). */
int sfa_pwd_prepare(const sfa_pwd_shape* s, const sfa_c128* Y_L, const sfa_c128* Y_Q, void* work, size_t work_bytes,
                    void* stream);
int sfa_pwd_run(const sfa_pwd_shape* s, const sfa_c128* Y_L, const sfa_c128* d, const void* X, void* out,
                float* power, void* work, size_t work_bytes, void* stream);
int sfa_pwd_is_fused(const sfa_pwd_shape* s);

/* Batched complex GEMM building block used by sfa_pwd_apply, exposed for tests:
 *   out[b, n, m] = sum_k P[b?, n, k] * R[b, k, m]  (* rowscale[b, n] if given)
 * complex64 on tcgen05, 3xTF32 (hybrid = 0) or TF32+BF16 (hybrid = 1).  ldP, ldR, ldo in elements;
 * strideP = 0 shares P across the batch. */
int sfa_cgemm3x(int64_t B, int64_t Nn, int64_t Mm, int64_t K,
                const sfa_c64* P, int64_t ldP, int64_t strideP,
                const sfa_c64* R, int64_t ldR, int64_t strideR,
                sfa_c64* out, int64_t ldo, int64_t strideO,
                const sfa_c64* rowscale, int64_t strideS, int hybrid, void* stream);

/* Host-buffer plane-wave decomposition: the call a reference user makes with NumPy
 * arrays.  Creates device buffers/streams once; `run` copies X (pinned or pageable
 * host memory) to the device in frequency chunks, computes, and copies the beams back,
 * overlapping the three on separate streams. */
typedef struct sfa_pwd_host_ctx sfa_pwd_host_ctx;
int sfa_pwd_host_create(sfa_pwd_host_ctx** ctx, const sfa_pwd_shape* s, int device,
                        int64_t chunk_bins);
int sfa_pwd_host_set_operators(sfa_pwd_host_ctx* ctx, const sfa_c128* Y_L_host,
                               const sfa_c128* Y_Q_host, const sfa_c128* d_host);
int sfa_pwd_host_run(sfa_pwd_host_ctx* ctx, const void* X_host, void* out_host);
/* Operators that already live on the device (the outputs of sfa_sht_matrix / sfa_radial_filter on `stream`):
 * device-to-device copies, no round trip through the host. */
int sfa_pwd_host_set_operators_dev(sfa_pwd_host_ctx* ctx, const sfa_c128* Y_L_dev, const sfa_c128* Y_Q_dev,
                                   const sfa_c128* d_dev, void* stream);
/* As sfa_pwd_host_run, with the beam power map mean_t |q|^2 ([F, L] float32, host) as a second -- or, when
 * out_host is NULL, the only -- result: the "beam map" reading of the metric moves 4 F L bytes back over PCIe
 * instead of 8 F L T. */
int sfa_pwd_host_run_power(sfa_pwd_host_ctx* ctx, const void* X_host, void* out_host, float* power_host);
int sfa_pwd_host_destroy(sfa_pwd_host_ctx* ctx);
/* pinned host memory helpers for the host path */
int sfa_host_alloc(void** p, size_t bytes);
int sfa_host_free(void* p);

/* Multi-GPU exchange of the beam tensor ("only an all-gather of the output beam maps"): DMA copy of
 * `bytes` from `src` (device src_dev) into `dst` (device dst_dev; typically a peer rank's output buffer mapped
 * over CUDA IPC), asynchronous on `stream` (a stream of src_dev).  Runs on the copy engines over
 * NVLink / NVSwitch, so it overlaps the persistent tensor-core kernels that own every SM.  Enables peer access on
 * first use; SFA_ERR_UNSUPPORTED when the two devices have no peer path. */
int sfa_peer_copy(void* dst, int dst_dev, const void* src, int src_dev, size_t bytes, void* stream);

/* Beam power map mean_t |q[f,l,t]|^2 (what the reference's examples plot as db(q), e.g.
 * doc/examples/modal_beamforming_rigid_spherical_array.py:43-57, and the all-gathered quantity of the
 * multi-GPU path).  q: rows = F*L rows of T complex64 with row pitch ld (elements); out: [rows] float32. */
int sfa_power_map(int64_t rows, int64_t T, int64_t ld, const sfa_c64* q, float* out, void* stream);

/* ---- time <-> frequency ends of the path -------------------------------- */

/* np.fft.irfft(q, n, axis=0), optionally followed by np.fft.fftshift(., axes=0): the last step of every
 * example script, q_t = fftshift(irfft(q, axis=0), axes=0)
 * (doc/examples/modal_beamforming_rigid_spherical_array.py:40, ..._open_circular_array.py:40).
 *   q   [F, C] complex (C = product of the trailing axes, e.g. L or L*T; row pitch ld elements)
 *   n   output samples; 0 = NumPy's default 2 (F - 1).  Bins beyond n/2 are ignored, missing bins are zero,
 *       the imaginary parts of the DC (and, for even n, Nyquist) bins are ignored -- as np.fft.irfft does.
 *   out [n, C] real (row pitch ldo): float32 for complex64 input (is_double = 0), float64 for complex128.
 * Hand-written shared-memory FFT (radix-4 Stockham, two real columns per complex transform): powers of two
 * n <= 8192 (float) / 4096 (double) directly, other lengths through Bluestein's chirp-z identity on the next
 * power of two >= 2n - 1 (so n <= 4096 / 2048); beyond that SFA_ERR_UNSUPPORTED.
 * `work`: device scratch of sfa_irfft_work_bytes(n, is_double) bytes (twiddle / chirp tables). */
size_t sfa_irfft_work_bytes(int64_t n, int is_double);
int sfa_irfft_axis0(int64_t F, int64_t C, int64_t n, const void* q, int64_t ld, void* out, int64_t ldo,
                    int is_double, int shift, void* work, size_t work_bytes, void* stream);

/* STFT front end producing the spectra directly in the steering operator's layout:
 *   X[f, q, t] = sum_j window[j] x[q, t*hop + j] exp(-2 pi i f j / nfft),  f < F_keep (0 = nfft/2 + 1),
 *   t < T = 1 + (S - nfft) / hop; x [Q, S] real, window [nfft] real or NULL (rectangular), X [F_keep, Q, T].
 * nfft: power of two <= 16384.  float32/complex64 (is_double = 0) or float64/complex128. */
size_t sfa_stft_work_bytes(int nfft, int is_double);
int sfa_stft(int64_t Q, int64_t S, int nfft, int hop, int64_t F_keep, const void* x, const void* window,
             void* X, int is_double, void* work, size_t work_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SFA_B200_H */
