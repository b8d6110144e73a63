// Batched complex GEMM on the 5th-generation tensor cores (sm_100a):
//
//     out[b, n, m] = sum_k P[b|shared, n, k] * R[b, k, m]      (* rowscale[b, n])
//
// complex64 data, "3xTF32": every real operand x is split as x = hi + lo with
// hi = tf32(x) (cvt.rna) and lo = x - hi; the product is accumulated in FP32 as
// hi*hi + hi*lo + lo*hi (the dropped lo*lo term is ~2^-22 relative).  A complex
// product is 4 real products into two accumulators (Re, Im); the minus sign of
// Re -= Ri*Pi uses the instruction descriptor's negate-A bit.
//
// Second mode ("hybrid", kHybrid): hi*hi runs as one kind::tf32 pass and the two cross terms
// hi*lo + lo*hi run as ONE kind::f16 (bf16) pass over a K extent interleaved as
// (hi_bf16(k), lo_bf16(k)) on the A side and (lo_bf16(k), hi_bf16(k)) on the B side, packed two
// bf16 per 32-bit slot -- same addresses, same 8 slots per k-step, 2/3 of the tensor work.
// The cross terms are ~2^-11 of the result, so their bf16 rounding (2^-9) costs ~2^-19 relative.
//
// This is stage 3 of the hot path: the plane-wave-decomposition/steering
// contraction of the reference's example idiom
//     A = matmul(matmul(Y_q, D), conj(Y_p.T)); q = matmul(A, p)
// (doc/examples/modal_beamforming_rigid_spherical_array.py:38-39), for which
// pwd_apply.cu issues one or two calls of this kernel per frequency chunk.
//
// Mapping onto the machine (one persistent CTA per SM, 448 threads):
//   * UMMA M (128 TMEM lanes) = 128 consecutive m (time frames) of one batch b.
//     The R^T tile [128 m x K<=64] is the A operand and lives in TENSOR MEMORY
//     (TS-form tcgen05.mma): four "loader" warps read R straight from global
//     memory (coalesced along m), split hi/lo in registers and write the four
//     planes (re_hi, im_hi, re_lo, im_lo; 64 columns each) with tcgen05.st.
//     It stays there for the whole task (b, m-tile) -> R is read exactly once.
//   * UMMA N = 64 rows n of P per step.  P is always produced by this library
//     (steering matrices, SH matrices), so it is stored pre-split and planar
//     ([4 planes][n][k] fp32).  One elected thread streams it with TMA
//     (cp.async.bulk.tensor, SWIZZLE_128B) into a 2-stage shared-memory ring in
//     exactly the canonical K-major layout the UMMA descriptor reads; the two CTAs of a
//     cluster (m-tiles 2j, 2j+1 of one batch) fetch half of the boxes each and multicast them.
//   * One elected thread issues the tcgen05.mma.kind::tf32 stream:
//     3 passes x 4 real products x K/8 k-steps per n-tile, accumulating into a
//     double-buffered pair of TMEM accumulators (Re, Im: 64 columns each).
//   * Eight epilogue warps read the accumulators with tcgen05.ld (lane = m, 32 columns
//     per warp), release the accumulator at once, stage 16-row pieces of the tile in a
//     2-deep ring of 16 KB shared-memory buffers per column half and write them with TMA
//     stores (reduce-add for split-K passes).  Rows that are not 16-byte aligned (odd ldo) and
//     folded batches store complex64 directly from registers: a warp writes 32 consecutive m of
//     one row n = 256 contiguous bytes per store instruction.
// TMEM budget: 2 x (64 Re + 64 Im) accumulator columns + 256 operand columns = 512.
// SMEM: 2 x 64 KB operand stages + 2 x 2 x 16 KB store staging.
#include <cuda.h>
#include <stdlib.h>
#include "sfa_common.cuh"
#include "sfa_tcgen05.cuh"

namespace sfa {

constexpr int kTileM = 128;          // m per task (TMEM lanes)
constexpr int kTileN = 64;           // rows of P per MMA step
constexpr int kMaxK = 64;            // complex k resident in TMEM per launch
#ifndef SFA_GEMM_STAGES
#define SFA_GEMM_STAGES 2
#endif
#ifndef SFA_GEMM_STORE_RING
#define SFA_GEMM_STORE_RING 2
#endif
constexpr int kStagesB = SFA_GEMM_STAGES;
constexpr int kStoreRing = SFA_GEMM_STORE_RING;          // staging buffers per epilogue half
constexpr int kBoxK = 32;            // fp32 per 128-byte swizzle row
constexpr int kBoxBytes = kTileN * kBoxK * 4;            // 8 KB
constexpr int kStageBytes = 4 * 2 * kBoxBytes;           // 4 planes x 2 k-boxes = 64 KB
constexpr int kThreads = 480;          // TMA, MMA, 4 loader warps, 8 epilogue warps (+ optional 2nd MMA issuer)
constexpr int kImIssuerWarp = 14;
constexpr int kTmemCols = 512;
constexpr int kAccCols = 2 * kTileN;                     // Re + Im per accumulator stage
constexpr int kOperandCol0 = 2 * kAccCols;               // 256
constexpr int kStoreRows = 16;                           // rows (n) per TMA store box
constexpr int kStoreBytes = kStoreRows * kTileM * 8;     // 16 KB staging buffer per epilogue half

enum OutMode { OUT_C64 = 0, OUT_C64_FIRST_OF_MANY = 1 /* plain store, first split-K pass */, OUT_C64_ACCUM = 2 };

struct GemmParams {
    long long B, Nn, Mm;
    int K;                  // <= kMaxK for this launch
    const float2* R;        // [B][K][ldR]
    long long ldR, strideR;
    int p_shared;           // P has no batch dimension
    float* out;             // float2 view
    long long ldo, strideO; // in float2 elements
    const float2* rowscale; // [B][Nn] or null
    long long strideS;
    int out_mode;
    int m_tiles;            // ceil(Mm / 128)
    long long tasks;        // B * m_tiles
    int n_tiles;            // ceil(Nn / 64)
    int debug_no_tma;
    int fold;               // P shared and Mm <= 64: m-tiles run over the folded (batch, m) axis
    int issuers;            // MMA issuer warps per CTA (1 or 2)
    int csize;              // CTAs per cluster sharing the P stream by TMA multicast (1 or 2)
    int m_groups;           // ceil(m_tiles / csize)
    int tma_store;          // epilogue stages tiles in smem and writes them with TMA (needs even ldo)
};

// (PTX helpers and descriptors: sfa_tcgen05.cuh)

struct __align__(8) Barriers {
    uint64_t b_full[kStagesB];
    uint64_t b_empty[kStagesB];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint64_t a_full;
    uint64_t a_empty;
    uint32_t tmem_base;
    uint32_t pad;
};

template <bool kHybrid>
__global__ void __launch_bounds__(kThreads, 1)
cgemm3x_kernel(const __grid_constant__ CUtensorMap p_map, const __grid_constant__ CUtensorMap o_map,
               const GemmParams prm) {
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B tiles need 1024-byte alignment; the launch reserves 1 KB of slack.
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;                                    // kStagesB x 64 KB
    unsigned char* store_base = smem + (size_t)kStagesB * kStageBytes;   // 2 x 16 KB epilogue staging
    Barriers* bars = reinterpret_cast<Barriers*>(store_base + 2 * kStoreRing * kStoreBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // Clusters of `csize` CTAs walk the (batch, m-group) list together: CTA r of a cluster owns
    // m-tile  mg * csize + r  (a "ghost" tile beyond m_tiles computes zeros and stores nothing).
    const int csize = prm.csize;
    const uint32_t crank = (csize > 1) ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
    const long long cluster_id = blockIdx.x / csize, n_clusters = gridDim.x / csize;
    const long long n_groups = prm.fold ? prm.m_groups : prm.B * prm.m_groups;
    const int ksteps = (prm.K + 7) >> 3;               // MMA k-steps of 8
    const int kboxes = (prm.K + kBoxK - 1) / kBoxK;    // TMA boxes along k (1 or 2)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesB; ++s) {
            mbar_init(&bars->b_full[s], 1);
            mbar_init(&bars->b_empty[s], prm.issuers * csize);   // every MMA issuer of every CTA in the cluster
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->acc_full[a], prm.issuers);
            mbar_init(&bars->acc_empty[a], 8);
        }
        mbar_init(&bars->a_full, 4);
        mbar_init(&bars->a_empty, prm.issuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&bars->tmem_base)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();                 // peers' barriers are initialised before any remote arrive
    tcgen05_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer: stream P tiles =====================
        if (elect_one_sync()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&p_map) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (long long grp = cluster_id; grp < n_groups; grp += n_clusters) {
                const long long b = grp / prm.m_groups;
                const int pb = prm.p_shared ? 0 : (int)b;
                for (int nt = 0; nt < prm.n_tiles; ++nt) {
                    mbar_wait(&bars->b_empty[stage], phase ^ 1);
                    unsigned char* dst = stage_base + (size_t)stage * kStageBytes;
                    if (prm.debug_no_tma) {              // DEBUG: operand tiles not fetched (timing experiments only)
                        mbar_arrive(&bars->b_full[stage]);
                        if (++stage == kStagesB) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    mbar_arrive_expect_tx(&bars->b_full[stage], (uint32_t)(4 * kboxes * kBoxBytes));
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl) {
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
                            if (kb >= kboxes) continue;
                            unsigned char* d = dst + (pl * 2 + kb) * kBoxBytes;
                            if (csize == 1) {
                                tma_load_3d(&p_map, &bars->b_full[stage], d, kb * kBoxK, nt * kTileN, pb * 4 + pl);
                            } else if (((kboxes > 1 ? pl * 2 + kb : pl) & 1) == (int)crank) {
                                // each CTA of the pair fetches half of the boxes and multicasts them to both
                                tma_load_3d_mc(&p_map, &bars->b_full[stage], d, kb * kBoxK, nt * kTileN, pb * 4 + pl, cmask);
                            }
                        }
                    }
                    if (++stage == kStagesB) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1 || (warp == kImIssuerWarp && prm.issuers == 2)) {
        // ===================== MMA issuers (whole warp waits, one elected lane issues) =====================
        // Optionally (SFA_ISSUERS=2) two warps share the issue work: warp 1 feeds the Re accumulator,
        // warp 14 the Im accumulator.  Measured on B200: no gain, the tensor pipe is not issue-bound.
        const bool im_part = (warp == kImIssuerWarp);
        const bool both = prm.issuers == 1;            // a single issuer feeds both accumulators
        // the last n-tile may be narrow: issue it with UMMA N = remaining rows rounded up to 16
        const int n_tail = (int)(prm.Nn - (long long)(prm.n_tiles - 1) * kTileN);
        const int n_tail16 = ((n_tail + 15) >> 4) << 4;
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0, a_phase = 0;
        for (long long grp = cluster_id; grp < n_groups; grp += n_clusters) {
            mbar_wait(&bars->a_full, a_phase);
            for (int nt = 0; nt < prm.n_tiles; ++nt) {
                mbar_wait(&bars->acc_empty[acc], acc_phase ^ 1);
                mbar_wait(&bars->b_full[stage], phase);
                tcgen05_fence_after();
                const int n_mma = (nt == prm.n_tiles - 1) ? n_tail16 : kTileN;
                const uint32_t idesc = make_idesc(kTileM, n_mma, false);
                const uint32_t idesc_neg = make_idesc(kTileM, n_mma, true);
                const uint32_t idesc_bf = make_idesc(kTileM, n_mma, false, true);
                const uint32_t idesc_bf_neg = make_idesc(kTileM, n_mma, true, true);
                if (elect_one_sync()) {
                    const uint32_t sbase = smem_u32(stage_base + (size_t)stage * kStageBytes);
                    const uint64_t bdesc = make_b_desc(sbase);          // plane 0, k-box 0, k-step 0
                    const uint32_t d_re = tmem + acc * kAccCols;
                    const uint32_t d_im = d_re + kTileN;
                    const uint32_t a0 = tmem + kOperandCol0;
                    constexpr int kPasses = kHybrid ? 2 : 3;
#pragma unroll
                    for (int pass = 0; pass < kPasses; ++pass) {
                        // planes: {re_hi, im_hi, re_x, im_x}; x = lo (fp32) or packed bf16 (hi, lo) pairs.
                        // 3xTF32: (A,B) = (hi,hi) (hi,lo) (lo,hi).  hybrid: tf32 (hi,hi), then bf16 (x,x).
                        const int ap = kHybrid ? (pass == 1 ? 2 : 0) : ((pass == 2) ? 2 : 0);
                        const int bp = kHybrid ? (pass == 1 ? 2 : 0) : ((pass == 1) ? 2 : 0);
                        const bool bf = kHybrid && pass == 1;
#pragma unroll
                        for (int ks = 0; ks < kMaxK / 8; ++ks) {
                            if (ks < ksteps) {
                                const int kb = ks >> 2, kk = ks & 3;
                                // the start-address field counts 16-byte units; offsets stay inside it
                                const uint64_t b_re = bdesc + (uint64_t)((((bp + 0) * 2 + kb) * kBoxBytes + kk * 32) >> 4);
                                const uint64_t b_im = bdesc + (uint64_t)((((bp + 1) * 2 + kb) * kBoxBytes + kk * 32) >> 4);
                                const uint32_t a_re = a0 + (ap + 0) * kMaxK + ks * 8;
                                const uint32_t a_im = a0 + (ap + 1) * kMaxK + ks * 8;
                                const uint32_t accum = (pass == 0 && ks == 0) ? 0u : 1u;
                                if (bf) {
                                    if (both || !im_part) {
                                        umma_bf16_ts(d_re, a_re, b_re, idesc_bf, accum);
                                        umma_bf16_ts(d_re, a_im, b_im, idesc_bf_neg, 1u);
                                    }
                                    if (both || im_part) {
                                        umma_bf16_ts(d_im, a_re, b_im, idesc_bf, accum);
                                        umma_bf16_ts(d_im, a_im, b_re, idesc_bf, 1u);
                                    }
                                } else {
                                    if (both || !im_part) {
                                        umma_tf32_ts(d_re, a_re, b_re, idesc, accum);
                                        umma_tf32_ts(d_re, a_im, b_im, idesc_neg, 1u);
                                    }
                                    if (both || im_part) {
                                        umma_tf32_ts(d_im, a_re, b_im, idesc, accum);
                                        umma_tf32_ts(d_im, a_im, b_re, idesc, 1u);
                                    }
                                }
                            }
                        }
                    }
                    if (csize == 1) tcgen05_commit(&bars->b_empty[stage]);
                    else tcgen05_commit_mc(&bars->b_empty[stage], cmask);   // frees the slot in every CTA of the cluster
                    tcgen05_commit(&bars->acc_full[acc]);
                    if (nt == prm.n_tiles - 1) tcgen05_commit(&bars->a_empty);
                }
                __syncwarp();
                if (++stage == kStagesB) {
                    stage = 0;
                    phase ^= 1;
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
            a_phase ^= 1;
        }
    } else if (warp < 6) {
        // ===================== A-operand loaders: global -> regs -> TMEM =====================
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may touch
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        uint32_t a_phase = 0;
        for (long long grp = cluster_id; grp < n_groups; grp += n_clusters) {
            // (b, m) of this lane's row; in folded mode consecutive rows run through several batches
            long long b = grp / prm.m_groups;
            int mt = (int)(grp - b * prm.m_groups) * csize + (int)crank;
            long long m = (long long)mt * kTileM + quarter * 32 + lane;
            bool m_ok = m < prm.Mm;
            if (prm.fold) {
                mt = (int)grp * csize + (int)crank;
                const long long mg = (long long)mt * kTileM + quarter * 32 + lane;
                m_ok = mg < prm.B * prm.Mm;
                b = m_ok ? mg / prm.Mm : 0;
                m = mg - b * prm.Mm;
            }
            const float2* src = prm.R + b * prm.strideR + (m_ok ? m : 0);
            float2 v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j)
                v[j] = (m_ok && j < prm.K) ? __ldg(src + (long long)j * prm.ldR) : make_float2(0.f, 0.f);
            // wait until the MMAs of the previous task no longer read the operand columns
            mbar_wait(&bars->a_empty, a_phase ^ 1);
            tcgen05_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                if (half == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int k = 32 + j;
                        v[j] = (m_ok && k < prm.K) ? __ldg(src + (long long)k * prm.ldR) : make_float2(0.f, 0.f);
                    }
                }
                uint32_t w[32];
                const uint32_t col = tmem + lane_base + kOperandCol0 + half * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) w[j] = to_tf32(v[j].x);
                SFA_TMEM_ST32(col + 0 * kMaxK, w);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float h = __uint_as_float(to_tf32(v[j].x));
                    w[j] = kHybrid ? pack_bf16(h, v[j].x - h) : __float_as_uint(v[j].x - h);   // A side: (hi, lo)
                }
                SFA_TMEM_ST32(col + 2 * kMaxK, w);
#pragma unroll
                for (int j = 0; j < 32; ++j) w[j] = to_tf32(v[j].y);
                SFA_TMEM_ST32(col + 1 * kMaxK, w);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float h = __uint_as_float(to_tf32(v[j].y));
                    w[j] = kHybrid ? pack_bf16(h, v[j].y - h) : __float_as_uint(v[j].y - h);
                }
                SFA_TMEM_ST32(col + 3 * kMaxK, w);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->a_full);
            a_phase ^= 1;
        }
    } else {
        // ===================== epilogue: TMEM -> regs -> global =====================
        // lane = m: one store instruction of a warp writes 32 consecutive complex64 of row n
        // (256 contiguous bytes).  The column loop is kept to a pointer bump + store per value;
        // mode / row-scale / ragged-tile variants are hoisted out of it.
        // Eight warps: warp w and w+4 share TMEM lane quarter (w & 3) and take 32 columns each.
        const int quarter = warp & 3;
        const int chalf = (warp - 6) >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        int acc = 0, ring = 0;
        uint32_t acc_phase = 0;
        for (long long grp = cluster_id; grp < n_groups; grp += n_clusters) {
            long long b = grp / prm.m_groups;
            int mt = (int)(grp - b * prm.m_groups) * csize + (int)crank;
            long long m = (long long)mt * kTileM + quarter * 32 + lane;
            bool m_ok = m < prm.Mm;
            if (prm.fold) {
                mt = (int)grp * csize + (int)crank;
                const long long mg = (long long)mt * kTileM + quarter * 32 + lane;
                m_ok = mg < prm.B * prm.Mm;
                b = m_ok ? mg / prm.Mm : 0;
                m = mg - b * prm.Mm;
            }
            const bool all_ok = __all_sync(0xffffffffu, m_ok);   // false only in a ragged last m-tile
            float2* const out_bm = reinterpret_cast<float2*>(prm.out) + b * prm.strideO + (m_ok ? m : 0);
            const float2* const scale_b = prm.rowscale ? prm.rowscale + b * prm.strideS : nullptr;
            for (int nt = 0; nt < prm.n_tiles; ++nt) {
                mbar_wait(&bars->acc_full[acc], acc_phase);
                tcgen05_fence_after();
                const uint32_t t_re = tmem + lane_base + acc * kAccCols + chalf * 32;
                uint32_t re[32], im[32];
                SFA_TMEM_LD32(t_re, re);
                SFA_TMEM_LD32(t_re + kTileN, im);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // this warp's part of the accumulator is in registers -> release it before storing
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->acc_empty[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
                const long long nbase = (long long)nt * kTileN + chalf * 32;
                const long long left = prm.Nn - nbase;
                const int ncols = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
                if (prm.tma_store) {
                    // ---- TMEM -> regs -> smem (row = n, 128 frames contiguous) -> TMA store ----
                    // The four warps of this column half share one 16 KB buffer: 2 x 16 rows per n-tile.
                    // TMA clips rows >= Nn and frames >= Mm, so ragged tiles need no predicates.
                    if (prm.out_mode == 99) continue;
                    if (scale_b) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (j < ncols) {
                                const float2 sc = __ldg(scale_b + nbase + j);
                                const float xr = __uint_as_float(re[j]), xi = __uint_as_float(im[j]);
                                re[j] = __float_as_uint(xr * sc.x - xi * sc.y);
                                im[j] = __float_as_uint(xr * sc.y + xi * sc.x);
                            }
                        }
                    }
                    const bool issuer = (quarter == 0) && (lane == 0);
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        unsigned char* const ring_buf = store_base + (size_t)(chalf * kStoreRing + ring) * kStoreBytes;
                        float2* sbuf = reinterpret_cast<float2*>(ring_buf) + quarter * 32 + lane;
                        // the store that used this buffer kStoreRing parts ago must have finished reading shared memory
                        if (issuer) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kStoreRing - 1) : "memory");
                        named_bar_sync(1 + chalf, 128);
#pragma unroll
                        for (int j = 0; j < kStoreRows; ++j)
                            sbuf[j * kTileM] = make_float2(__uint_as_float(re[part * kStoreRows + j]),
                                                           __uint_as_float(im[part * kStoreRows + j]));
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        named_bar_sync(1 + chalf, 128);
                        if (issuer && ncols > part * kStoreRows) {
                            tma_store_3d(&o_map, ring_buf, 2 * mt * kTileM,
                                         (int)nbase + part * kStoreRows, (int)b, prm.out_mode == OUT_C64_ACCUM);
                        }
                        // (an empty group keeps the group <-> ring-slot correspondence when nothing was stored)
                        if (issuer) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        if (++ring == kStoreRing) ring = 0;
                    }
                    continue;
                }
                if (ncols == 0 || prm.out_mode == 99) continue;          // warp-uniform
                if (!all_ok && !m_ok) continue;                           // ragged m-tile: scalar path, no shuffles
                float2* o = out_bm + nbase * prm.ldo;
                long long ldo_eff = prm.ldo;
                if (scale_b) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < ncols) {
                            const float2 sc = __ldg(scale_b + nbase + j);
                            const float xr = __uint_as_float(re[j]), xi = __uint_as_float(im[j]);
                            re[j] = __float_as_uint(xr * sc.x - xi * sc.y);
                            im[j] = __float_as_uint(xr * sc.y + xi * sc.x);
                        }
                    }
                }
                if (prm.out_mode == OUT_C64_ACCUM) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < ncols) {
                            const float2 old = o[(long long)j * prm.ldo];
                            re[j] = __float_as_uint(__uint_as_float(re[j]) + old.x);
                            im[j] = __float_as_uint(__uint_as_float(im[j]) + old.y);
                        }
                    }
                }
                // Full tiles with 16-byte aligned rows: lane pairs swap one value per column pair so that
                // every lane owns two adjacent frames of one row -> 16-byte stores, half as many store
                // instructions in flight for the same bytes (the LSU store queue is what limits here).
                const bool vec_ok = all_ok && (ncols == 32) && ((ldo_eff & 1) == 0) &&
                                    __all_sync(0xffffffffu, (((size_t)o) & 15) == ((lane & 1) ? 8u : 0u));
                if (vec_ok) {
                    const bool odd = lane & 1;
                    float2* ov = odd ? (o + ldo_eff - 1) : o;       // odd lanes own row j+1, frames (t-1, t)
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const uint32_t sr = odd ? re[j] : re[j + 1];
                        const uint32_t si = odd ? im[j] : im[j + 1];
                        const uint32_t gr = __shfl_xor_sync(0xffffffffu, sr, 1);
                        const uint32_t gi = __shfl_xor_sync(0xffffffffu, si, 1);
                        float4 v;
                        if (odd) v = make_float4(__uint_as_float(gr), __uint_as_float(gi), __uint_as_float(re[j + 1]), __uint_as_float(im[j + 1]));
                        else v = make_float4(__uint_as_float(re[j]), __uint_as_float(im[j]), __uint_as_float(gr), __uint_as_float(gi));
                        *reinterpret_cast<float4*>(ov) = v;
                        ov += 2 * ldo_eff;
                    }
                } else if (ncols == 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        *o = make_float2(__uint_as_float(re[j]), __uint_as_float(im[j]));
                        o += ldo_eff;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < ncols) *o = make_float2(__uint_as_float(re[j]), __uint_as_float(im[j]));
                        o += ldo_eff;
                    }
                }
            }
        }
    }

    // outstanding bulk stores of this thread (epilogue issuers) must complete before the CTA exits
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();                 // no CTA leaves while a peer may still write into it
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}

// ------------------------------------------------------------------ host side
// P planar-split layout: [batches][4 planes][Nn][Kp] fp32, Kp % 4 == 0.
int launch_cgemm3x(long long B, long long Nn, long long Mm, int K, const float* Pplanar, long long Kp,
                   int p_shared, const float2* R, long long ldR, long long strideR, float* out, long long ldo,
                   long long strideO, int out_mode, const float2* rowscale, long long strideS, int hybrid,
                   cudaStream_t st) {
    if (B <= 0 || Nn <= 0 || Mm <= 0) return SFA_OK;
    if (K <= 0 || K > kMaxK || (Kp & 3) || Kp < K) return SFA_ERR_INVALID;
    if (((uintptr_t)Pplanar & 15) || ((uintptr_t)R & 7)) return SFA_ERR_INVALID;
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return SFA_ERR_UNSUPPORTED;
    CUtensorMap map;
    const long long pb = p_shared ? 1 : B;
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)Nn, (cuuint64_t)(4 * pb)};
    cuuint64_t gstr[2] = {(cuuint64_t)Kp * 4, (cuuint64_t)Nn * Kp * 4};
    cuuint32_t box[3] = {(cuuint32_t)kBoxK, (cuuint32_t)kTileN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)Pplanar, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
        return SFA_ERR_CUDA;
    }
    // Output tensor map for the TMA-store epilogue: fp32 view [B][Nn][2*ldo], box = 16 rows x 128 complex.
    // Small m extents with a shared P (factored form at small T): fold batches into the m-tiles.
    // Only for single-pass products: the split-K accumulate passes rely on the TMA reduce-add store,
    // which needs whole tiles inside one batch (measured: folding + register read-modify-write is slower).
    const int fold = (p_shared && Mm <= 64 && B > 1 && out_mode == OUT_C64 && !options().no_fold) ? 1 : 0;
    // (a TMA store clips a box in 16-byte units: with an odd number of frames per row -- 8 Mm bytes -- it would also
    // write the element behind the row end, which belongs to the caller when the rows are padded)
    const int tma_store = (!fold && (ldo & 1) == 0 && (Mm & 1) == 0 && (strideO & 1) == 0 && (((uintptr_t)out) & 15) == 0 &&
                           !options().no_tma_store) ? 1 : 0;
    CUtensorMap omap = map;
    if (tma_store) {
        cuuint64_t odim[3] = {(cuuint64_t)(2 * Mm), (cuuint64_t)Nn, (cuuint64_t)B};
        cuuint64_t ostr[2] = {(cuuint64_t)ldo * 8, (cuuint64_t)strideO * 8};
        cuuint32_t obox[3] = {(cuuint32_t)(2 * kTileM), (cuuint32_t)kStoreRows, 1};
        CUresult r2 = encode(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)out, odim, ostr, obox, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r2 != CUDA_SUCCESS) {
            set_last_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(out)");
            return SFA_ERR_CUDA;
        }
    }
    GemmParams prm;
    prm.tma_store = tma_store;
    prm.B = B;
    prm.Nn = Nn;
    prm.Mm = Mm;
    prm.K = K;
    prm.R = R;
    prm.ldR = ldR;
    prm.strideR = strideR;
    prm.p_shared = p_shared;
    prm.out = out;
    prm.ldo = ldo;
    prm.strideO = strideO;
    prm.rowscale = rowscale;
    prm.strideS = strideS;
#ifdef SFA_DEBUG_HOOKS   // timing-attribution experiments only; never compiled into the shipped library
    prm.out_mode = getenv("SFA_DEBUG_NOSTORE") ? 99 : out_mode;
    prm.debug_no_tma = getenv("SFA_DEBUG_NOTMA") ? 1 : 0;
#else
    prm.out_mode = out_mode;
    prm.debug_no_tma = 0;
#endif
    prm.fold = fold;
    prm.m_tiles = (int)ceil_div(fold ? B * Mm : Mm, kTileM);
    prm.tasks = B * prm.m_tiles;
    prm.n_tiles = (int)ceil_div(Nn, kTileN);
    const size_t smem = (size_t)kStagesB * kStageBytes + 2 * kStoreRing * kStoreBytes + sizeof(Barriers) + 1024;
    if (hybrid) SFA_ENSURE_DYN_SMEM(cgemm3x_kernel<true>, smem);
    else SFA_ENSURE_DYN_SMEM(cgemm3x_kernel<false>, smem);
    // CTA pairs share the P stream by TMA multicast whenever a batch has at least two m-tiles.
    const int csize = (prm.m_tiles >= 2 && !options().no_cluster) ? 2 : 1;
    prm.csize = csize;
    prm.m_groups = (prm.m_tiles + csize - 1) / csize;
    const long long n_groups = (fold ? 1 : B) * prm.m_groups;
    const long long max_clusters = persistent_sms() / csize;
    const long long grid = (n_groups < max_clusters ? n_groups : max_clusters) * csize;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    // One issuer warp is enough (measured: a second one, splitting Re/Im, changes nothing) and keeps
    // the CTA at 448 threads x 128 registers, which leaves room for a steering-build CTA on the SM.
    prm.issuers = options().issuers == 2 ? 2 : 1;
    cfg.blockDim = dim3(prm.issuers == 2 ? kThreads : kThreads - 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = csize;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    prof_begin(st);
    cudaError_t le = hybrid ? cudaLaunchKernelEx(&cfg, cgemm3x_kernel<true>, map, omap, prm)
                            : cudaLaunchKernelEx(&cfg, cgemm3x_kernel<false>, map, omap, prm);
    prof_end(st);
    if (le != cudaSuccess) {
        set_last_cuda_error(le, "cudaLaunchKernelEx(cgemm3x_kernel)");
        return SFA_ERR_CUDA;
    }
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

}  // namespace sfa
