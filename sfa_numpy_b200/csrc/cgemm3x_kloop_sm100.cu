// Complex GEMM with a shared left operand and a LONG inner dimension (K > 64) on tcgen05:
//
//     out[b, n, m] = (sum_k P[n, k] * R[b, k, m]) * rowscale[b, n]
//
// This is the factored form of the plane-wave decomposition when the array has many channels
// (BASELINE configs 4 and 5: K = 289 ... 1089):
//     Z_f = d_f * (Y_Q^H X_f)      P = conj(Y_Q)^T [M x Q],  R = X [F][Q][T],  rowscale = d
//     q_f = Y_L Z_f                P = Y_L [L x M],          R = Z [F][M][T]
// i.e. the two np.matmul of the reference idiom with the diagonal applied as a row scale
// (doc/examples/modal_beamforming_rigid_spherical_array.py:38-39, radial.py:269-314).
//
// cgemm3x_sm100.cu keeps the whole K extent (<= 64) of the R tile resident in tensor memory and
// needs one launch (and one pass over the output) per 64-wide K chunk.  Here the K loop runs
// INSIDE the kernel: the accumulators stay in TMEM for the whole inner dimension and the output
// is written exactly once.
//
// Mapping (one persistent CTA per SM, 448 threads = TMA + MMA + 8 loader + 4 epilogue warps; CTA pairs
// share the P stream):
//   * rows (TMEM lanes) = 128 consecutive entries of the folded (batch, m) axis -- all batches
//     share P, so small m extents (T = 32, 64 frames) fill the tile with several bins;
//   * columns = 128 rows n of P per task; accumulators Re, Im: 2 x 128 TMEM columns;
//   * a task is (pair of m-tiles, n-tile); its K loop advances 32 complex k per step:
//       - four loader warps read the R chunk [32 k x 128 m] from global memory (coalesced along
//         m), split it (tf32 hi + lo / packed bf16) and write the four operand planes into a
//         TMEM operand buffer (tcgen05.st) -- the A operand of a TS-form tcgen05.mma.  There are
//         two such buffers, each owned by its own set of four loader warps (even / odd chunks),
//         so that the global-load latency of one chunk hides behind the MMAs of two;
//       - one elected thread streams the matching P chunk [128 n x 32 k x 4 planes] = 64 KB by
//         TMA into a 3-stage shared-memory ring; each CTA of the pair fetches half of the planes
//         and multicasts them to both;
//       - one elected thread issues 2 (hybrid) or 3 (3xTF32) passes x 4 real products x 4
//         k-steps of tcgen05.mma M128 x N128 x K8 and releases the stage / operand buffer with
//         tcgen05.commit;
//   * after the last chunk four epilogue warps read the accumulators (tcgen05.ld), apply the
//     row scale and store complex64: a warp writes 32 consecutive m of one row n per
//     instruction (256 contiguous bytes).
//   * tasks are ordered in bands of 8 n-tiles so that the P rows of a band (and the R tiles the
//     concurrently running pairs share) stay in L2 while R streams through once per band.
//   * BLOCKED ACCUMULATION.  The tensor core adds into the FP32 accumulator with truncation, which
//     biases every add by ~3e-8 of the running sum: measured normwise error 1.0e-8 K (hybrid) /
//     1.5e-8 K (3xTF32), i.e. > 1e-5 at K = 1089.  The K loop is therefore cut into blocks of
//     `block_chunks` chunks; after each block the epilogue warps move the accumulators through two
//     16 KB shared-memory staging buffers into an L2-resident per-CTA scratch tile with bulk
//     copies -- a plain copy for the first block, cp.reduce.async.bulk ... .add.f32 (FP32
//     round-to-nearest add at L2) for the others -- and release them; after the last block they
//     read the tile back, apply the row scale and store.  The error then stays at the
//     single-block level whatever K is.  K <= one block: TMEM -> registers -> output directly.
// TMEM: 256 accumulator columns + 2 x 128 operand columns = 512.  SMEM: 3 x 64 KB + 2 x 16 KB.
#include <cuda.h>
#include <stdlib.h>
#include "sfa_common.cuh"
#include "sfa_tcgen05.cuh"

namespace sfa {
namespace kloop {

constexpr int kTileM = 128;                       // folded (batch, m) rows per CTA task
constexpr int kTileN = 128;                       // rows of P per task
constexpr int kChunkK = 32;                       // complex k per pipeline step (one 128-byte swizzle row)
constexpr int kStages = 3;
constexpr int kPlaneBytes = kTileN * kChunkK * 4; // 16 KB
constexpr int kStageBytes = 4 * kPlaneBytes;      // 64 KB
constexpr int kThreads = 448;                     // TMA, MMA, 2 x 4 loader warps, 4 epilogue warps
constexpr int kEpilogueWarp0 = 10;
constexpr int kTmemCols = 512;
constexpr int kAccIm = kTileN;                    // Im accumulator column offset
constexpr int kOperandCol0 = 2 * kTileN;          // 256
constexpr int kOperandCols = 4 * kChunkK;         // 128 per buffer
constexpr int kBand = 8;                          // n-tiles per L2 band
constexpr int kPieceCols = 16;                    // accumulator columns per staging piece
constexpr int kPieceBytes = kPieceCols * kTileM * 8;   // 16 KB
constexpr int kWarpPieceBytes = kPieceCols * 32 * 8;   // 4 KB: one epilogue warp's share of a piece

struct Params {
    long long B, Nn, Mm;          // batches, rows of P, m extent per batch
    long long rows;               // B * Mm (folded)
    int K;
    const float2* R;              // [B][K][ldR]
    long long ldR, strideR;
    float2* out;                  // [B][Nn][ldo]
    long long ldo, strideO;
    const float2* rowscale;       // [B][Nn] or null
    long long strideS;
    int n_tiles, m_groups, csize;
    long long tasks;              // n_tiles * m_groups
    int kchunks;
    int block_chunks, nblocks;    // accumulation blocks (chunks per block, blocks per task)
    float2* scratch;              // [grid][4 lane quarters][128 n][32 rows] running sum of the block partials
    int store_evict_first;        // output much larger than L2: its stores carry the evict_first policy, so that the
                                  // written lines do not push the left-operand band and the X / Z chunks out of L2
};

struct __align__(8) Barriers {
    uint64_t b_full[kStages];
    uint64_t b_empty[kStages];
    uint64_t a_full[2];
    uint64_t a_empty[2];
    uint64_t acc_full;
    uint64_t acc_empty;
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void task_coords(const Params& prm, long long task, int& mg, int& nt) {
    const long long per_band = (long long)kBand * prm.m_groups;
    const long long band = task / per_band;
    const long long r = task - band * per_band;
    const int left = prm.n_tiles - (int)band * kBand;
    const int nb = left < kBand ? left : kBand;
    mg = (int)(r / nb);
    nt = (int)band * kBand + (int)(r - (long long)mg * nb);
}

// 32 accumulator columns of one row m (this lane): row scale, clip at Nn, store complex64.
__device__ __forceinline__ void store_columns(const Params& prm, uint32_t (&re)[32], uint32_t (&im)[32],
                                              float2* out_bm, const float2* scale_b, long long nbase) {
    const long long left = prm.Nn - nbase;
    const int ncols = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
    if (ncols == 0) return;
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
    float2* o = out_bm + nbase * prm.ldo;
    if (ncols == 32 && prm.store_evict_first) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(o), "r"(re[j]), "r"(im[j]), "l"(pol)
                         : "memory");
            o += prm.ldo;
        }
    } else if (ncols == 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            *o = make_float2(__uint_as_float(re[j]), __uint_as_float(im[j]));
            o += prm.ldo;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j < ncols) *o = make_float2(__uint_as_float(re[j]), __uint_as_float(im[j]));
            o += prm.ldo;
        }
    }
}

template <bool kHybrid>
__global__ void __launch_bounds__(kThreads, 1)
cgemm3x_kloop_kernel(const __grid_constant__ CUtensorMap p_map, const Params prm) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;
    unsigned char* piece_base = smem + (size_t)kStages * kStageBytes;          // 2 x 16 KB staging
    Barriers* bars = reinterpret_cast<Barriers*>(piece_base + 2 * kPieceBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int csize = prm.csize;
    const uint32_t crank = (csize > 1) ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
    const long long cluster_id = blockIdx.x / csize, n_clusters = gridDim.x / csize;
    const int kchunks = prm.kchunks;
    const int tail_ksteps = (prm.K - (kchunks - 1) * kChunkK + 7) >> 3;      // MMA k-steps of the last chunk

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars->b_full[s], 1);
            mbar_init(&bars->b_empty[s], csize);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->a_full[a], 4);
            mbar_init(&bars->a_empty[a], 1);
        }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->acc_empty, 4);
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
    if (csize > 1) cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer: P chunks [128 n x 32 k x 4 planes] =====================
        if (elect_one_sync()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&p_map) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (long long task = cluster_id; task < prm.tasks; task += n_clusters) {
                int mg, nt;
                task_coords(prm, task, mg, nt);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&bars->b_empty[stage], phase ^ 1);
                    unsigned char* dst = stage_base + (size_t)stage * kStageBytes;
                    mbar_arrive_expect_tx(&bars->b_full[stage], (uint32_t)kStageBytes);
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl) {
                        if (csize == 1)
                            tma_load_3d(&p_map, &bars->b_full[stage], dst + pl * kPlaneBytes, kc * kChunkK, nt * kTileN, pl);
                        else if ((pl & 1) == (int)crank)
                            tma_load_3d_mc(&p_map, &bars->b_full[stage], dst + pl * kPlaneBytes, kc * kChunkK,
                                           nt * kTileN, pl, cmask);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const int n_tail = (int)(prm.Nn - (long long)(prm.n_tiles - 1) * kTileN);
        const int n_tail16 = ((n_tail + 15) >> 4) << 4;
        int stage = 0, ob = 0;
        uint32_t phase = 0, a_phase = 0, acc_phase = 0;
        for (long long task = cluster_id; task < prm.tasks; task += n_clusters) {
            int kin = 0;                                         // chunk index inside the accumulation block
            int mg, nt;
            task_coords(prm, task, mg, nt);
            const int n_mma = (nt == prm.n_tiles - 1) ? n_tail16 : kTileN;
            const uint32_t idesc = make_idesc(kTileM, n_mma, false);
            const uint32_t idesc_neg = make_idesc(kTileM, n_mma, true);
            const uint32_t idesc_bf = make_idesc(kTileM, n_mma, false, true);
            const uint32_t idesc_bf_neg = make_idesc(kTileM, n_mma, true, true);
            for (int kc = 0; kc < kchunks; ++kc) {
                if (kin == 0) {                                  // the epilogue has drained the previous block
                    mbar_wait(&bars->acc_empty, acc_phase ^ 1);
                    acc_phase ^= 1;
                }
                const bool block_end = (kin == prm.block_chunks - 1) || (kc == kchunks - 1);
                mbar_wait(&bars->a_full[ob], a_phase);
                mbar_wait(&bars->b_full[stage], phase);
                tcgen05_fence_after();
                const int ksteps = (kc == kchunks - 1) ? tail_ksteps : kChunkK / 8;
                if (elect_one_sync()) {
                    const uint64_t bdesc = make_b_desc(smem_u32(stage_base + (size_t)stage * kStageBytes));
                    const uint32_t d_re = tmem, d_im = tmem + kAccIm;
                    const uint32_t a0 = tmem + kOperandCol0 + ob * kOperandCols;
                    constexpr int kPasses = kHybrid ? 2 : 3;
#pragma unroll
                    for (int pass = 0; pass < kPasses; ++pass) {
                        // planes {re_hi, im_hi, re_x, im_x}.  3xTF32: (A,B) = (hi,hi) (hi,lo) (lo,hi);
                        // hybrid: tf32 (hi,hi), then one bf16 pass over the packed (hi,lo) x (lo,hi) planes.
                        const int ap = kHybrid ? (pass == 1 ? 2 : 0) : ((pass == 2) ? 2 : 0);
                        const int bp = kHybrid ? (pass == 1 ? 2 : 0) : ((pass == 1) ? 2 : 0);
                        const bool bf = kHybrid && pass == 1;
#pragma unroll
                        for (int ks = 0; ks < kChunkK / 8; ++ks) {
                            if (ks < ksteps) {
                                const uint64_t b_re = bdesc + (uint64_t)(((bp + 0) * kPlaneBytes + ks * 32) >> 4);
                                const uint64_t b_im = bdesc + (uint64_t)(((bp + 1) * kPlaneBytes + ks * 32) >> 4);
                                const uint32_t a_re = a0 + (ap + 0) * kChunkK + ks * 8;
                                const uint32_t a_im = a0 + (ap + 1) * kChunkK + ks * 8;
                                const uint32_t accum = (kin == 0 && pass == 0 && ks == 0) ? 0u : 1u;
                                if (bf) {
                                    umma_bf16_ts(d_re, a_re, b_re, idesc_bf, accum);
                                    umma_bf16_ts(d_re, a_im, b_im, idesc_bf_neg, 1u);
                                    umma_bf16_ts(d_im, a_re, b_im, idesc_bf, accum);
                                    umma_bf16_ts(d_im, a_im, b_re, idesc_bf, 1u);
                                } else {
                                    umma_tf32_ts(d_re, a_re, b_re, idesc, accum);
                                    umma_tf32_ts(d_re, a_im, b_im, idesc_neg, 1u);
                                    umma_tf32_ts(d_im, a_re, b_im, idesc, accum);
                                    umma_tf32_ts(d_im, a_im, b_re, idesc, 1u);
                                }
                            }
                        }
                    }
                    if (csize == 1) tcgen05_commit(&bars->b_empty[stage]);
                    else tcgen05_commit_mc(&bars->b_empty[stage], cmask);
                    tcgen05_commit(&bars->a_empty[ob]);
                    if (block_end) tcgen05_commit(&bars->acc_full);
                }
                __syncwarp();
                kin = block_end ? 0 : kin + 1;
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1;
                }
                if (++ob == 2) {
                    ob = 0;
                    a_phase ^= 1;
                }
            }
        }
    } else if (warp < kEpilogueWarp0) {
        // ===================== R loaders: global -> regs -> split -> TMEM operand buffer =====================
        // set 0 (warps 2-5) owns operand buffer 0 and the even chunks of the CTA's chunk sequence,
        // set 1 (warps 6-9) buffer 1 and the odd chunks.
        const int quarter = warp & 3;
        const int ob = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const uint32_t col = tmem + lane_base + kOperandCol0 + ob * kOperandCols;
        uint32_t a_phase = 0;
        long long seq = 0;                                   // running chunk index of this CTA
        for (long long task = cluster_id; task < prm.tasks; task += n_clusters) {
            int mg, nt;
            task_coords(prm, task, mg, nt);
            const long long row = ((long long)mg * csize + crank) * kTileM + quarter * 32 + lane;
            const bool m_ok = row < prm.rows;
            const long long b = m_ok ? row / prm.Mm : 0;
            const long long m = m_ok ? row - b * prm.Mm : 0;
            const float2* src = prm.R + b * prm.strideR + m;
            for (int kc = (int)((seq ^ ob) & 1); kc < kchunks; kc += 2) {
                const int k0 = kc * kChunkK;
                float2 v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = (m_ok && k0 + j < prm.K) ? __ldg(src + (long long)(k0 + j) * prm.ldR) : make_float2(0.f, 0.f);
                mbar_wait(&bars->a_empty[ob], a_phase ^ 1);      // the MMAs that read this buffer have completed
                tcgen05_fence_after();
                uint32_t w[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) w[j] = to_tf32(v[j].x);
                SFA_TMEM_ST32(col + 0 * kChunkK, w);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float h = __uint_as_float(to_tf32(v[j].x));
                    w[j] = kHybrid ? pack_bf16(h, v[j].x - h) : __float_as_uint(v[j].x - h);
                }
                SFA_TMEM_ST32(col + 2 * kChunkK, w);
#pragma unroll
                for (int j = 0; j < 32; ++j) w[j] = to_tf32(v[j].y);
                SFA_TMEM_ST32(col + 1 * kChunkK, w);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float h = __uint_as_float(to_tf32(v[j].y));
                    w[j] = kHybrid ? pack_bf16(h, v[j].y - h) : __float_as_uint(v[j].y - h);
                }
                SFA_TMEM_ST32(col + 3 * kChunkK, w);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->a_full[ob]);
                a_phase ^= 1;
            }
            seq += kchunks;
        }
    } else {
        // ===================== epilogue: TMEM -> regs -> (row scale) -> global =====================
        const int quarter = warp & 3;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        uint32_t acc_phase = 0;
        for (long long task = cluster_id; task < prm.tasks; task += n_clusters) {
            int mg, nt;
            task_coords(prm, task, mg, nt);
            const long long row = ((long long)mg * csize + crank) * kTileM + quarter * 32 + lane;
            const bool m_ok = row < prm.rows;
            const long long b = m_ok ? row / prm.Mm : 0;
            const long long m = m_ok ? row - b * prm.Mm : 0;
            float2* const out_bm = prm.out + b * prm.strideO + m;
            const float2* const scale_b = prm.rowscale ? prm.rowscale + b * prm.strideS : nullptr;
            if (prm.nblocks == 1) {
                // ---- single block: TMEM -> registers -> output ----
                mbar_wait(&bars->acc_full, acc_phase);
                acc_phase ^= 1;
                tcgen05_fence_after();
#pragma unroll 1
                for (int round = 0; round < kTileN / 32; ++round) {
                    const int c0 = round * 32;
                    uint32_t re[32], im[32];
                    __syncwarp();                                // converged for the .sync.aligned TMEM loads
                    SFA_TMEM_LD32(tmem + lane_base + c0, re);
                    SFA_TMEM_LD32(tmem + lane_base + kAccIm + c0, im);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (round == kTileN / 32 - 1) {              // the whole accumulator has been read
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->acc_empty);
                    }
                    if (m_ok) store_columns(prm, re, im, out_bm, scale_b, (long long)nt * kTileN + c0);
                }
                __syncwarp();
                continue;
            }
            // ---- several blocks: accumulators -> staging -> scratch tile (bulk copy / reduce-add) ----
            // Everything is warp-local: a warp owns the 32 rows of its TMEM lane quarter, two 4 KB staging
            // buffers and the [128 n][32 rows] quarter of the CTA's scratch tile, so that a piece of 16
            // columns is one contiguous 4 KB bulk copy and no cross-warp barrier is needed.
            float2* const tile = prm.scratch + ((size_t)blockIdx.x * 4 + quarter) * kTileN * 32;
            unsigned char* const wbuf = piece_base + quarter * (2 * kWarpPieceBytes);
            for (int blk = 0; blk < prm.nblocks; ++blk) {
                mbar_wait(&bars->acc_full, acc_phase);
                acc_phase ^= 1;
                tcgen05_fence_after();
                uint32_t re[16], im[16];
                SFA_TMEM_LD16(tmem + lane_base, re);
                SFA_TMEM_LD16(tmem + lane_base + kAccIm, im);
#pragma unroll 1
                for (int piece = 0; piece < kTileN / kPieceCols; ++piece) {
                    unsigned char* buf = wbuf + (piece & 1) * kWarpPieceBytes;
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    SFA_PIN16(re);
                    SFA_PIN16(im);
                    float2 v[16];
#pragma unroll
                    for (int j = 0; j < kPieceCols; ++j) v[j] = make_float2(__uint_as_float(re[j]), __uint_as_float(im[j]));
                    if (piece + 1 < kTileN / kPieceCols) {       // next piece's TMEM read overlaps this piece's staging
                        SFA_TMEM_LD16(tmem + lane_base + (piece + 1) * kPieceCols, re);
                        SFA_TMEM_LD16(tmem + lane_base + kAccIm + (piece + 1) * kPieceCols, im);
                    } else {                                     // the whole accumulator has been read
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->acc_empty);
                    }
                    // the bulk copy that read this buffer two pieces ago has finished reading it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                    float2* sb = reinterpret_cast<float2*>(buf) + lane;
#pragma unroll
                    for (int j = 0; j < kPieceCols; ++j) sb[j * 32] = v[j];
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        bulk_store_1d(tile + (size_t)piece * kPieceCols * 32, buf, kWarpPieceBytes, blk > 0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            // all partial sums are in the scratch tile: read it back (L2), scale, store
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
            if (m_ok) {
#pragma unroll 1
                for (int round = 0; round < kTileN / 32; ++round) {
                    const int c0 = round * 32;
                    const long long nbase = (long long)nt * kTileN + c0;
                    if (nbase >= prm.Nn) break;
                    const float2* sp = tile + (size_t)c0 * 32 + lane;
                    uint32_t re[32], im[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float2 v = __ldcg(sp + (size_t)j * 32);
                        re[j] = __float_as_uint(v.x);
                        im[j] = __float_as_uint(v.y);
                    }
                    store_columns(prm, re, im, out_bm, scale_b, nbase);
                }
            }
            __syncwarp();
        }
    }

    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();                 // no CTA leaves while a peer may still multicast into it
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}

}  // namespace kloop

// Accumulation block (in 32-k chunks): keeps the truncation bias of the tensor-core accumulate at
// ~6e-6 of the result through both GEMMs of the factored form (measured 2.1e-8 / 3.2e-8 per k).
// At most 10 (hybrid) / 7 (3xTF32) chunks per block, the chunks spread evenly over the blocks: K = 289
// (cfg 4, second product) is ONE block of 10 -- accumulators straight to the output, no scratch round trip:
// 1.98 -> 1.64 ms for cfg 4 -- and K = 1089 (cfg 5) stays at 4 blocks of 9 (error 5.9e-6; 3 blocks of 12
// would be 1 % faster at 7.6e-6).  tools/kloop_block_probe.py.
static int kloop_block_chunks(long long kchunks, int hybrid) {
    if (options().kloop_block > 0) return options().kloop_block;
    const long long cap = hybrid ? 10 : 7;
    const long long nblocks = ceil_div(kchunks, cap);
    return (int)ceil_div(kchunks, nblocks < 1 ? 1 : nblocks);
}

// Scratch the launch needs for a product with inner dimension K (0 when one block covers K).
size_t cgemm3x_kloop_scratch_bytes(long long K, int hybrid) {
    using namespace kloop;
    const long long kchunks = ceil_div(K, kChunkK);
    const long long nblocks = ceil_div(kchunks, kloop_block_chunks(kchunks, hybrid));
    if (nblocks <= 1) return 0;
    return (size_t)sm_count() * kTileN * kTileM * sizeof(float2);
}

// P planar-split layout: [4 planes][Nn][Kp] fp32 (Kp % 4 == 0), shared by all batches.
int launch_cgemm3x_kloop(long long B, long long Nn, long long Mm, long long K, const float* Pplanar, long long Kp,
                         const float2* R, long long ldR, long long strideR, float2* out, long long ldo,
                         long long strideO, const float2* rowscale, long long strideS, int hybrid,
                         void* scratch, size_t scratch_bytes, cudaStream_t st) {
    using namespace kloop;
    if (B <= 0 || Nn <= 0 || Mm <= 0) return SFA_OK;
    if (K <= 0 || K > (1 << 20) || (Kp & 3) || Kp < K) return SFA_ERR_INVALID;
    if (((uintptr_t)Pplanar & 15) || ((uintptr_t)R & 7) || ((uintptr_t)out & 7)) return SFA_ERR_INVALID;
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return SFA_ERR_UNSUPPORTED;
    CUtensorMap map;
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)Nn, 4};
    cuuint64_t gstr[2] = {(cuuint64_t)Kp * 4, (cuuint64_t)Nn * Kp * 4};
    cuuint32_t box[3] = {(cuuint32_t)kChunkK, (cuuint32_t)kTileN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)Pplanar, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(kloop P)");
        return SFA_ERR_CUDA;
    }
    Params prm;
    prm.B = B;
    prm.Nn = Nn;
    prm.Mm = Mm;
    prm.rows = B * Mm;
    prm.K = (int)K;
    prm.R = R;
    prm.ldR = ldR;
    prm.strideR = strideR;
    prm.out = out;
    prm.ldo = ldo;
    prm.store_evict_first = (options().kloop_no_hint == 0 && (double)B * (double)Nn * (double)ldo * 8.0 > 256e6) ? 1 : 0;
    prm.strideO = strideO;
    prm.rowscale = rowscale;
    prm.strideS = strideS;
    const long long m_tiles = ceil_div(prm.rows, kTileM);
    const int csize = (m_tiles >= 2 && !options().no_cluster) ? 2 : 1;
    prm.csize = csize;
    prm.m_groups = (int)ceil_div(m_tiles, csize);
    prm.n_tiles = (int)ceil_div(Nn, kTileN);
    prm.tasks = (long long)prm.n_tiles * prm.m_groups;
    prm.kchunks = (int)ceil_div(K, kChunkK);
    prm.block_chunks = kloop_block_chunks(prm.kchunks, hybrid);
    prm.nblocks = (int)ceil_div(prm.kchunks, prm.block_chunks);
    prm.scratch = reinterpret_cast<float2*>(scratch);
    if (prm.nblocks > 1 && (!scratch || scratch_bytes < cgemm3x_kloop_scratch_bytes(K, hybrid) || ((uintptr_t)scratch & 7)))
        return SFA_ERR_WORKSPACE;
    const size_t smem = (size_t)kStages * kStageBytes + 2 * kPieceBytes + sizeof(Barriers) + 1024;
    if (hybrid) SFA_ENSURE_DYN_SMEM(cgemm3x_kloop_kernel<true>, smem);
    else SFA_ENSURE_DYN_SMEM(cgemm3x_kloop_kernel<false>, smem);
    const long long max_clusters = persistent_sms() / csize;
    const long long grid = (prm.tasks < max_clusters ? prm.tasks : max_clusters) * csize;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
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
    cudaError_t le = hybrid ? cudaLaunchKernelEx(&cfg, cgemm3x_kloop_kernel<true>, map, prm)
                            : cudaLaunchKernelEx(&cfg, cgemm3x_kloop_kernel<false>, map, prm);
    prof_end(st);
    if (le != cudaSuccess) {
        set_last_cuda_error(le, "cudaLaunchKernelEx(cgemm3x_kloop_kernel)");
        return SFA_ERR_CUDA;
    }
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

}  // namespace sfa
