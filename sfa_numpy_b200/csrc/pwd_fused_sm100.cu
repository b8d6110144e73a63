// Stage 3, steering form, in ONE persistent launch per step (sm_100a):
//
//     q_f = A_f X_f,   A_f = Y_L diag(d_f) Y_Q^H = sum_g d_f[g] G_g      for every frequency bin f
//
// i.e. the example idiom  A = matmul(matmul(Y_q, D), conj(Y_p.T)); q = matmul(A, p)
// (doc/examples/modal_beamforming_rigid_spherical_array.py:36-39, radial.diagonal_mode_mat
// micarray/modal/radial.py:269-314) with the [F, L, Q] steering tensor A NEVER written to memory.
//
// Round 1 built A_f into HBM slabs with one kernel and multiplied with another (cgemm3x_sm100.cu):
// 2.7 GB written + ~3 GB read back per step on the headline config, a 34 us window per 74-bin slab in which
// the GEMM could not run, and an output tile pattern (64 look directions x 128 frames per CTA) that HBM
// absorbs at 4.9 TB/s.  Here the roles of the two GEMM operands are swapped:
//
//   * TMEM lanes (UMMA M = 128) = 128 look directions l of one bin.  The steering tile A_f[128 l x K<=64 mics]
//     is the A operand IN TENSOR MEMORY (TS-form tcgen05.mma).  Eight "builder" warps compute it with FP32 FMAs
//     in registers, split it (tf32 hi + lo / packed bf16) and write the four operand planes with tcgen05.st.
//     Spherical-harmonic operators: Legendre recurrence on cos(angle(l, q)) (addition theorem; one float per
//     entry out of L2); circular-harmonic operators: repeated rotation of (cos, sin)(phi_l - phi_q); anything
//     else: the group products G (shared by all bins, L2 resident, a warp's load is 512 contiguous bytes) times
//     the bin's radial filters.  Which one is decided per call by an FP64 check of the given matrices (see the
//     "Addition-theorem build" comment below).  The tile of task i+1 is built while the MMAs of task i run; only
//     the split + tcgen05.st (a few hundred cycles per ~40k cycle task) waits for the operand columns.
//   * UMMA N = 64 frames per step.  The spectra are split into the K-major planes the tensor core reads
//     (Xs[slot][plane][t][q]) by the two spare warps of every CTA, a round of bins AHEAD of the bins being
//     multiplied, into a ring of plane slots that lives in L2 (`convert_ahead`; launches with few bins use the
//     pre-pass kernel `x_split_tail_kernel`), and streamed by TMA (SWIZZLE_128B, mbarrier complete_tx) into a
//     2-stage shared-memory ring.  The two CTAs of a cluster work on two l-tiles of the SAME bin and share that
//     stream: each fetches half of the boxes and multicasts them.
//   * One elected thread issues the tcgen05.mma stream (2 passes hybrid / 3 passes 3xTF32 x 4 real products
//     x K/8 k-steps) into double-buffered (Re, Im) accumulators.
//   * Four epilogue warps read the accumulators (lane = look direction, columns = frames), transpose them
//     through warp-private SWIZZLE_128B staging boxes in shared memory (conflict-free 16-byte stores) and
//     write them with TMA stores.  A CTA owns 128 consecutive rows of out[f] for ALL frames, i.e. over a
//     task it writes one contiguous ~1 MB region of the beam tensor front to back; rows end on whole 128-byte
//     lines when the caller owns the row padding.  The beam power map is accumulated from the same registers.
//
// TMEM: 2 x (64 Re + 64 Im) accumulator columns + 4 x 64 operand columns = 512.
// SMEM: 2 x 64 KB operand stages + 4 x 16 KB store staging + 25 KB conversion tile.
#include <cuda.h>
#include "sfa_common.cuh"
#include "sfa_tcgen05.cuh"

namespace sfa {
namespace fused {

constexpr int kTileL = 128;                         // look directions per task (TMEM lanes)
constexpr int kTileT = 64;                          // frames per MMA step
constexpr int kMaxK = 64;                           // microphones (complex k) resident in TMEM
#ifndef SFA_FUSED_STAGES
#define SFA_FUSED_STAGES 2
#endif
constexpr int kStages = SFA_FUSED_STAGES;
#ifndef SFA_FUSED_PREFETCH
#define SFA_FUSED_PREFETCH 0
#endif
constexpr int kPrefetch = SFA_FUSED_PREFETCH;       // L2 prefetch distance of the spectra stream, in frame tiles
constexpr int kTailMaxRows = 16;                    // L mod 128 up to this many rows goes to the CUDA-core tail kernel
constexpr int kBoxK = 32;                           // fp32 per 128-byte swizzle row
constexpr int kBoxBytes = kTileT * kBoxK * 4;       // 8 KB
constexpr int kStageBytes = 4 * 2 * kBoxBytes;      // 4 planes x 2 k-boxes = 64 KB
// Four warpgroups: {TMA producer, MMA issuer, 2 spare warps}, {4 epilogue warps}, {8 builder warps}.  The roles are
// aligned to warpgroups so that the register file can be re-divided with setmaxnreg: the builders hold a
// 64-value accumulator tile per thread PLUS a deep batch of loads in flight (the group products come from L2 at
// ~1000 cycles under the store traffic; with the 128 registers of a 512-thread CTA the loads serialised and the
// builders, not the tensor pipe, set the pace -- profiles/r2_b_fused_first.txt).
constexpr int kEpiWarp0 = 4, kEpiWarps = 4;
constexpr int kBuildWarp0 = 8, kBuildWarps = 8;
constexpr int kThreads = 512;
constexpr int kRegsCtl = 56, kRegsEpi = 104, kRegsBuild = 176;   // 128 * (56 + 104) + 256 * 176 = 65536
constexpr int kG2Batch = 16;                                     // group-product loads in flight per builder thread
// (Tried: the epilogue warps pulling the WHOLE accumulator tile into registers -- 152 registers each, 152 for the
// builders -- so that the accumulator goes back to the MMA warp before any staging: no gain, 5.37 against 5.21 ms.
// Clusters of four CTAs sharing one bin's plane stream -- half the L2 reads of the planes -- : 6.2 against 5.6 ms.)
// CTA-pair mode (k2: tcgen05.mma.cta_group::2, M = 256 = the two l-tiles of a pair): every CTA holds HALF of the
// frame tile (32 frames) of every plane, so an operand stage is 32 KB instead of 64 KB, the TMA writes and the
// tensor core's B reads of a CTA's shared memory halve, and the freed 64 KB go into a third operand stage and a
// third store slot per epilogue warp.  Bit-equal to the default mode on every shape tested (same MMA sequence per
// output element) -- and SLOWER: 6.2-6.4 ms per step on the headline shape against 4.7-4.8 ms on the same box, whatever
// the split of the freed shared memory (3 stages + 3 slots, 4 + 2, 2 + 4).  One MMA stream for two CTAs locks the
// pair together: the leader cannot start a tile before BOTH epilogues have handed the accumulator back and both
// builders have written their operand, so every hiccup of either store stream stalls both SMs, where the default
// mode couples the two CTAs only through the release of an operand stage.  Kept behind the "fused_pair" option.
#ifndef SFA_FUSED_STAGES2
#define SFA_FUSED_STAGES2 3
#endif
#ifndef SFA_FUSED_RING2
#define SFA_FUSED_RING2 3
#endif
constexpr int kStages2 = SFA_FUSED_STAGES2;
constexpr int kStoreRing2 = SFA_FUSED_RING2;
constexpr int kMaxStages = 4;
static_assert(kStages <= kMaxStages && kStages2 <= kMaxStages, "barrier block");
constexpr int kTmemCols = 512;
constexpr int kAccCols = 2 * kTileT;                // Re + Im per accumulator stage
constexpr int kOperandCol0 = 2 * kAccCols;          // 256
constexpr int kStoreBoxBytes = 32 * 128;            // 32 rows x 16 frames (128 B): one SWIZZLE_128B store box
#ifndef SFA_FUSED_RING
#define SFA_FUSED_RING 2
#endif
constexpr int kStoreRing = SFA_FUSED_RING;          // staging slots (one pass = 2 boxes = 8 KB each) per epilogue warp
// L2 policies (createpolicy + .L2::cache_hint).  When the beam tensor is much larger than L2 its stores carry
// evict_first: written lines are never read again by this kernel, and without the hint they push the
// group-product table and the spectra planes (both re-read by every task of a bin) out of L2 -- measured on the
// headline shape: 4.99 -> 4.64 ms per step.  (evict_last on the table loads alone gives 4.72 ms and adds
// nothing on top; a third staging slot per epilogue warp costs 4 %.)
// Store path: a pure-store kernel with this row-tile pattern moves 4.95 TB/s through tensor-map boxes and 6.1 TB/s
// through one 1-D bulk copy per 512-byte row piece (tools/store_probe3.cu, store_probe4.cu) -- but a bulk copy
// costs its issuing lane ~95 cycles, so the 64 copies per warp and frame tile that the epilogue would need take
// longer than the tile itself (measured in this kernel: 13.0 ms per step against 5.45 ms); one tensor-map store
// moves 32 row pieces per instruction and stays.
// the converters run ahead until the plane ring is full and then wait for a slot most of the time: a slot frees
// every ~20 us, so the poll can be slow
#ifndef SFA_FUSED_RING_SLEEP
#define SFA_FUSED_RING_SLEEP 2000
#endif
#ifndef SFA_FUSED_G2_HINT
#define SFA_FUSED_G2_HINT 0
#endif

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg_hint(const float4* p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* map, const void* src, int c0, int c1, int c2,
                                                  uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
                 : "memory");
}
constexpr int kStoreWarpBytes = kStoreRing * 2 * kStoreBoxBytes;


__device__ __forceinline__ float2 ld_filter(const double2* p) {
    const double2 v = __ldg(p);
    return make_float2((float)v.x, (float)v.y);
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void red_release_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_f32_hint(float* p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

// Addition-theorem build.  When Y_L and Y_Q are spherical-harmonic matrices (angular.sht_matrix; Y_Q may carry real
// or complex quadrature weights) and the filters are per order, the group products collapse to
//     G_n[l, q] = sum_m Y_n^m(l) conj(Y_n^m(q)) w_q = s_q (2n + 1) P_n(x_lq),    s_q = G_0[., q],  x_lq = cos angle(l, q)
// so the steering tile is  A_f[l, q] = s_q sum_n (2n + 1) d_f[n] P_n(x_lq):  the builders run the three-term Legendre
// recurrence on ONE float per entry instead of reading NG complex group products per entry -- 32 KB instead of 590 KB
// per task out of L2 on the headline shape (13 GB per step; with the beam stores in flight that stream, not the
// tensor pipe, set the pace of the kernel: 6.2 ms with it, 4.2 ms with the build switched off, tools/pad_ab.py).
// Nothing is assumed: the preparation kernel computes s, x and the model in FP64 from the given matrices and
// records the largest deviation of every G_n from it next to the largest |G_n|; the builders take this path only
// if every group agrees to 1e-8 (relative to its largest entry), otherwise they read the table as before.
// The same for circular arrays (circular-harmonic matrices, angular.cht_matrix, one filter per column; orders
// [0..N, -N..-1]):  G_c[l, q] = s_q z_lq^m(c),  z_lq = e^{i (phi_l - phi_q)}, so
//     A_f[l, q] = s_q [d_0 + sum_{m=1..N} (d_m + d_{-m}) cos(m D) + i (d_m - d_{-m}) sin(m D)]
// -- the "rotation build": (cos m D, sin m D) by repeated rotation, 8 FP32 operations per order and entry against
// 2 N + 1 complex table entries (BASELINE config 2, 33 columns: 1 MB of table per 3 us task, which made the fused
// kernel slower than the factored form there).  Same FP64 verification, same fallback.
constexpr int kLegMaxGroups = 16;
constexpr int kLegHdrWords = 144;                                                  // 64 dev | 64 mag | model | pad
constexpr size_t kLegHdrBytes = kLegHdrWords * sizeof(unsigned) + 64 * sizeof(float2);      // record | s_q ; then the table
constexpr int kLegTableRow = 32;                                                   // float4 per look direction and l-tile
constexpr float kLegTol = 1e-8f;
enum { kModelNone = 0, kModelLegendre = 1, kModelRotation = 2 };

constexpr int kConvT = 32;                           // frames per conversion tile
constexpr int kConvSmemBytes = kMaxK * (kConvT + 1) * 8 + kTailMaxRows * kMaxK * 8 + kTailMaxRows * 4;
struct Params;
struct Barriers;
template <bool kHybrid>
__device__ void convert_ahead(const Params& prm, Barriers* bars, unsigned char* smem, int cw, int lane,
                              long long cluster_id, long long n_clusters, int crank);

struct Params {
    long long F, L, T;           // bins, look directions handled by this kernel (a multiple of 128 when the tail kernel takes the rest), frames
    long long Lrows;             // rows per bin of the beam tensor / power map (all look directions)
    int K;                       // microphones (<= 64)
    int ksteps, kboxes;          // ceil(K/8), ceil(K/32)
    int ngroups;                 // columns of d (<= NG; any number up to 64 for NG = 0)
    const float4* G2;            // [ltiles][ngroups][Kh][128] : (G[g][l][2j], G[g][l][2j+1]) per lane l
    int Kh;                      // ksteps * 4 column pairs
    const double2* d;            // [F][ngroups] radial filters as the caller holds them (complex128), narrowed on load
    int ltiles, lt_groups, csize, t_tiles;
    long long groups;            // F * lt_groups
    float* power;                // nullable: [F][L] mean_t |q|^2, accumulated in the epilogue (a CTA owns whole rows)
    int store_evict_first;       // beam stores with the L2 evict_first policy (output much larger than L2)
    int no_store;                // power map only: the beams are not written (out == nullptr)
    long long Tst;               // frames per row covered by the store tensor map (T, or T - 1: see odd_last)
    int odd_last;                // odd T without writable row padding: the tensor map ends at frame T - 1 (a TMA store
                                 // clips a box in 16-byte units, so a map of 2 T floats would also write frame T) and
                                 // the epilogue writes the last frame of its rows with plain 8-byte stores
    // stream-ahead conversion of the spectra by the two spare warps of every CTA (conv_D = 0: all done by the pre-pass)
    int conv_D;                  // task on bin f converts one slice of bin f + conv_D; bins < conv_D come from the pre-pass
    int nsl;                     // slices per bin = lt_groups * csize (one per CTA-task of a bin)
    int Q, Kp, rem;              // microphones, padded k extent of the planes, tail look directions (L mod 128 <= 16)
    const float2* X;             // [F][Q][T] raw spectra
    float* Xs;                   // [F][4][T][Kp] split planes
    const float4* G2_last;       // group products of the tail rows
    float2* out;                 // beam tensor (tail rows are written with plain stores)
    long long ldo, Lm;           // row pitch of out, first tail row
    int* ready;                  // [F] slices of the bin converted so far
    int* done;                   // [F] CTA-tasks of the bin that have received all their planes
    int ring;                    // plane slots: bin f lives in slot f % ring (conv_D > 0); = F otherwise
    float* tail_part;            // [F][nsl][16] partial power sums of the tail rows
    // Addition-theorem build (spherical-harmonic operators, see legendre block below): nullable
    const unsigned* leg_devmag;  // [64] largest deviation of G_n from the model, [64] largest |G_n| (float bits), [1] model
    const float2* leg_sq;        // [64] s_q = G_0[., q]
    const float4* leg_xt;        // [ltiles][32][128] per lane l: Legendre x[l][4 j .. 4 j + 3] (x = cos of the angle
                                 // (l, q); the first 16 of a row), rotation (cos, sin)[l][2 j .. 2 j + 1]
    int debug;                   // timing attribution (SFA_DEBUG_HOOKS builds only): 1 no TMA stores, 2 no build FMAs,
                                 // 4 no staging writes, 8 no MMAs, 16 no operand loads, 32 converters read no X,
                                 // 64 converters write no planes
};

struct __align__(8) Barriers {
    uint64_t b_full[kMaxStages];
    uint64_t b_empty[kMaxStages];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint64_t a_full;
    uint64_t a_empty;
    uint32_t tmem_base;
    uint32_t pad;
};

template <bool kHybrid, int NG, bool k2>
__global__ void __launch_bounds__(kThreads, 1)
pwd_fused_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap o_map,
                 const Params prm) {
    constexpr int kSt = k2 ? kStages2 : kStages;                         // operand stages
    constexpr int kBoxB = k2 ? kBoxBytes / 2 : kBoxBytes;                // one (plane, k-box) of a stage: 32 / 64 frames
    constexpr int kStageB = 8 * kBoxB;
    constexpr int kRing = k2 ? kStoreRing2 : kStoreRing;                 // store slots per epilogue warp
    constexpr int kStoreWarpB = kRing * 2 * kStoreBoxBytes;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* stage_base = smem;                                   // kSt x 64 KB (32 KB in pair mode)
    unsigned char* store_base = smem + (size_t)kSt * kStageB;           // 4 x 16 KB (24 KB)
    Barriers* bars = reinterpret_cast<Barriers*>(store_base + kEpiWarps * kStoreWarpB);
    unsigned char* conv_base = reinterpret_cast<unsigned char*>(bars) + 256;      // converter warps' tile (conv_D > 0 only)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int csize = prm.csize;
    const uint32_t crank = (csize > 1) ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
    const long long cluster_id = blockIdx.x / csize, n_clusters = gridDim.x / csize;

    if (threadIdx.x == 0) {
        // pair mode: ONE MMA issuer (the leader, rank 0) for both CTAs.  Its barriers collect the arrivals of both
        // CTAs (builders: a_full, epilogue warps: acc_empty, TMA bytes of both halves: b_full); its commits are
        // multicast to the barriers of both (b_empty, acc_full, a_empty)
        for (int s = 0; s < kSt; ++s) {
            mbar_init(&bars->b_full[s], 1);
            mbar_init(&bars->b_empty[s], k2 ? 1 : csize);    // the MMA issuer of every CTA in the cluster
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->acc_full[a], 1);
            mbar_init(&bars->acc_empty[a], k2 ? 2 * kEpiWarps : kEpiWarps);
        }
        mbar_init(&bars->a_full, k2 ? 2 * kBuildWarps : kBuildWarps);
        mbar_init(&bars->a_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (k2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(&bars->tmem_base)),
                         "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(&bars->tmem_base)),
                         "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem = bars->tmem_base;

    // group -> (bin f, l-tile of this CTA); a "ghost" l-tile beyond ltiles only keeps the multicast protocol alive
    auto decode = [&](long long grp, long long& f, int& lt) {
        f = grp / prm.lt_groups;
        lt = (int)(grp - f * prm.lt_groups) * csize + (int)crank;
    };

    // setmaxnreg sits INSIDE each role's branch so that it dominates that role's code only (ptxas sizes the
    // register allocation of a region by the setmaxnreg that dominates it)
    if (warp < kEpiWarp0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
    if (warp == 0) {
        // ===================== TMA producer: stream the pre-split spectra of the bin =====================
        if (elect_one_sync()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&x_map) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            for (long long grp = cluster_id; grp < prm.groups; grp += n_clusters) {
                const int f = (int)(grp / prm.lt_groups);
                const int fslot = f % prm.ring;
                if (prm.conv_D > 0) {
                    // planes of bin f are written by the converter warps of the CTAs working conv_D bins back
                    // (the first conv_D bins by the whole grid in its prologue)
                    while (ld_acquire_gpu(prm.ready + f) < prm.nsl) __nanosleep(64);
                    asm volatile("fence.proxy.async;" ::: "memory");          // generic-proxy writes -> TMA reads
                }
                for (int tt = 0; tt < prm.t_tiles; ++tt) {
                    mbar_wait(&bars->b_empty[stage], phase ^ 1);
                    unsigned char* dst = stage_base + (size_t)stage * kStageB;
                    if (prm.debug & 16) {
                        if (!k2 || crank == 0) mbar_arrive(&bars->b_full[stage]);
                        if (++stage == kSt) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    // pair mode: the leader's barrier expects the bytes of BOTH halves (the peer's loads complete
                    // on it; they may land before this arrive -- the phase cannot end without it)
                    if (!k2) mbar_arrive_expect_tx(&bars->b_full[stage], (uint32_t)(4 * prm.kboxes * kBoxB));
                    else if (crank == 0) mbar_arrive_expect_tx(&bars->b_full[stage], (uint32_t)(2 * 4 * prm.kboxes * kBoxB));
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl) {
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
                            if (kb >= prm.kboxes) continue;
                            unsigned char* dbox = dst + (pl * 2 + kb) * kBoxB;
                            if (k2) {
                                tma_load_3d_2sm(&x_map, &bars->b_full[stage], dbox, kb * kBoxK,
                                                tt * kTileT + (int)crank * (kTileT / 2), fslot * 4 + pl);
                            } else if (csize == 1) {
                                tma_load_3d(&x_map, &bars->b_full[stage], dbox, kb * kBoxK, tt * kTileT, fslot * 4 + pl);
                            } else if (((prm.kboxes > 1 ? pl * 2 + kb : pl) % csize) == (int)crank) {
                                tma_load_3d_mc(&x_map, &bars->b_full[stage], dbox, kb * kBoxK, tt * kTileT, fslot * 4 + pl, cmask);
                            }
                        }
                    }
                    if (++stage == kSt) {
                        stage = 0;
                        phase ^= 1;
                    }
                    // optional L2 prefetch of the tile kPrefetch steps ahead.  Measured on the headline shape: 2 or 4
                    // tiles ahead cost 3-5 % (the prefetch requests compete with the store stream), so it is off.
                    if (kPrefetch > 0) {
                        int ptt = tt + kPrefetch;
                        long long pgrp = grp;
                        while (ptt >= prm.t_tiles) {
                            ptt -= prm.t_tiles;
                            pgrp += n_clusters;
                        }
                        if (pgrp < prm.groups) {
                            const int pf = (int)(pgrp / prm.lt_groups);
#pragma unroll
                            for (int pl = 0; pl < 4; ++pl) {
#pragma unroll
                                for (int kb = 0; kb < 2; ++kb) {
                                    if (kb >= prm.kboxes) continue;
                                    if (csize == 1 || ((prm.kboxes > 1 ? pl * 2 + kb : pl) % csize) == (int)crank)
                                        tma_prefetch_3d(&x_map, kb * kBoxK, ptt * kTileT, (pf % prm.ring) * 4 + pl);
                                }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const int t_tail = (int)(prm.T - (long long)(prm.t_tiles - 1) * kTileT);
        const int t_tail16 = ((t_tail + 15) >> 4) << 4;
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0, a_phase = 0;
        // pair mode: the leader issues for both CTAs (M = 256: its own l-tile and the peer's, each in its own tensor
        // memory; N = 64 frames, 32 from each CTA's shared memory); the peer's warp 1 has nothing to do
        for (long long grp = cluster_id; grp < prm.groups && !(k2 && crank != 0); grp += n_clusters) {
            long long f;
            int lt;
            decode(grp, f, lt);
            if (!k2 && lt >= prm.ltiles) {
                // ghost tile: consume the multicast stream, compute nothing
                for (int tt = 0; tt < prm.t_tiles; ++tt) {
                    mbar_wait(&bars->b_full[stage], phase);
                    if (elect_one_sync()) tcgen05_commit_mc(&bars->b_empty[stage], cmask);
                    __syncwarp();
                    if (++stage == kSt) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (prm.conv_D > 0 && lane == 0) red_release_add(prm.done + f, 1);
                continue;
            }
            mbar_wait(&bars->a_full, a_phase);
            for (int tt = 0; tt < prm.t_tiles; ++tt) {
                mbar_wait(&bars->acc_empty[acc], acc_phase ^ 1);
                mbar_wait(&bars->b_full[stage], phase);
                tcgen05_fence_after();
                // (pair mode: always all 64 frames -- the halves of the two CTAs are fixed at 32 frames each; the
                // frames beyond T are zeros)
                const int n_mma = (tt == prm.t_tiles - 1 && !k2) ? t_tail16 : kTileT;
                constexpr int kM = k2 ? 2 * kTileL : kTileL;
                const uint32_t idesc = make_idesc(kM, n_mma, false);
                const uint32_t idesc_neg = make_idesc(kM, n_mma, true);
                const uint32_t idesc_bf = make_idesc(kM, n_mma, false, true);
                const uint32_t idesc_bf_neg = make_idesc(kM, n_mma, true, true);
                if (elect_one_sync()) {
                    const uint32_t sbase = smem_u32(stage_base + (size_t)stage * kStageB);
                    const uint64_t bdesc = make_b_desc(sbase);
                    const uint32_t d_re = tmem + acc * kAccCols;
                    const uint32_t d_im = d_re + kTileT;
                    const uint32_t a0 = tmem + kOperandCol0;
                    constexpr int kPasses = kHybrid ? 2 : 3;
#pragma unroll
                    for (int pass = 0; pass < kPasses; ++pass) {
                        // planes {re_hi, im_hi, re_x, im_x}; x = lo (fp32) or packed bf16 pairs.
                        // 3xTF32: (A, B) = (hi, hi) (hi, lo) (lo, hi).  hybrid: tf32 (hi, hi), then bf16 (x, x).
                        const int ap = kHybrid ? (pass == 1 ? 2 : 0) : ((pass == 2) ? 2 : 0);
                        const int bp = kHybrid ? (pass == 1 ? 2 : 0) : ((pass == 1) ? 2 : 0);
                        const bool bf = kHybrid && pass == 1;
#pragma unroll
                        for (int ks = 0; ks < kMaxK / 8; ++ks) {
                            if (ks < prm.ksteps && !(prm.debug & 8)) {
#ifdef SFA_DEBUG_HOOKS
                                // timing probe (results are garbage): the hi*hi pass as a 16-bit-kind pass with half the
                                // k-steps, i.e. the instruction stream an fp16-split variant would issue
                                if ((prm.debug & 256) && pass == 0 && (ks & 1)) continue;
#endif
                                const int kb = ks >> 2, kk = ks & 3;
                                const uint64_t b_re = bdesc + (uint64_t)((((bp + 0) * 2 + kb) * kBoxB + kk * 32) >> 4);
                                const uint64_t b_im = bdesc + (uint64_t)((((bp + 1) * 2 + kb) * kBoxB + kk * 32) >> 4);
                                const uint32_t a_re = a0 + (ap + 0) * kMaxK + ks * 8;
                                const uint32_t a_im = a0 + (ap + 1) * kMaxK + ks * 8;
                                const uint32_t accum = (pass == 0 && ks == 0) ? 0u : 1u;
                                if (k2) {
                                    if (bf) {
                                        umma_bf16_ts2(d_re, a_re, b_re, idesc_bf, accum);
                                        umma_bf16_ts2(d_re, a_im, b_im, idesc_bf_neg, 1u);
                                        umma_bf16_ts2(d_im, a_re, b_im, idesc_bf, accum);
                                        umma_bf16_ts2(d_im, a_im, b_re, idesc_bf, 1u);
                                    } else {
                                        umma_tf32_ts2(d_re, a_re, b_re, idesc, accum);
                                        umma_tf32_ts2(d_re, a_im, b_im, idesc_neg, 1u);
                                        umma_tf32_ts2(d_im, a_re, b_im, idesc, accum);
                                        umma_tf32_ts2(d_im, a_im, b_re, idesc, 1u);
                                    }
                                } else if (bf || ((prm.debug & 256) && pass == 0)) {
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
                    if (k2) {
                        tcgen05_commit_mc2(&bars->b_empty[stage], 3);
                        tcgen05_commit_mc2(&bars->acc_full[acc], 3);
                        if (tt == prm.t_tiles - 1) {
                            tcgen05_commit_mc2(&bars->a_empty, 3);
                            if (prm.conv_D > 0) red_release_add(prm.done + f, 2);       // both halves have landed
                        }
                    } else {
                    if (csize == 1) tcgen05_commit(&bars->b_empty[stage]);
                    else tcgen05_commit_mc(&bars->b_empty[stage], cmask);
                    tcgen05_commit(&bars->acc_full[acc]);
                    if (tt == prm.t_tiles - 1) {
                        tcgen05_commit(&bars->a_empty);
                        // every plane box of this task has landed in shared memory (b_full of the last tile was
                        // waited for above): the slot of bin f may be overwritten once all nsl tasks have said so
                        if (prm.conv_D > 0) red_release_add(prm.done + f, 1);
                    }
                    }
                }
                __syncwarp();
                if (++stage == kSt) {
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
    } else if (prm.conv_D > 0) {
        // ===================== converters (warps 2, 3): split the spectra of a bin conv_D bins ahead =====================
        convert_ahead<kHybrid>(prm, bars, conv_base, warp - 2, lane, cluster_id, n_clusters, (int)crank);
    }
    } else if (warp >= kBuildWarp0) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsBuild));
        // ===================== builders: A_f tile = sum_g d_f[g] G_g  ->  registers -> TMEM =====================
        // warp (quarter, half): rows quarter*32 + lane of the l-tile, mic columns [half*32, half*32 + 32)
        const int quarter = warp & 3;
        const int half = (warp - kBuildWarp0) >> 2;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const int row = quarter * 32 + lane;
        // column pairs this warp owns that actually exist (Kh = ksteps * 4 pairs; pairs beyond it stay zero)
        int npairs = prm.Kh - half * 16;
        npairs = npairs < 0 ? 0 : (npairs > 16 ? 16 : npairs);
        uint32_t a_phase = 0;
#if SFA_FUSED_G2_HINT
        const uint64_t g2_pol = l2_policy_evict_last();
#endif
        // addition-theorem build: only if the preparation found every group product to follow the Legendre model
        bool use_leg = false, use_rot = false;
        if (prm.leg_xt != nullptr && prm.ngroups >= 2) {
            const unsigned model = __ldg(prm.leg_devmag + 128);
            bool ok = true;
            for (int g = 0; g < prm.ngroups; ++g) {
                const float dev = __uint_as_float(__ldg(prm.leg_devmag + g));
                const float mag = __uint_as_float(__ldg(prm.leg_devmag + 64 + g));
                if (!(dev <= kLegTol * mag)) ok = false;
            }
            use_leg = ok && model == kModelLegendre && NG > 0 && NG <= kLegMaxGroups;
            use_rot = ok && model == kModelRotation && (prm.ngroups & 1);
        }
        for (long long grp = cluster_id; grp < prm.groups; grp += n_clusters) {
            long long f;
            int lt;
            decode(grp, f, lt);
            if (lt >= prm.ltiles) {
                if (k2) {
                    // ghost tile of a pair: the leader's MMAs cover these rows as well (nobody reads the result);
                    // keep the operand hand-over alive
                    mbar_wait(&bars->a_empty, a_phase ^ 1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta(&bars->a_full, 0);
                    a_phase ^= 1;
                }
                continue;
            }
            float are[32], aim[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) are[j] = aim[j] = 0.f;
            const float4* gp = prm.G2 + ((size_t)lt * prm.ngroups * prm.Kh + (size_t)half * 16) * kTileL + row;
            // one group: are/aim += d_f[g] * G_g[row, this warp's 32 mic columns]
            auto add_group = [&](const float2 dgv, const float4* gg) {
                // all loads of a batch are issued before its first FMA (memory-level parallelism)
#pragma unroll
                for (int j0 = 0; j0 < 16; j0 += kG2Batch) {
                    float4 v[kG2Batch];
#pragma unroll
                    for (int jj = 0; jj < kG2Batch; ++jj) {
                        const int j = j0 + jj;
#if SFA_FUSED_G2_HINT
                        v[jj] = (j < npairs) ? ldg_hint(gg + (size_t)j * kTileL, g2_pol) : make_float4(0.f, 0.f, 0.f, 0.f);
#else
                        v[jj] = (j < npairs) ? __ldg(gg + (size_t)j * kTileL) : make_float4(0.f, 0.f, 0.f, 0.f);
#endif
                    }
#pragma unroll
                    for (int jj = 0; jj < kG2Batch; ++jj) {
                        const int j = j0 + jj;
                        are[2 * j] = fmaf(dgv.x, v[jj].x, are[2 * j]);
                        are[2 * j] = fmaf(-dgv.y, v[jj].y, are[2 * j]);
                        aim[2 * j] = fmaf(dgv.x, v[jj].y, aim[2 * j]);
                        aim[2 * j] = fmaf(dgv.y, v[jj].x, aim[2 * j]);
                        are[2 * j + 1] = fmaf(dgv.x, v[jj].z, are[2 * j + 1]);
                        are[2 * j + 1] = fmaf(-dgv.y, v[jj].w, are[2 * j + 1]);
                        aim[2 * j + 1] = fmaf(dgv.x, v[jj].w, aim[2 * j + 1]);
                        aim[2 * j + 1] = fmaf(dgv.y, v[jj].z, aim[2 * j + 1]);
                    }
                }
            };
            if (use_rot) {
                // rotation build (circular-harmonic operators): columns [0..N, -N..-1]
                if (!(prm.debug & 2)) {
                    const int ng = prm.ngroups, Nn = (ng - 1) >> 1;
                    const double2* dF = prm.d + f * ng;
                    const float2 d0 = ld_filter(dF);
                    const float4* tr = prm.leg_xt + ((size_t)lt * kLegTableRow + (size_t)half * 16) * kTileL + row;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (half * 32 + h * 16 >= 2 * prm.Kh) continue;          // columns beyond the padded K: zero
                        float c[16], s[16], C[16], S[16];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 v = __ldg(tr + (size_t)(h * 8 + i) * kTileL);
                            c[2 * i] = v.x;
                            s[2 * i] = v.y;
                            c[2 * i + 1] = v.z;
                            s[2 * i + 1] = v.w;
                        }
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            C[e] = c[e];
                            S[e] = s[e];
                            are[16 * h + e] = d0.x;
                            aim[16 * h + e] = d0.y;
                        }
#pragma unroll 1
                        for (int m = 1; m <= Nn; ++m) {
                            const float2 dp = ld_filter(dF + m), dm = ld_filter(dF + ng - m);
                            const float ex = dp.x + dm.x, ey = dp.y + dm.y;      // d_m + d_-m
                            const float ox = dm.y - dp.y, oy = dp.x - dm.x;      // i (d_m - d_-m)
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                are[16 * h + e] = fmaf(ex, C[e], fmaf(ox, S[e], are[16 * h + e]));
                                aim[16 * h + e] = fmaf(ey, C[e], fmaf(oy, S[e], aim[16 * h + e]));
                                const float cn = fmaf(C[e], c[e], -S[e] * s[e]);
                                S[e] = fmaf(S[e], c[e], C[e] * s[e]);
                                C[e] = cn;
                            }
                        }
                    }
                    const float2* sq = prm.leg_sq + half * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float2 sv = __ldg(sq + j);
                        const float r = are[j], i = aim[j];
                        are[j] = fmaf(r, sv.x, -i * sv.y);
                        aim[j] = fmaf(r, sv.y, i * sv.x);
                    }
                }
            } else if (NG > 0 && NG <= kLegMaxGroups && use_leg) {
                // A[l, q] = s_q sum_n (2n + 1) d[n] P_n(x_lq): Legendre recurrence in registers, 16 entries at a time
                constexpr int NGA = (NG > 0 && NG <= kLegMaxGroups) ? NG : 2;
                float2 dg[NGA];
#pragma unroll
                for (int g = 0; g < NGA; ++g) {
                    dg[g] = (g < prm.ngroups) ? ld_filter(prm.d + f * prm.ngroups + g) : make_float2(0.f, 0.f);
                    dg[g].x *= (float)(2 * g + 1);
                    dg[g].y *= (float)(2 * g + 1);
                }
                const float4* xr = prm.leg_xt + ((size_t)lt * kLegTableRow + (size_t)half * 8) * kTileL + row;
                if (!(prm.debug & 2)) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float x[16], p0[16], p1[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 v = __ldg(xr + (size_t)(h * 4 + i) * kTileL);
                            x[4 * i] = v.x;
                            x[4 * i + 1] = v.y;
                            x[4 * i + 2] = v.z;
                            x[4 * i + 3] = v.w;
                        }
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            p0[e] = 1.f;
                            p1[e] = x[e];
                            are[16 * h + e] = fmaf(dg[1].x, x[e], dg[0].x);
                            aim[16 * h + e] = fmaf(dg[1].y, x[e], dg[0].y);
                        }
#pragma unroll
                        for (int n = 1; n + 1 < NGA; ++n) {
                            if (n + 1 < prm.ngroups) {
                                const float ca = (float)(2 * n + 1) / (float)(n + 1), cb = (float)n / (float)(n + 1);
#pragma unroll
                                for (int e = 0; e < 16; ++e) {
                                    const float pn = fmaf(ca * x[e], p1[e], -cb * p0[e]);
                                    p0[e] = p1[e];
                                    p1[e] = pn;
                                    are[16 * h + e] = fmaf(dg[n + 1].x, pn, are[16 * h + e]);
                                    aim[16 * h + e] = fmaf(dg[n + 1].y, pn, aim[16 * h + e]);
                                }
                            }
                        }
                    }
                    const float2* sq = prm.leg_sq + half * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float2 s = __ldg(sq + j);
                        const float r = are[j], i = aim[j];
                        are[j] = fmaf(r, s.x, -i * s.y);
                        aim[j] = fmaf(r, s.y, i * s.x);
                    }
                }
            } else if (NG > 0) {
                // few groups (spherical arrays: one per order): the bin's filters sit in registers, loop unrolled
                constexpr int NGA = NG > 0 ? NG : 1;
                float2 dg[NGA];
#pragma unroll
                for (int g = 0; g < NGA; ++g)
                    dg[g] = (g < prm.ngroups) ? ld_filter(prm.d + f * prm.ngroups + g) : make_float2(0.f, 0.f);
#pragma unroll
                for (int g = 0; g < NGA; ++g)
                    if (g < prm.ngroups && !(prm.debug & 2)) add_group(dg[g], gp + (size_t)g * prm.Kh * kTileL);
            } else {
                // many groups (circular arrays: one per column, up to 64): plain loop
#pragma unroll 1
                for (int g = 0; g < prm.ngroups; ++g)
                    add_group(ld_filter(prm.d + f * prm.ngroups + g), gp + (size_t)g * prm.Kh * kTileL);
            }
            // the MMAs of the previous task must have finished reading the operand columns
            mbar_wait(&bars->a_empty, a_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t col = tmem + lane_base + kOperandCol0 + half * 32;
            uint32_t w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = to_tf32(are[j]);
            SFA_TMEM_ST32(col + 0 * kMaxK, w);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float h = __uint_as_float(to_tf32(are[j]));
                w[j] = kHybrid ? pack_bf16(h, are[j] - h) : __float_as_uint(are[j] - h);     // A side: (hi, lo)
            }
            SFA_TMEM_ST32(col + 2 * kMaxK, w);
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = to_tf32(aim[j]);
            SFA_TMEM_ST32(col + 1 * kMaxK, w);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float h = __uint_as_float(to_tf32(aim[j]));
                w[j] = kHybrid ? pack_bf16(h, aim[j] - h) : __float_as_uint(aim[j] - h);
            }
            SFA_TMEM_ST32(col + 3 * kMaxK, w);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (k2) mbar_arrive_cta(&bars->a_full, 0);          // the leader's barrier collects both CTAs' builders
                else mbar_arrive(&bars->a_full);
            }
            a_phase ^= 1;
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
        // ===================== epilogue: TMEM -> regs -> swizzled smem boxes -> TMA store =====================
        // warp owns TMEM lane quarter (warp & 3) = 32 look directions; per frame tile two passes of 32 frames.
        const int quarter = warp & 3;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        unsigned char* const my_store = store_base + (size_t)(warp - kEpiWarp0) * kStoreWarpB;
        // SWIZZLE_128B: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
        const uint32_t row_off = (uint32_t)lane * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
        int acc = 0, buf = 0;
        uint32_t acc_phase = 0;
        const uint64_t st_pol = l2_policy_evict_first();
        const bool st_hint = prm.store_evict_first != 0;
        for (long long grp = cluster_id; grp < prm.groups; grp += n_clusters) {
            long long f;
            int lt;
            decode(grp, f, lt);
            if (lt >= prm.ltiles) {
                if (k2) {
                    // ghost tile of a pair: hand every accumulator straight back
                    for (int tt = 0; tt < prm.t_tiles; ++tt) {
                        mbar_wait(&bars->acc_full[acc], acc_phase);
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cta(&bars->acc_empty[acc], 0);
                        if (++acc == 2) {
                            acc = 0;
                            acc_phase ^= 1;
                        }
                    }
                }
                continue;
            }
            const int l0 = lt * kTileL + quarter * 32;
            float psum = 0.f;                                    // sum_t |q[f, l0 + lane, t]|^2
            for (int tt = 0; tt < prm.t_tiles; ++tt) {
                mbar_wait(&bars->acc_full[acc], acc_phase);
                tcgen05_fence_after();
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    const uint32_t t_re = tmem + lane_base + acc * kAccCols + pass * 32;
                    uint32_t re[32], im[32];
                    SFA_TMEM_LD32(t_re, re);
                    SFA_TMEM_LD32(t_re + kTileT, im);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (pass == 1) {
                        // the whole accumulator of this warp's lanes is in registers: release it
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (k2) mbar_arrive_cta(&bars->acc_empty[acc], 0);
                            else mbar_arrive(&bars->acc_empty[acc]);
                        }
                    }
                    const long long t0 = (long long)tt * kTileT + pass * 32;
                    if (t0 >= prm.T || l0 >= prm.L) continue;                   // warp-uniform
                    if (prm.power) {
                        const int nval = (prm.T - t0 >= 32) ? 32 : (int)(prm.T - t0);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (j < nval) {
                                const float xr = __uint_as_float(re[j]), xi = __uint_as_float(im[j]);
                                psum = fmaf(xr, xr, fmaf(xi, xi, psum));
                            }
                        }
                    }
                    if (prm.no_store) continue;                                 // power map only (warp-uniform)
                    if (prm.odd_last && prm.T - 1 >= t0 && prm.T - 1 < t0 + 32) {
                        const int jj = (int)(prm.T - 1 - t0);
                        uint32_t xr = 0u, xi = 0u;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j == jj) {
                                xr = re[j];
                                xi = im[j];
                            }
                        if (l0 + lane < prm.L)
                            prm.out[(f * prm.Lrows + l0 + lane) * prm.ldo + prm.T - 1] =
                                make_float2(__uint_as_float(xr), __uint_as_float(xi));
                        if (t0 >= prm.Tst) continue;                            // the tile held that frame only
                    }
                    unsigned char* const sbuf = my_store + (size_t)buf * 2 * kStoreBoxBytes;
                    // the stores issued from this slot kStoreRing passes ago must have finished reading it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kRing - 1) : "memory");
                    __syncwarp();
                    if (!(prm.debug & 4))
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        unsigned char* const box = sbuf + b * kStoreBoxBytes + row_off;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const int j = 16 * b + 2 * c;
                            *reinterpret_cast<float4*>(box + (((uint32_t)c ^ sw) << 4)) =
                                make_float4(__uint_as_float(re[j]), __uint_as_float(im[j]),
                                            __uint_as_float(re[j + 1]), __uint_as_float(im[j + 1]));
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && !(prm.debug & 1)) {
                        if (st_hint) {
                            tma_store_3d_hint(&o_map, sbuf, (int)(2 * t0), l0, (int)f, st_pol);
                            if (t0 + 16 < prm.Tst)
                                tma_store_3d_hint(&o_map, sbuf + kStoreBoxBytes, (int)(2 * (t0 + 16)), l0, (int)f, st_pol);
                        } else {
                            tma_store_3d(&o_map, sbuf, (int)(2 * t0), l0, (int)f, false);
                            if (t0 + 16 < prm.Tst)
                                tma_store_3d(&o_map, sbuf + kStoreBoxBytes, (int)(2 * (t0 + 16)), l0, (int)f, false);
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (++buf == kRing) buf = 0;
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
            if (prm.power && l0 + lane < prm.L) prm.power[f * prm.Lrows + l0 + lane] = psum / (float)prm.T;
        }
    }

    // outstanding bulk stores of this thread (epilogue issuers) must complete before the CTA exits
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tcgen05_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        if (k2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
    }
}

// Stream-ahead conversion (warps 2 and 3 of every CTA, 64 threads, <= 56 registers).
// `convert_slice` turns slice j of bin f' -- the 32-frame tiles j, j + nsl, ... of X[f'] -- into the four K-major
// planes the TMA producer streams (same arithmetic as x_split_tail_kernel): the tile comes in with 8-byte
// cp.async copies (all 32 per thread in flight at once: no registers, one DRAM latency per tile), is read back
// transposed and written along q.  The tail look directions of those frames are computed from the tile that
// is in shared memory anyway.  Schedule:
//   * prologue: the grid converts the first conv_D bins, unit (bin, slice) = blockIdx.x, + gridDim.x, ...;
//   * main loop: the CTA-task (bin f, slice j) converts slice j of bin f + conv_D.  The planes live in a RING of
//     `ring` bin slots (bin f in slot f mod ring, ~40 MB, L2-resident): a slot is overwritten while its lines are
//     still dirty in L2, so the planes are never written back -- DRAM sees X once and the beams once -- and the
//     pre-pass over all bins (0.38 ms of a 5.4 ms step) disappears.  The converters run ahead of their CTA as far
//     as the ring allows: before writing slot s they wait until every task of the bin that held it has received
//     its planes (done[f' - ring] == nsl, signalled by the MMA warps).
// Hand-over: every thread fences its stores, the 64 threads meet on named barrier 1, one thread adds 1 to
// ready[f'] (release); the TMA producer of a task on bin f' spins on ready[f'] == nsl (acquire) and issues a
// proxy fence before its first load.  The converters never wait for anything but their own CTA reaching the
// task, so the dependency chain runs strictly towards lower bins and ends at the prologue.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

template <bool kHybrid>
__device__ __forceinline__ void convert_slice(const Params& prm, unsigned char* smem, int cw, int lane, long long fc,
                                              int slice) {
    float2 (*tile)[kConvT + 1] = reinterpret_cast<float2 (*)[kConvT + 1]>(smem);
    float2 (*sA)[kMaxK] = reinterpret_cast<float2 (*)[kMaxK]>(smem + kMaxK * (kConvT + 1) * 8);
    float* spow = reinterpret_cast<float*>(smem + kMaxK * (kConvT + 1) * 8 + kTailMaxRows * kMaxK * 8);
    const int tid = cw * 32 + lane;
    const long long T = prm.T;
    const int Q = prm.Q, Kp = prm.Kp, rem = prm.rem;
    const int n32 = (int)((T + kConvT - 1) / kConvT);
    const long long plane = T * Kp;
    const float2* const xf = prm.X + fc * Q * T;
    float* const base = prm.Xs + (fc % prm.ring) * 4 * plane;
    const uint64_t keep = l2_policy_evict_last();
    if (fc >= prm.ring) {
        // the slot still holds bin fc - ring until every task of that bin has received its planes (two rounds of
        // tasks ago in the normal course of events, so this does not spin)
        if (tid == 0)
            while (ld_acquire_gpu(prm.done + (fc - prm.ring)) < prm.nsl) __nanosleep(SFA_FUSED_RING_SLEEP);
        named_bar_sync(1, 64);
    }
    if (rem > 0) {
        for (int i = tid; i < rem * prm.Kh; i += 64) {
            const int r = i / prm.Kh, j = i - r * prm.Kh;
            float ar0 = 0.f, ai0 = 0.f, ar1 = 0.f, ai1 = 0.f;
            for (int g = 0; g < prm.ngroups; ++g) {
                const float2 dg = ld_filter(prm.d + fc * prm.ngroups + g);
                const float4 v = __ldg(prm.G2_last + ((size_t)g * prm.Kh + j) * kTileL + r);
                ar0 = fmaf(dg.x, v.x, ar0);
                ar0 = fmaf(-dg.y, v.y, ar0);
                ai0 = fmaf(dg.x, v.y, ai0);
                ai0 = fmaf(dg.y, v.x, ai0);
                ar1 = fmaf(dg.x, v.z, ar1);
                ar1 = fmaf(-dg.y, v.w, ar1);
                ai1 = fmaf(dg.x, v.w, ai1);
                ai1 = fmaf(dg.y, v.z, ai1);
            }
            sA[r][2 * j] = make_float2(ar0, ai0);
            sA[r][2 * j + 1] = make_float2(ar1, ai1);
        }
        for (int r = cw + 2 * lane; r < rem; r += 64) spow[r] = 0.f;      // row r belongs to warp r & 1
        __syncwarp();
    }
    for (int tile_no = slice; tile_no < n32; tile_no += prm.nsl) {
        const long long t0 = (long long)tile_no * kConvT;
        named_bar_sync(1, 64);                                 // previous tile consumed, sA complete
        // load: warp cw takes the microphones [32 cw, 32 cw + 32), lane = frame (coalesced along t)
        {
            const long long t = t0 + lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int q = cw * 32 + i;
                if (q < Q && t < T && !(prm.debug & 32)) cp_async8(&tile[q][lane], xf + (long long)q * T + t);
                else tile[q][lane] = make_float2(0.f, 0.f);
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        named_bar_sync(1, 64);
        // planes: lane = microphone within the half, loop over the frames of the tile (writes along q)
        {
            const int q = cw * 32 + lane;
            if (q < Kp && !(prm.debug & 64)) {
#pragma unroll 4
                for (int i = 0; i < kConvT; ++i) {
                    const long long t = t0 + i;
                    if (t < T) {
                        const float2 v = tile[q][i];
                        const float hr = __uint_as_float(to_tf32(v.x)), hi = __uint_as_float(to_tf32(v.y));
                        float xr = v.x - hr, xi = v.y - hi;
                        if (kHybrid) {                          // B side: (lo, hi) = (even, odd) bf16 k
                            xr = __uint_as_float(pack_bf16(xr, hr));
                            xi = __uint_as_float(pack_bf16(xi, hi));
                        }
                        float* o = base + t * Kp + q;
#ifdef SFA_DEBUG_HOOKS
                        if (prm.debug & 1024) {                 // plane stores without the evict_last policy
                            o[0] = hr;
                            o[plane] = hi;
                            o[2 * plane] = xr;
                            o[3 * plane] = xi;
                            continue;
                        }
#endif
                        st_f32_hint(o, hr, keep);
                        st_f32_hint(o + plane, hi, keep);
                        st_f32_hint(o + 2 * plane, xr, keep);
                        st_f32_hint(o + 3 * plane, xi, keep);
                    }
                }
            }
        }
        // tail look directions: warp cw takes the rows cw, cw + 2, ...; lane = frame
        for (int r = cw; r < rem; r += 2) {
            const long long t = t0 + lane;
            float ar = 0.f, ai = 0.f;
            for (int q = 0; q < Q; ++q) {
                const float2 a = sA[r][q];
                const float2 x = tile[q][lane];
                ar = fmaf(a.x, x.x, ar);
                ar = fmaf(-a.y, x.y, ar);
                ai = fmaf(a.x, x.y, ai);
                ai = fmaf(a.y, x.x, ai);
            }
            float p = 0.f;
            if (t < T) {
                if (!prm.no_store) prm.out[(fc * prm.Lrows + prm.Lm + r) * prm.ldo + t] = make_float2(ar, ai);
                p = ar * ar + ai * ai;
            }
            if (prm.power) {
                for (int off = 16; off; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
                if (lane == 0) spow[r] += p;
            }
        }
    }
    if (prm.power && rem > 0) {
        __syncwarp();
        for (int r = cw + 2 * lane; r < rem; r += 64)
            prm.tail_part[((size_t)fc * prm.nsl + slice) * kTailMaxRows + r] = spow[r];
    }
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    named_bar_sync(1, 64);
    if (tid == 0) {
        const int old = atomicAdd(prm.ready + fc, 1);
        if (old == prm.nsl - 1 && prm.power && rem > 0) {
            // last slice of the bin: sum the partial power sums of the tail rows in a fixed order
            __threadfence();
            for (int r = 0; r < rem; ++r) {
                float acc = 0.f;
                for (int sl = 0; sl < prm.nsl; ++sl)
                    acc += __ldcg(prm.tail_part + ((size_t)fc * prm.nsl + sl) * kTailMaxRows + r);
                prm.power[fc * prm.Lrows + prm.Lm + r] = acc / (float)T;
            }
        }
    }
}

template <bool kHybrid>
__device__ void convert_ahead(const Params& prm, Barriers* bars, unsigned char* smem, int cw, int lane,
                              long long cluster_id, long long n_clusters, int crank) {
    (void)bars;
    // prologue: the first conv_D bins, spread over the whole grid
    const long long units = (long long)prm.conv_D * prm.nsl;
    for (long long u = blockIdx.x; u < units; u += gridDim.x)
        convert_slice<kHybrid>(prm, smem, cw, lane, u / prm.nsl, (int)(u % prm.nsl));
    // main loop: not paced by the CTA's own tasks -- the converters run ahead as far as the plane ring allows
    // (convert_slice waits for the previous tenant of the slot), which absorbs the DRAM latency of the X reads
    // under the kernel's store traffic.  (An L2 prefetch of the next slice's rows was measured: +0.4 ms per step.)
    for (long long grp = cluster_id; grp < prm.groups; grp += n_clusters) {
        const long long f = grp / prm.lt_groups;
        const long long fc = f + prm.conv_D;
        if (fc >= prm.F) break;                                    // bins only grow with grp
        const int slice = (int)(grp - f * prm.lt_groups) * prm.csize + crank;
        convert_slice<kHybrid>(prm, smem, cw, lane, fc, slice);
    }
}

// Pre-pass over the spectra, one CTA per bin:
//  (1) X[f][q][t] complex64 -> Xs[f][4 planes][t][Kp] fp32, K-major (mic index contiguous), split for the B side
//      of the MMA: planes (re_hi, im_hi, re_x, im_x), x = lo (3xTF32) or packed bf16 (lo, hi) (hybrid); a
//      [64 q x 32 t] tile is transposed through shared memory (reads coalesced along t, writes along q);
//  (2) the last few look directions when L is not a multiple of 128 (icosahedral / Fibonacci grids: 42, 162,
//      642, 2562 = 20 x 128 + 2).  A whole 128-row tensor-core tile -- and a ghost tile for its cluster partner --
//      for two rows costs 9 % of the fused kernel's time; on the CUDA cores, from the tile that is in shared
//      memory anyway, they are almost free:
//        out[f, Lm + r, t] = sum_q A_f[r, q] X[f, q, t],   A_f[r, q] = sum_g d_f[g] G_g[Lm + r, q]   (FP32 FMA)
//      (+ their power-map entries: the CTA sees every frame of its bin, so no atomics).
constexpr int kSplitT = 32;                          // frames per tile of the pre-pass
__global__ void __launch_bounds__(256)
x_split_tail_kernel(const float2* __restrict__ X, int Q, long long T, int Kp, int hybrid, float* __restrict__ Xs,
                    const float4* __restrict__ G2_last, int ngroups, int Kh, const double2* __restrict__ d,
                    float2* __restrict__ out, long long L, long long Lm, int rem, long long ldo,
                    float* __restrict__ power) {
    __shared__ float2 tile[kMaxK][kSplitT + 1];
    __shared__ float2 sA[kTailMaxRows][kMaxK];                      // steering rows of the tail look directions
    __shared__ float spow[kTailMaxRows];
    const long long f = blockIdx.x;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    if (rem > 0) {
        for (int i = threadIdx.x; i < rem * Kh; i += blockDim.x) {
            const int r = i / Kh, j = i - r * Kh;
            float ar0 = 0.f, ai0 = 0.f, ar1 = 0.f, ai1 = 0.f;
            for (int g = 0; g < ngroups; ++g) {
                const float2 dg = ld_filter(d + f * ngroups + g);
                const float4 v = __ldg(G2_last + ((size_t)g * Kh + j) * kTileL + r);
                ar0 = fmaf(dg.x, v.x, ar0);
                ar0 = fmaf(-dg.y, v.y, ar0);
                ai0 = fmaf(dg.x, v.y, ai0);
                ai0 = fmaf(dg.y, v.x, ai0);
                ar1 = fmaf(dg.x, v.z, ar1);
                ar1 = fmaf(-dg.y, v.w, ar1);
                ai1 = fmaf(dg.x, v.w, ai1);
                ai1 = fmaf(dg.y, v.z, ai1);
            }
            sA[r][2 * j] = make_float2(ar0, ai0);
            sA[r][2 * j + 1] = make_float2(ar1, ai1);
        }
        if (threadIdx.x < kTailMaxRows) spow[threadIdx.x] = 0.f;
    }
    const long long plane = T * Kp;
    float* const base = Xs + f * 4 * plane;
    const float2* const xf = X + f * Q * T;
    float psum = 0.f;                                                // thread (r = tid / 32, t = tid % 32) of the tail
    for (long long t0 = 0; t0 < T; t0 += kSplitT) {
        __syncthreads();                                             // previous tile consumed (and sA complete)
#pragma unroll
        for (int i = 0; i < kMaxK / 8; ++i) {
            const int q = ty + 8 * i;
            const long long t = t0 + tx;
            tile[q][tx] = (q < Q && t < T) ? __ldg(xf + (long long)q * T + t) : make_float2(0.f, 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kSplitT / 8; ++i) {
            const long long t = t0 + ty + 8 * i;
            if (t < T) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int q = tx + 32 * h;
                    if (q < Kp) {
                        const float2 v = tile[q][ty + 8 * i];
                        const float hr = __uint_as_float(to_tf32(v.x)), hi = __uint_as_float(to_tf32(v.y));
                        float xr = v.x - hr, xi = v.y - hi;
                        if (hybrid) {                                // B side: (lo, hi) = (even, odd) bf16 k
                            xr = __uint_as_float(pack_bf16(xr, hr));
                            xi = __uint_as_float(pack_bf16(xi, hi));
                        }
                        float* o = base + t * Kp + q;
                        o[0] = hr;
                        o[plane] = hi;
                        o[2 * plane] = xr;
                        o[3 * plane] = xi;
                    }
                }
            }
        }
        // tail rows: thread (r, t) = (tid / 32 + 8 k, tid % 32)
        for (int r = ty; r < rem; r += 8) {
            const long long t = t0 + tx;
            float ar = 0.f, ai = 0.f;
            for (int q = 0; q < Q; ++q) {
                const float2 a = sA[r][q];
                const float2 x = tile[q][tx];
                ar = fmaf(a.x, x.x, ar);
                ar = fmaf(-a.y, x.y, ar);
                ai = fmaf(a.x, x.y, ai);
                ai = fmaf(a.y, x.x, ai);
            }
            if (t < T) {
                if (out) out[(f * L + Lm + r) * ldo + t] = make_float2(ar, ai);
                if (power) {
                    float p = ar * ar + ai * ai;
                    for (int off = 16; off; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
                    if (tx == 0) spow[r] += p;                       // row r is always handled by the same warp
                }
            } else if (power) {
                float p = 0.f;
                for (int off = 16; off; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
            }
        }
    }
    (void)psum;
    if (power && rem > 0) {
        __syncthreads();
        if (threadIdx.x < rem) power[f * L + Lm + threadIdx.x] = spow[threadIdx.x] / (float)T;
    }
}

__device__ __forceinline__ int order_of_col(int i) {     // radial.repeat_n_m: column i -> order n
    int n = (int)sqrtf((float)i);
    while ((n + 1) * (n + 1) <= i) ++n;
    while (n * n > i) --n;
    return n;
}

// Group products in the builders' layout (FP64 accumulate, stored as float):
//   G2[lt][g][j][lane] = (G_g[l][2j], G_g[l][2j+1]),  l = lt*128 + lane,  G_g[l][q] = sum_{m in g} Y_L[l,m] conj(Y_Q[q,m])
// Rows l >= L and columns q >= Q are zero.  One block per (l-tile, group), one thread per row: the thread's
// Y_L[l, m0..m1) stays in L1 over the 32 column pairs, conj(Y_Q[q, m]) is the same address for the whole warp,
// and every store is 16 contiguous bytes per lane (2 KB per block and column pair); blockIdx.y splits the column
// pairs so that the grid fills the machine.  Reads Y_Q as it is (no conjugate-transposed copy).
__global__ void __launch_bounds__(kTileL)
group_products_g2_kernel(const double2* __restrict__ YL, const double2* __restrict__ YQ, long long L, int Q, int M,
                         int G, int per_order, int Kh, float4* __restrict__ out, int model,
                         unsigned* __restrict__ devmag, float2* __restrict__ sq_out, float* __restrict__ xt) {
    extern __shared__ double2 sq[];                              // conj-free copy of Y_Q[q0 .. q1)[m0 .. m1): [q][nm]
    const int lane = threadIdx.x;
    const long long lt = blockIdx.x / G;
    const int g = (int)(blockIdx.x - lt * G);
    const long long l = lt * kTileL + lane;
    const int m0 = per_order ? g * g : g;
    int m1 = per_order ? (g + 1) * (g + 1) : g + 1;
    if (m1 > M) m1 = M;
    const int nm = m1 - m0;
    float4* o = out + ((size_t)(lt * G + g) * Kh) * kTileL + lane;
    const int jper = (Kh + gridDim.y - 1) / gridDim.y;       // column pairs per block (blockIdx.y splits them)
    const int j0 = blockIdx.y * jper, j1 = (j0 + jper < Kh) ? j0 + jper : Kh;
    const int nq = 2 * (j1 - j0);
    for (int i = lane; i < nq * nm; i += kTileL) {
        const int qq = i / nm, mm = i - qq * nm;
        const int q = 2 * j0 + qq;
        sq[i] = (q < Q) ? YQ[(long long)q * M + m0 + mm] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    // Structure models, everything in FP64, s_q = Y_L[0, 0] conj(Y_Q[q, 0]):
    //   Legendre (per-order groups, M >= 4): x_lq = Re(G_1[l, q] / (3 s_q)), model G_g[l, q] = s_q (2g + 1) P_g(x_lq)
    //   rotation (one group per column, M = 2N + 1 >= 3, column c of order m = c or c - M):
    //            z_lq = G_1[l, q] / s_q, model G_c[l, q] = s_q z^m (m >= 0), s_q conj(z)^|m| (m < 0)
    const bool leg = model != kModelNone;
    double2 yl1 = make_double2(0.0, 0.0), yl2 = yl1, yl3 = yl1, y00 = yl1;
    if (leg) {
        y00 = YL[0];
        if (l < L) {
            yl1 = YL[l * M + 1];
            if (model == kModelLegendre) {
                yl2 = YL[l * M + 2];
                yl3 = YL[l * M + 3];
            }
        }
        if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) devmag[128] = (unsigned)model;
    }
    double ldev = 0.0, lmag = 0.0;
    for (int j = j0; j < j1; ++j) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (l < L) {
            const double2* arow = YL + l * M + m0;
            const double2* b0 = sq + (size_t)(2 * (j - j0)) * nm;
            const double2* b1 = b0 + nm;
            for (int m = 0; m < nm; ++m) {
                const double2 a = arow[m];
                const double2 x = b0[m], y = b1[m];              // conj(x) = (x.x, -x.y)
                s0 += a.x * x.x + a.y * x.y;
                s1 += a.y * x.x - a.x * x.y;
                s2 += a.x * y.x + a.y * y.y;
                s3 += a.y * y.x - a.x * y.y;
            }
        }
        o[(size_t)j * kTileL] = make_float4((float)s0, (float)s1, (float)s2, (float)s3);
        if (leg && l < L) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int q = 2 * j + c;
                if (q >= Q) continue;
                const double gre = c ? s2 : s0, gim = c ? s3 : s1;
                const double2* yq = YQ + (long long)q * M;
                const double2 q0 = yq[0], q1 = yq[1];
                const double sre = y00.x * q0.x + y00.y * q0.y, sim = y00.y * q0.x - y00.x * q0.y;       // y00 conj(q0)
                if (model == kModelRotation) {
                    const double h1re = yl1.x * q1.x + yl1.y * q1.y, h1im = yl1.y * q1.x - yl1.x * q1.y;  // G_1[l, q]
                    const double den = sre * sre + sim * sim;
                    const double zr = den > 0.0 ? (h1re * sre + h1im * sim) / den : 0.0;                 // G_1 conj(s) / |s|^2
                    const double zi = den > 0.0 ? (h1im * sre - h1re * sim) / den : 0.0;
                    const int mo = (g <= (M - 1) / 2) ? g : g - M;
                    double wr = 1.0, wi = 0.0;
                    for (int kk = 0; kk < (mo < 0 ? -mo : mo); ++kk) {
                        const double t = wr * zr - wi * zi;
                        wi = wr * zi + wi * zr;
                        wr = t;
                    }
                    if (mo < 0) wi = -wi;
                    const double dre = fabs(gre - (sre * wr - sim * wi)), dim = fabs(gim - (sre * wi + sim * wr));
                    if (!(dre <= ldev)) ldev = dre;
                    if (!(dim <= ldev)) ldev = dim;
                    lmag = fmax(lmag, fmax(fabs(gre), fabs(gim)));
                    if (g == 1) {
                        float* tq = xt + (((size_t)lt * kLegTableRow + (q >> 1)) * kTileL + lane) * 4 + (q & 1) * 2;
                        tq[0] = (float)zr;
                        tq[1] = (float)zi;
                    }
                    if (g == 0 && l == 0) sq_out[q] = make_float2((float)sre, (float)sim);
                    continue;
                }
                const double2 q2 = yq[2], q3 = yq[3];
                const double g1re = yl1.x * q1.x + yl1.y * q1.y + yl2.x * q2.x + yl2.y * q2.y + yl3.x * q3.x + yl3.y * q3.y;
                const double g1im = yl1.y * q1.x - yl1.x * q1.y + yl2.y * q2.x - yl2.x * q2.y + yl3.y * q3.x - yl3.x * q3.y;
                const double den = 3.0 * (sre * sre + sim * sim);
                double x = den > 0.0 ? (g1re * sre + g1im * sim) / den : 0.0;                            // Re(g1 / (3 s))
                x = x > 1.0 ? 1.0 : (x < -1.0 ? -1.0 : x);
                double p0 = 1.0, p1 = x;
                for (int n = 1; n < g; ++n) {
                    const double pn = ((2 * n + 1) * x * p1 - n * p0) / (n + 1);
                    p0 = p1;
                    p1 = pn;
                }
                const double pg = (g == 0 ? 1.0 : p1) * (2 * g + 1);
                const double dre = fabs(gre - sre * pg), dim = fabs(gim - sim * pg);
                // (a NaN compares false everywhere: fmax would drop it, so it is carried by hand)
                if (!(dre <= ldev)) ldev = dre;
                if (!(dim <= ldev)) ldev = dim;
                lmag = fmax(lmag, fmax(fabs(gre), fabs(gim)));
                if (g == 1) xt[(((size_t)lt * kLegTableRow + (q >> 2)) * kTileL + lane) * 4 + (q & 3)] = (float)x;
                if (g == 0 && l == 0) sq_out[q] = make_float2((float)sre, (float)sim);
            }
        }
    }
    if (leg) {
        // non-negative floats (and NaN, whose bits lie above +inf) order like their bit patterns
        unsigned ud = __float_as_uint(fabsf((float)ldev)), um = __float_as_uint(fabsf((float)lmag));
        if (ldev > 0.0 && ud == 0u) ud = 1u;                       // deviations below the float range still count
        for (int off = 16; off; off >>= 1) {
            ud = max(ud, __shfl_xor_sync(0xffffffffu, ud, off));
            um = max(um, __shfl_xor_sync(0xffffffffu, um, off));
        }
        if ((lane & 31) == 0) {
            atomicMax(devmag + g, ud);
            atomicMax(devmag + 64 + g, um);
        }
    }
}

}  // namespace fused

// ------------------------------------------------------------------ host side
bool pwd_fused_supported(long long L, long long Q, long long T, int Cd, long long out_ld, const void* out) {
    (void)L;
    return !options().no_fused && Q >= 1 && Q <= fused::kMaxK && Cd >= 1 && Cd <= 64 && T >= 1 &&
           (out_ld & 1) == 0 && (((uintptr_t)out) & 15) == 0;
}

// bytes of the builders' group-product table and of the pre-split spectra of `nf` bins
static size_t fused_g2_table_bytes(long long L, long long Q, int Cd) {
    const long long ltiles = ceil_div(L, fused::kTileL);
    const long long Kh = ceil_div(Q, 8) * 4;
    return ((size_t)ltiles * Cd * Kh * fused::kTileL * sizeof(float4) + 255) / 256 * 256;
}
static size_t fused_leg_bytes(long long L) {
    return fused::kLegHdrBytes + (size_t)ceil_div(L, fused::kTileL) * fused::kLegTableRow * fused::kTileL * sizeof(float4);
}
// group-product table + the addition-theorem block (deviation record, s_q, x table)
size_t pwd_fused_g2_bytes(long long L, long long Q, int Cd) {
    return fused_g2_table_bytes(L, Q, Cd) + fused_leg_bytes(L);
}
static size_t fused_planes_bytes(long long nf, long long Q, long long T) {
    const long long Kp = ceil_div(Q, 4) * 4;
    return ((size_t)nf * 4 * T * Kp * sizeof(float) + 255) / 256 * 256;
}
static size_t fused_ready_bytes(long long nf) { return ((size_t)2 * nf * sizeof(int) + 255) / 256 * 256; }   // ready + done
// split planes + hand-over counters + partial power sums of the tail rows (stream-ahead conversion)
size_t pwd_fused_xs_bytes(long long nf, long long L, long long Q, long long T) {
    const long long max_slices = ceil_div(L, fused::kTileL) + 4;     // slices per bin <= l-tiles rounded up to the cluster size
    return fused_planes_bytes(nf, Q, T) + fused_ready_bytes(nf) +
           (size_t)nf * max_slices * fused::kTailMaxRows * sizeof(float);
}

int launch_group_products_g2(const double2* YL, const double2* YQ, long long L, long long Q, int M, int Cd,
                             int per_order, float* G2, cudaStream_t st) {
    const long long ltiles = ceil_div(L, fused::kTileL);
    const int Kh = (int)ceil_div(Q, 8) * 4;
    // about eight blocks per SM
    long long ysplit = ceil_div((long long)sm_count() * 8, ltiles * Cd);
    if (ysplit > Kh) ysplit = Kh;
    if (ysplit < 1) ysplit = 1;
    const long long jper = ceil_div(Kh, ysplit);
    const int nm_max = per_order ? 2 * Cd - 1 : 1;           // the largest group: order Cd - 1 has 2 Cd - 1 columns
    const size_t smem = (size_t)2 * jper * nm_max * sizeof(double2);
    if (smem > 48 * 1024) return SFA_ERR_UNSUPPORTED;
    // addition-theorem block: zeroed (x = 0, s = 0 for padding rows / columns; deviations start at 0), or -- when the
    // model does not apply -- a deviation record that can never pass (0x7f7f7f7f = 3.4e38 against 1e-8 * 3.4e38)
    unsigned char* legb = reinterpret_cast<unsigned char*>(G2) + fused_g2_table_bytes(L, Q, Cd);
    int model = fused::kModelNone;
    if (!options().no_legendre) {
        if (per_order && Cd >= 2 && Cd <= fused::kLegMaxGroups && M >= 4) model = fused::kModelLegendre;
        else if (!per_order && Cd == M && (M & 1) && M >= 3 && M <= 63) model = fused::kModelRotation;
    }
    SFA_CUDA_CHECK(cudaMemsetAsync(legb, 0, fused_leg_bytes(L), st));
    if (model == fused::kModelNone) SFA_CUDA_CHECK(cudaMemsetAsync(legb, 0x7f, 128 * sizeof(unsigned), st));
    fused::group_products_g2_kernel<<<dim3((unsigned)(ltiles * Cd), (unsigned)ysplit), fused::kTileL, smem, st>>>(
        YL, YQ, L, (int)Q, M, Cd, per_order, Kh, reinterpret_cast<float4*>(G2), model,
        reinterpret_cast<unsigned*>(legb), reinterpret_cast<float2*>(legb + fused::kLegHdrWords * sizeof(unsigned)),
        reinterpret_cast<float*>(legb + fused::kLegHdrBytes));
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

// q[f] = (sum_g d[f][g] G_g) X[f] for nf bins.  d: complex128 [nf][Cd]; X: [nf][Q][T]; out: [nf][L][ldo].
int launch_pwd_fused(long long nf, long long L, long long Q, long long T, int Cd, const float* G2, const double2* d,
                     const float2* X, float* Xs, float2* out, long long ldo, float* power, int hybrid,
                     cudaStream_t st, int pad_writable) {
    using namespace fused;
    if (nf <= 0) return SFA_OK;
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return SFA_ERR_UNSUPPORTED;
    const int Kp = (int)ceil_div(Q, 4) * 4;
    // look directions beyond the last full 128-row tile: few of them -> CUDA cores (no ragged tile, no ghost CTA)
    const long long rem = L % kTileL;
    const bool tail = rem > 0 && rem <= kTailMaxRows && L > kTileL;
    const long long Lmain = tail ? L - rem : L;
    const int Kh = (int)ceil_div(Q, 8) * 4;
    const float4* g2_last = reinterpret_cast<const float4*>(G2) + (size_t)(Lmain / kTileL) * Cd * Kh * kTileL;
    Params prm;
    prm.F = nf;
    prm.L = Lmain;
    prm.Lrows = L;
    prm.T = T;
    prm.K = (int)Q;
    prm.ksteps = (int)ceil_div(Q, 8);
    prm.kboxes = (int)ceil_div(Q, kBoxK);
    prm.ngroups = Cd;
    prm.G2 = reinterpret_cast<const float4*>(G2);
    prm.Kh = prm.ksteps * 4;
    prm.d = d;
    prm.ltiles = (int)ceil_div(Lmain, kTileL);
    prm.csize = (prm.ltiles >= 2 && !options().no_cluster) ? 2 : 1;
    // CTA pairs on one tcgen05.mma (cta_group::2) instead of two MMA streams sharing a multicast plane stream
    const bool pair = prm.csize == 2 && options().fused_pair != 0;
    prm.lt_groups = (prm.ltiles + prm.csize - 1) / prm.csize;
    prm.t_tiles = (int)ceil_div(T, kTileT);
    prm.groups = nf * prm.lt_groups;
    prm.power = power;
    prm.store_evict_first = ((double)nf * (double)L * (double)ldo * 8.0 > 256e6) ? 1 : 0;     // >> L2 (126 MB)
#ifdef SFA_DEBUG_HOOKS
    if (options().fused_debug & 512) prm.store_evict_first = 0;
#endif
    prm.no_store = out ? 0 : 1;
    if (!out && !power) return SFA_ERR_INVALID;
    const long long max_clusters = persistent_sms() / prm.csize;
    const long long n_clusters = prm.groups < max_clusters ? prm.groups : max_clusters;
    // Stream-ahead conversion: the CTAs working on bin f split the spectra of bin f + D, D = the number of bins
    // the grid works on at once, plus two of slack; the first D bins come from the pre-pass.  Only worth it when
    // the launch covers several rounds of bins.
    const long long in_flight = ceil_div(n_clusters, prm.lt_groups);
    // (look-ahead of ONE round of bins and a ring of three rounds measured best on the headline shape: 5.0 ms per step
    // against 5.6 ms with two rounds ahead and five rounds of planes -- the smaller ring leaves more of the L2 to the
    // store stream; tools/fused_knobs.py)
    long long D = options().fused_D > 0 ? options().fused_D : in_flight;
    if (options().fused_prepass || nf < 2 * D) D = 0;
    prm.conv_D = (int)D;
    // plane slots: the bins reuse a ring of D + 2 rounds of bins, so the planes are overwritten while still dirty in
    // L2 and never written back: DRAM sees X once and the beams once
    long long ring = D + 2 * in_flight;
    if (options().fused_ring > 0) ring = options().fused_ring > D ? options().fused_ring : D + 1;
    if (D <= 0 || ring >= nf) ring = nf;
    prm.ring = (int)ring;
    prm.nsl = prm.lt_groups * prm.csize;
    prm.Q = (int)Q;
    prm.Kp = Kp;
    prm.rem = tail ? (int)rem : 0;
    prm.X = X;
    prm.Xs = Xs;
    prm.G2_last = g2_last;
    {
        const unsigned char* legb = reinterpret_cast<const unsigned char*>(G2) + fused_g2_table_bytes(L, Q, Cd);
        prm.leg_devmag = reinterpret_cast<const unsigned*>(legb);
        prm.leg_sq = reinterpret_cast<const float2*>(legb + kLegHdrWords * sizeof(unsigned));
        prm.leg_xt = reinterpret_cast<const float4*>(legb + kLegHdrBytes);
    }
    prm.out = out;
    prm.ldo = ldo;
    prm.Lm = Lmain;
    unsigned char* extra = reinterpret_cast<unsigned char*>(Xs) + fused_planes_bytes(nf, Q, T);
    prm.ready = reinterpret_cast<int*>(extra);
    prm.done = prm.ready + nf;
    prm.tail_part = reinterpret_cast<float*>(extra + fused_ready_bytes(nf));
    if (D > 0) {
        SFA_CUDA_CHECK(cudaMemsetAsync(prm.ready, 0, (size_t)2 * nf * sizeof(int), st));
    } else {
        prof_aux_begin(st);
        x_split_tail_kernel<<<(unsigned)nf, 256, 0, st>>>(X, (int)Q, T, Kp, hybrid, Xs, g2_last, Cd, Kh, d, out, L, Lmain,
                                                          tail ? (int)rem : 0, ldo, power);
        prof_aux_end(st);
        SFA_LAUNCH_CHECK();
    }
    CUtensorMap xmap, omap;
    {
        cuuint64_t gdim[3] = {(cuuint64_t)Q, (cuuint64_t)T, (cuuint64_t)(4 * prm.ring)};
        cuuint64_t gstr[2] = {(cuuint64_t)Kp * 4, (cuuint64_t)T * Kp * 4};
        // (pair mode: a CTA loads its own half of the frame tile)
        cuuint32_t box[3] = {(cuuint32_t)kBoxK, (cuuint32_t)(pair ? kTileT / 2 : kTileT), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)Xs, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_last_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(Xs)");
            return SFA_ERR_CUDA;
        }
        // beams as fp32 [nf][L][2*ldo]; store box = 32 rows x 32 fp32 (16 frames), SWIZZLE_128B staging
        // (power map only: the map is encoded over the plane buffer and never used).
        // Row ends: with T = 938 a row is 58.6 lines of 128 bytes, and a tensor map of extent T clips the last box
        // of every row to a partial line -- which costs the WHOLE store stream 15-20 % of its bandwidth (pure-store
        // probe, tools/store_probe5.cu: 4.76 TB/s with the clipped row ends, 5.74 TB/s with whole lines; the L2
        // has to fetch the rest of a partially written line from DRAM before it can write it back).  When the
        // caller owns the row padding (sfa_pwd_shape.out_pad_writable) the map extends to the next multiple of 16
        // frames: the frames [T, T16) are written as zeros (the TMA load of the spectra zero-fills frames >= T).
        const long long T16 = ceil_div(T, 16) * 16;
        long long Tmap = (pad_writable && ldo >= T16) ? T16 : T;
        // a TMA store clips in 16-byte units: an odd T (8 T bytes per row) would take frame T with it
        prm.odd_last = (Tmap == T && (T & 1) && ldo > T) ? 1 : 0;
        if (prm.odd_last) Tmap = T - 1;
        prm.Tst = Tmap < T ? Tmap : T;
        if (Tmap < 1) Tmap = 1;                                      // T = 1: every store is the epilogue's own
        void* const obase = out ? (void*)out : (void*)Xs;
        cuuint64_t odim[3] = {(cuuint64_t)(2 * Tmap), (cuuint64_t)Lmain, (cuuint64_t)nf};
        cuuint64_t ostr[2] = {(cuuint64_t)ldo * 8, (cuuint64_t)L * ldo * 8};
        cuuint32_t obox[3] = {32, 32, 1};
        r = encode(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, obase, odim, ostr, obox, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_last_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(beams)");
            return SFA_ERR_CUDA;
        }
    }
#ifdef SFA_DEBUG_HOOKS
    prm.debug = options().fused_debug;
#else
    prm.debug = 0;
#endif
    const size_t smem = (pair ? (size_t)kStages2 * (kStageBytes / 2) + kEpiWarps * (kStoreRing2 * 2 * kStoreBoxBytes)
                              : (size_t)kStages * kStageBytes + kEpiWarps * kStoreWarpBytes) +
                        256 + kConvSmemBytes + 1024;
    static_assert(sizeof(Barriers) <= 256, "barrier block");
    const long long grid = n_clusters * prm.csize;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = prm.csize;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaError_t le;
#define SFA_FUSED_LAUNCH(HYB, NGV)                                                                  \
    do {                                                                                            \
        if (pair) {                                                                                 \
            SFA_ENSURE_DYN_SMEM((pwd_fused_kernel<HYB, NGV, true>), smem);                          \
            prof_begin(st);                                                                         \
            le = cudaLaunchKernelEx(&cfg, pwd_fused_kernel<HYB, NGV, true>, xmap, omap, prm);       \
            prof_end(st);                                                                           \
        } else {                                                                                    \
            SFA_ENSURE_DYN_SMEM((pwd_fused_kernel<HYB, NGV, false>), smem);                         \
            prof_begin(st);                                                                         \
            le = cudaLaunchKernelEx(&cfg, pwd_fused_kernel<HYB, NGV, false>, xmap, omap, prm);      \
            prof_end(st);                                                                           \
        }                                                                                           \
    } while (0)
    if (Cd <= 9) {
        if (hybrid) SFA_FUSED_LAUNCH(true, 9);
        else SFA_FUSED_LAUNCH(false, 9);
    } else if (Cd <= 16) {
        if (hybrid) SFA_FUSED_LAUNCH(true, 16);
        else SFA_FUSED_LAUNCH(false, 16);
    } else {
        if (hybrid) SFA_FUSED_LAUNCH(true, 0);
        else SFA_FUSED_LAUNCH(false, 0);
    }
#undef SFA_FUSED_LAUNCH
    if (le != cudaSuccess) {
        set_last_cuda_error(le, "cudaLaunchKernelEx(pwd_fused_kernel)");
        return SFA_ERR_CUDA;
    }
    SFA_LAUNCH_CHECK();
    return SFA_OK;
}

}  // namespace sfa
