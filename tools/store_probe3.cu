// HBM write-pattern probe 3: the store stream of the fused steering kernel (pwd_fused_sm100.cu) in isolation.
// 148 persistent CTAs; a task = (bin f, block of 128 look directions); per frame tile (64 frames) the four
// "epilogue" warps each write their 32 rows x 64 frames as TMA store boxes.  Variants: box shape, L2 cache hint,
// row pitch, frames per step, plain st.global instead of TMA.  Pure stores, no compute.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/store_probe3 tools/store_probe3.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

constexpr long long F = 1024, L = 2562, T = 938;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// hint: 0 none, 1 evict_last, 2 evict_first, 3 evict_normal (explicit)
__device__ __forceinline__ void tma_store(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int hint,
                                          uint64_t pol) {
    if (hint == 0)
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                     "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                     : "memory");
    else
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(map),
                     "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
                     : "memory");
}

// box_cols: fp32 per box row (32 = 128 B, 64 = 256 B, 128 = 512 B); W: frames per step (64 or 128)
__global__ void __launch_bounds__(128) probe_tma(const __grid_constant__ CUtensorMap map, int box_cols, int W, int hint,
                                                 int pair_order) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t pol = 0;
    if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (hint == 3) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    const long long ltiles = (L + 127) / 128;
    const long long tasks = F * ltiles;
    unsigned char* my = smem + warp * 16384;
    const int boxes_per_step = (2 * W) / box_cols;                // per warp: 32 rows x 2W fp32
    const long long lt_groups = (ltiles + 1) / 2;
    const long long iters = pair_order ? (F * lt_groups + gridDim.x / 2 - 1) / (gridDim.x / 2) : (tasks + gridDim.x - 1) / gridDim.x;
    for (long long it = 0; it < iters; ++it) {
        long long f, lt;
        if (pair_order) {                                          // CTAs 2j, 2j+1 take adjacent l-tiles of one bin
            const long long grp = blockIdx.x / 2 + it * (gridDim.x / 2);
            f = grp / lt_groups;
            lt = (grp % lt_groups) * 2 + (blockIdx.x & 1);
        } else {
            const long long task = blockIdx.x + it * gridDim.x;
            f = task / ltiles;
            lt = task % ltiles;
        }
        if (f >= F || lt >= ltiles) continue;
        const int l0 = (int)lt * 128 + warp * 32;
        if (l0 >= L) continue;
        for (long long t0 = 0; t0 < T; t0 += W) {
            if (lane == 0) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                for (int b = 0; b < boxes_per_step; ++b) {
                    const long long c0 = 2 * t0 + (long long)b * box_cols;
                    if (c0 < 2 * T) tma_store(&map, my, (int)c0, l0, (int)f, hint, pol);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            __syncwarp();
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// same pattern with plain 16-byte stores: lane = row, each lane writes its row piece (the register layout of the
// epilogue) -- mode 0; or coalesced (a warp writes 512 contiguous bytes of one row) -- mode 1
__global__ void __launch_bounds__(128) probe_stg(float4* out, long long pitch, int W, int coalesced) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ltiles = (L + 127) / 128;
    const long long tasks = F * ltiles;
    for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
        const long long f = task / ltiles, lt = task % ltiles;
        const long long l0 = lt * 128 + warp * 32;
        for (long long t0 = 0; t0 < T; t0 += W) {
            if (coalesced) {
                for (int r = 0; r < 32; ++r) {
                    const long long row = l0 + r;
                    for (int c = lane; c < W / 2; c += 32) {
                        const long long t = t0 + 2 * c;
                        if (row < L && t + 1 < T) out[((f * L + row) * pitch + t) / 2] = make_float4(1.f, 2.f, 3.f, 4.f);
                    }
                }
            } else {
                const long long row = l0 + lane;
                for (int c = 0; c < W / 2; ++c) {
                    const long long t = t0 + 2 * c;
                    if (row < L && t + 1 < T) out[((f * L + row) * pitch + t) / 2] = make_float4(1.f, 2.f, 3.f, 4.f);
                }
            }
        }
    }
}

int main() {
    float* out;
    const size_t bytes_alloc = (size_t)F * L * 1024 * 8;
    const size_t bytes = (size_t)F * L * T * 8;
    cudaMalloc(&out, bytes_alloc);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(p);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaFuncSetAttribute(probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
    struct Cfg { long long pitch; int box_cols, W, hint, pair, swz; };
    std::vector<Cfg> cfgs = {
        {960, 32, 64, 0, 0, 1},  {960, 32, 64, 1, 0, 1},  {960, 32, 64, 2, 0, 1},  {960, 32, 64, 3, 0, 1},
        {960, 64, 64, 0, 0, 0},  {960, 128, 64, 0, 0, 0}, {960, 128, 64, 1, 0, 0},
        {960, 32, 128, 0, 0, 1}, {960, 128, 128, 0, 0, 0},
        {960, 32, 64, 0, 1, 1},  {960, 128, 64, 0, 1, 0},
        {1024, 32, 64, 0, 0, 1}, {1024, 128, 64, 0, 0, 0}, {944, 32, 64, 0, 0, 1}, {992, 32, 64, 0, 0, 1},
        {938, 32, 64, 0, 0, 1},
    };
    for (auto& c : cfgs) {
        CUtensorMap map;
        cuuint64_t odim[3] = {(cuuint64_t)(2 * T), (cuuint64_t)L, (cuuint64_t)F};
        cuuint64_t ostr[2] = {(cuuint64_t)c.pitch * 8, (cuuint64_t)L * c.pitch * 8};
        cuuint32_t obox[3] = {(cuuint32_t)c.box_cols, 32, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)out, odim, ostr, obox, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, c.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            printf("encode failed %d\n", (int)r);
            continue;
        }
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            probe_tma<<<148, 128, 65536 + 1024>>>(map, c.box_cols, c.W, c.hint, c.pair);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
        }
        printf("TMA pitch %4lld box %3d fp32 x 32 rows, %3d frames/step, hint %d, pair %d: %6.3f ms %5.2f TB/s  (%s)\n", c.pitch,
               c.box_cols, c.W, c.hint, c.pair, ms, bytes / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    for (long long pitch : {960LL, 1024LL}) {
        for (int W : {64, 128}) {
            for (int co = 0; co < 2; ++co) {
                float ms = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(a);
                    probe_stg<<<148, 128>>>(reinterpret_cast<float4*>(out), pitch, W, co);
                    cudaEventRecord(b);
                    cudaEventSynchronize(b);
                    cudaEventElapsedTime(&ms, a, b);
                }
                printf("STG pitch %4lld %3d frames/step %s: %6.3f ms %5.2f TB/s\n", pitch, W,
                       co ? "coalesced (warp = 512 B of one row)" : "lane = row (16 B per lane)", ms, bytes / ms / 1e9);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
