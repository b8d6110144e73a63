// HBM write-pattern probe 5: tensor-map stores against 1-D bulk copies for the fused kernel's row-tile pattern, in ONE
// binary (probes 3 and 4 ran on different boxes of a pool whose boxes differ by up to 20 %), and the issue cost of a
// bulk copy when every lane of a warp issues one (lane = row) instead of one lane issuing 32.
//   task = (bin, 128 rows); per 64-frame step each of 4 warps writes 32 row pieces of 512 B
//   mode 0: tensor-map box 128 fp32 x 32 rows (no swizzle), lane 0 issues
//   mode 1: tensor-map box 32 fp32 x 32 rows, SWIZZLE_128B, 4 per step, lane 0 issues (the kernel's store today)
//   mode 2: 1-D bulk copies, lane 0 issues 32 per step
//   mode 3: 1-D bulk copies, every lane issues the copy of its own row (one warp instruction per step)
//   mode 4: as 3 with 256-byte pieces (two instructions per step: the epilogue's two passes of 32 frames)
//   mode 5: 4-D tensor map {32 fp32, 4 chunks, 32 rows, bin}: one instruction per step, 512 contiguous bytes per row
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/store_probe5 tools/store_probe5.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr long long F = 1024, L = 2562, PITCH = 960;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes, int hint, uint64_t pol) {
    if (hint)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
                     "r"(smem_u32(src)), "r"(bytes), "l"(pol)
                     : "memory");
    else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
                     "r"(bytes)
                     : "memory");
}
__device__ __forceinline__ void tma_store3(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store4(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, unsigned char* out, long long T,
                                             int mode, int hint, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* my = smem + warp * 17408;                       // 32 rows x (512 + 16) B fits, 1024-aligned
    uint64_t pol = 0;
    if (hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const long long ltiles = L / 128;
    const long long tasks = F * ltiles;
    const long long rowb = PITCH * 8;
    long long issue = 0;
    for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
        const long long f = task / ltiles, lt = task % ltiles;
        const int l0 = (int)lt * 128 + warp * 32;
        unsigned char* base = out + ((f * L + l0) * rowb);
        for (long long t0 = 0; t0 < T; t0 += 64) {
            const long long nfr = (T - t0) < 64 ? (T - t0) : 64;
            const long long c0 = clock64();
            if (mode <= 2 || mode == 5) {
                if (lane == 0) {
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    if (mode == 0) tma_store3(&map, my, (int)(2 * t0), l0, (int)f);
                    else if (mode == 1) {
                        for (int b = 0; b < 4; ++b)
                            if (2 * t0 + 32 * b < 2 * T) tma_store3(&map, my + b * 4096, (int)(2 * t0 + 32 * b), l0, (int)f);
                    } else if (mode == 5) tma_store4(&map, my, 0, (int)(t0 / 16), l0, (int)f);
                    else
                        for (int r = 0; r < 32; ++r) bulk_store(base + r * rowb + t0 * 8, my + r * 528, (uint32_t)(nfr * 8), hint, pol);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else if (mode == 3) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                bulk_store(base + lane * rowb + t0 * 8, my + lane * 528, (uint32_t)(nfr * 8), hint, pol);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            } else {
                for (int h = 0; h < 2; ++h) {
                    const long long n2 = nfr - 32 * h;
                    if (n2 <= 0) break;
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    bulk_store(base + lane * rowb + (t0 + 32 * h) * 8, my + lane * 528 + h * 256,
                               (uint32_t)((n2 < 32 ? n2 : 32) * 8), hint, pol);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            issue += clock64() - c0;
            __syncwarp();
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (blockIdx.x == 0 && threadIdx.x == 0) *cyc = issue;
}

int main() {
    unsigned char* out;
    long long* cyc;
    const long long total = F * L * PITCH * 8;
    cudaMalloc(&out, total);
    cudaMalloc(&cyc, 8);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(p);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 17408 + 1024);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    struct Cfg { int mode; long long T; int hint; };
    Cfg cfgs[] = {{0, 938, 0}, {0, 960, 0}, {1, 938, 0}, {1, 960, 0}, {2, 938, 0}, {2, 960, 0}, {3, 938, 0}, {3, 960, 0},
                  {3, 938, 1}, {4, 938, 0}, {4, 938, 1}, {5, 938, 0}, {5, 960, 0}, {0, 938, 0}, {3, 938, 0}};
    for (auto& c : cfgs) {
        CUtensorMap map;
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r;
        if (c.mode == 5) {
            cuuint64_t odim[4] = {32, (cuuint64_t)((2 * c.T + 31) / 32), (cuuint64_t)L, (cuuint64_t)F};
            cuuint64_t ostr[3] = {128, (cuuint64_t)PITCH * 8, (cuuint64_t)L * PITCH * 8};
            cuuint32_t obox[4] = {32, 4, 32, 1};
            r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)out, odim, ostr, obox, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            cuuint64_t odim[3] = {(cuuint64_t)(2 * c.T), (cuuint64_t)L, (cuuint64_t)F};
            cuuint64_t ostr[2] = {(cuuint64_t)PITCH * 8, (cuuint64_t)L * PITCH * 8};
            cuuint32_t obox[3] = {(cuuint32_t)(c.mode == 0 ? 128 : 32), 32, 1};
            r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)out, odim, ostr, obox, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, c.mode == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) {
            printf("mode %d: encode failed %d\n", c.mode, (int)r);
            continue;
        }
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a);
            probe<<<148, 128, 4 * 17408 + 1024>>>(map, out, c.T, c.mode, c.hint, cyc);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
        }
        long long h = 0;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)F * 2560 * c.T * 8;
        const double steps = (double)(F * 20 / 148) * ((c.T + 63) / 64);
        printf("mode %d T %3lld hint %d: %6.3f ms  %5.2f TB/s  issue %6.0f cycles/step  (%s)\n", c.mode, c.T, c.hint, ms,
               bytes / ms / 1e9, (double)h / steps, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
