// HBM write-pattern probe 4: is the ~4.95 TB/s of the fused kernel's store stream a property of the TMA store path
// or of the row-tile pattern?  1-D bulk copies (cp.async.bulk.global.shared::cta) of different chunk sizes:
//   mode 0: every CTA streams through its own contiguous 1/grid of the buffer in 16 KB chunks
//   mode 1: 16 KB chunks dealt out round-robin over the CTAs (compact active window, memset-like)
//   mode 2: the kernel's pattern -- task = (bin, 128 rows), per 64-frame step each warp writes 32 row pieces of
//           `piece` bytes (512 = one frame tile; 1024, 2048, 4096 = several tiles at once; 7680 = whole rows)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/store_probe4 tools/store_probe4.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr long long F = 1024, L = 2562, PITCH = 960;       // frames per row (padded): 7680 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}

__global__ void __launch_bounds__(128) probe(unsigned char* out, long long total, int mode, int piece, int depth) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* my = smem + warp * 16384;
    if (lane != 0) return;
    if (mode <= 1) {
        const long long chunks = total / 16384;
        const long long nw = (long long)gridDim.x * 4, me = (long long)blockIdx.x * 4 + warp;
        const long long per = (chunks + nw - 1) / nw;
        for (long long i = 0; i < per; ++i) {
            const long long c = mode == 0 ? me * per + i : i * nw + me;
            if (c >= chunks) break;
            if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            bulk_store(out + c * 16384, my, 16384);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    } else {
        const long long ltiles = L / 128;                          // 20 full tiles
        const long long tasks = F * ltiles;
        const long long rowb = PITCH * 8;
        for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
            const long long f = task / ltiles, lt = task % ltiles;
            unsigned char* base = out + ((f * L + lt * 128 + warp * 32) * rowb);
            for (long long off = 0; off < rowb; off += piece) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                const uint32_t n = (uint32_t)((rowb - off) < piece ? (rowb - off) : piece);
                for (int r = 0; r < 32; ++r) bulk_store(base + r * rowb + off, my, n);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    unsigned char* out;
    const long long total = F * L * PITCH * 8;
    cudaMalloc(&out, total);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    struct Cfg { int mode, piece, depth, grid; };
    Cfg cfgs[] = {{0, 0, 1, 148}, {1, 0, 1, 148}, {1, 0, 3, 148}, {1, 0, 1, 296}, {2, 512, 1, 148}, {2, 1024, 1, 148},
                  {2, 2048, 1, 148}, {2, 3840, 1, 148}, {2, 7680, 1, 148}, {2, 512, 1, 296}, {2, 7680, 1, 296}};
    for (auto& c : cfgs) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            probe<<<c.grid, 128, 65536 + 1024>>>(out, total, c.mode, c.piece, c.depth);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
        }
        const double bytes = c.mode <= 1 ? (double)total : (double)F * 2560 * PITCH * 8;
        printf("mode %d piece %5d depth %d grid %3d: %6.3f ms  %5.2f TB/s  (%s)\n", c.mode, c.piece, c.depth, c.grid, ms,
               bytes / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
