// HBM write-pattern probe: how fast can 148 persistent CTAs write a [F][L][T] complex64 tensor
// (the beam output of cfg 3) in different tile orders?  Pure stores, no compute.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

constexpr long long F = 1024, L = 2562, T = 938;

// mode 0: sequential grid-stride (memset-like)
// mode 1: GEMM-like: task = (f, m-tile of W frames), per step 64 rows x W frames; tasks round-robin over CTAs
// mode 2: full rows: task = (f, block of 41*64/R... rows), per step R full rows (R*7504 bytes contiguous)
__global__ void probe(float4* out, int mode, int W, int R) {
    const long long total16 = F * L * T / 2;       // float4 = 2 complex64
    if (mode == 0) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += (long long)gridDim.x * blockDim.x)
            out[i] = make_float4(1.f, 2.f, 3.f, 4.f);
        return;
    }
    if (mode == 1) {
        const int m_tiles = (int)((T + W - 1) / W);
        const long long tasks = F * m_tiles;
        const int n_tiles = (int)((L + 63) / 64);
        const int vec_per_row = W / 2;             // float4 per tile row
        for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
            const long long f = task / m_tiles;
            const int mt = (int)(task % m_tiles);
            for (int nt = 0; nt < n_tiles; ++nt) {
                for (int e = threadIdx.x; e < 64 * vec_per_row; e += blockDim.x) {
                    const int r = e / vec_per_row, c = e % vec_per_row;
                    const long long row = (long long)nt * 64 + r;
                    const long long t = (long long)mt * W + 2 * c;
                    if (row < L && t + 1 < T) out[((f * L + row) * T + t) / 2] = make_float4(1.f, 2.f, 3.f, 4.f);
                }
            }
        }
        return;
    }
    if (mode == 3) {
        // steering-stationary pattern: task = (f, block of 128 rows); per step 128 rows x W frames
        const long long blocks = (L + 127) / 128;
        const long long tasks = F * blocks;
        const int vec = W / 2;
        for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
            const long long f = task / blocks, r0 = (task % blocks) * 128;
            for (long long t0 = 0; t0 < T; t0 += W) {
                for (int e = threadIdx.x; e < 128 * vec; e += blockDim.x) {
                    const long long row = r0 + e / vec;
                    const long long t = t0 + 2 * (e % vec);
                    if (row < L && t + 1 < T) out[((f * L + row) * T + t) / 2] = make_float4(1.f, 2.f, 3.f, 4.f);
                }
            }
        }
        return;
    }
    if (mode == 4) {
        // GEMM-like tiles, but one CTA walks the m-tiles of its own bin one after the other
        const int m_tiles = (int)((T + W - 1) / W);
        const int n_tiles = (int)((L + 63) / 64);
        const int vec_per_row = W / 2;
        for (long long f = blockIdx.x; f < F; f += gridDim.x)
            for (int mt = 0; mt < m_tiles; ++mt)
                for (int nt = 0; nt < n_tiles; ++nt)
                    for (int e = threadIdx.x; e < 64 * vec_per_row; e += blockDim.x) {
                        const int r = e / vec_per_row, c = e % vec_per_row;
                        const long long row = (long long)nt * 64 + r;
                        const long long t = (long long)mt * W + 2 * c;
                        if (row < L && t + 1 < T) out[((f * L + row) * T + t) / 2] = make_float4(1.f, 2.f, 3.f, 4.f);
                    }
        return;
    }
    if (mode == 5) {
        // GEMM-like tiles, n-tile outer / m-tile inner inside one CTA (full rows complete every 8 tiles)
        const int m_tiles = (int)((T + W - 1) / W);
        const int n_tiles = (int)((L + 63) / 64);
        const int vec_per_row = W / 2;
        for (long long f = blockIdx.x; f < F; f += gridDim.x)
            for (int nt = 0; nt < n_tiles; ++nt)
                for (int mt = 0; mt < m_tiles; ++mt)
                    for (int e = threadIdx.x; e < 64 * vec_per_row; e += blockDim.x) {
                        const int r = e / vec_per_row, c = e % vec_per_row;
                        const long long row = (long long)nt * 64 + r;
                        const long long t = (long long)mt * W + 2 * c;
                        if (row < L && t + 1 < T) out[((f * L + row) * T + t) / 2] = make_float4(1.f, 2.f, 3.f, 4.f);
                    }
        return;
    }
    // mode 2: full rows, R rows per step; a task = (f, group of rows covering L / groups)
    const int groups = 8;                          // same number of tasks as mode 1 with 8 m-tiles
    const long long rows_per_task = (L + groups - 1) / groups;
    const long long tasks = F * groups;
    const int vec_per_row = (int)(T / 2);
    for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
        const long long f = task / groups;
        const long long r0 = (task % groups) * rows_per_task;
        for (long long rb = 0; rb < rows_per_task; rb += R) {
            for (int e = threadIdx.x; e < R * vec_per_row; e += blockDim.x) {
                const long long row = r0 + rb + e / vec_per_row;
                const int c = e % vec_per_row;
                if (row < L && row < r0 + rows_per_task) out[((f * L + row) * T) / 2 + c] = make_float4(1.f, 2.f, 3.f, 4.f);
            }
        }
    }
}

int main() {
    float4* out;
    const size_t bytes = (size_t)F * L * T * 8;
    cudaMalloc(&out, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    struct Cfg { int mode, W, R, grid; const char* name; };
    std::vector<Cfg> cfgs = {
        {0, 0, 0, 148 * 8, "sequential grid-stride"},
        {1, 128, 0, 148, "GEMM-like tiles 64 rows x 1 KB (8 m-tiles per bin on 8 CTAs)"},
        {1, 256, 0, 148, "GEMM-like tiles 64 rows x 2 KB"},
        {1, 512, 0, 148, "GEMM-like tiles 64 rows x 4 KB"},
        {1, 938, 0, 148, "GEMM-like tiles 64 rows x full row (1 CTA per bin)"},
        {3, 64, 0, 148, "steering-stationary: CTA owns 128 rows, 512 B per row per step"},
        {3, 128, 0, 148, "steering-stationary: CTA owns 128 rows, 1 KB per row per step"},
        {4, 128, 0, 148, "64x1KB tiles, one CTA per bin, m-tiles one after the other"},
        {5, 128, 0, 148, "64x1KB tiles, one CTA per bin, n-tile outer / m-tile inner"},
        {1, 128, 0, 74, "GEMM-like tiles 64 rows x 1 KB, 74 CTAs"},
        {2, 0, 8, 148, "full rows, 8 rows (60 KB contiguous) per step"},
        {2, 0, 64, 148, "full rows, 64 rows (480 KB contiguous) per step"},
    };
    for (auto& c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            probe<<<c.grid, 512>>>(out, c.mode, c.W, c.R);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
        }
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("%-72s %7.3f ms  %5.2f TB/s\n", c.name, ms, bytes / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
