// HBM write-pattern probe 2: GEMM-like tiles (64 rows x W frames, the m-tiles of one bin on
// different CTAs) for different ROW PITCHES of the output tensor.
#include <cstdio>
#include <cuda_runtime.h>

constexpr long long F = 1024, L = 2562, T = 938;

template <int W>
__global__ void __launch_bounds__(1024) probe(float4* out, long long pitch /* complex per row */) {
    const int m_tiles = (int)((T + W - 1) / W);
    const long long tasks = F * m_tiles;
    const int n_tiles = (int)((L + 63) / 64);
    constexpr int VPR = W / 2;                       // float4 per tile row
    for (long long task = blockIdx.x; task < tasks; task += gridDim.x) {
        const long long f = task / m_tiles;
        const int mt = (int)(task % m_tiles);
        for (int nt = 0; nt < n_tiles; ++nt) {
            for (int e = threadIdx.x; e < 64 * VPR; e += blockDim.x) {
                const int r = e / VPR, c = e % VPR;
                const long long row = (long long)nt * 64 + r;
                const long long t = (long long)mt * W + 2 * c;
                if (row < L && t + 1 < T) out[((f * L + row) * pitch + t) / 2] = make_float4(1.f, 2.f, 3.f, 4.f);
            }
        }
    }
}

int main() {
    float4* out;
    const size_t bytes_alloc = (size_t)F * L * 1280 * 8;
    const size_t bytes = (size_t)F * L * T * 8;
    cudaMalloc(&out, bytes_alloc);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const long long pitches[] = {938, 940, 944, 960, 992, 1024, 1040, 1056, 1088, 1152, 1280};
    for (long long pitch : pitches) {
        float ms1 = 0, ms2 = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            probe<128><<<148, 1024>>>(out, pitch);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms1, a, b);
            cudaEventRecord(a);
            probe<256><<<148, 1024>>>(out, pitch);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms2, a, b);
        }
        printf("pitch %5lld frames (%6lld B): 64x1KB tiles %6.3f ms %5.2f TB/s | 64x2KB tiles %6.3f ms %5.2f TB/s\n", pitch,
               pitch * 8, ms1, bytes / ms1 / 1e9, ms2, bytes / ms2 / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
