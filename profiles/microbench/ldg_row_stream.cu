// Microbenchmark: the observation matrix [65536 x 968] fp32 streamed through the load/store path the way converter threads
// would read it (thread = row, 64-byte pieces of a 128-byte chunk per thread, several chunks in flight), against the TMA
// figures of tma_box_throughput.cu (4.3-4.6 TB/s cold whatever the box).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldg_row_stream ldg_row_stream.cu && ./ldg_row_stream
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int kAhead>
__global__ void __launch_bounds__(256, 1) stream_kernel(const float* __restrict__ obs, int n_rows, int stride, unsigned* sink) {
    const int row = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int n_tiles = n_rows / 128;
    unsigned acc = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const float4* p = reinterpret_cast<const float4*>(obs + (size_t)(t * 128 + row) * stride + 16 * half);
#pragma unroll 1
        for (int c0 = 0; c0 < 30; c0 += kAhead) {
            float4 v[kAhead][4];
#pragma unroll
            for (int a = 0; a < kAhead; ++a)
#pragma unroll
                for (int k = 0; k < 4; ++k) v[a][k] = __ldcs(p + (c0 + a) * 8 + k);  // 32 columns per chunk = 8 float4
#pragma unroll
            for (int a = 0; a < kAhead; ++a)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc ^= __float_as_uint(v[a][k].x) ^ __float_as_uint(v[a][k].w);
        }
    }
    if (acc == 0x12345u) *sink = acc;
}

int main() {
    const int n_rows = 65536, stride = 968;
    float* d;
    cudaMalloc(&d, (size_t)n_rows * stride * 4);
    cudaMemset(d, 0, (size_t)n_rows * stride * 4);
    unsigned char* flush;
    cudaMalloc(&flush, 256u << 20);
    unsigned* sink;
    cudaMalloc(&sink, 4);
    for (int warm = 0; warm < 2; ++warm) {
        const int rows = warm ? 8192 : n_rows;
        for (int ahead : {1, 2, 3, 5}) {
            float best = 1e9f;
            for (int rep = 0; rep < 6; ++rep) {
                if (!warm) cudaMemset(flush, rep, 256u << 20);
                cudaEvent_t a, b;
                cudaEventCreate(&a), cudaEventCreate(&b);
                cudaEventRecord(a);
                if (ahead == 1) stream_kernel<1><<<148, 256>>>(d, rows, stride, sink);
                if (ahead == 2) stream_kernel<2><<<148, 256>>>(d, rows, stride, sink);
                if (ahead == 3) stream_kernel<3><<<148, 256>>>(d, rows, stride, sink);
                if (ahead == 5) stream_kernel<5><<<148, 256>>>(d, rows, stride, sink);
                cudaEventRecord(b);
                cudaDeviceSynchronize();
                float ms;
                cudaEventElapsedTime(&ms, a, b);
                if (rep > 0 && ms < best) best = ms;
            }
            const double bytes = (double)rows * 960 * 4;
            printf("%s, %d chunk(s) in flight per thread (%d KB per SM): %7.1f us  %6.0f GB/s\n", warm ? "L2-resident (31 MB)" : "cold (254 MB)",
                   ahead, ahead * 256 * 64 / 1024, best * 1e3, bytes / (best * 1e-3) / 1e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
