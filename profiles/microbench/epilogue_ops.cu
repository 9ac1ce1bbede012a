// Microbenchmark: issue cost of the epilogue's instructions on one SM (cycles per warp-instruction per scheduler).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epilogue_ops epilogue_ops.cu && ./epilogue_ops
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int kOp>
__global__ void __launch_bounds__(512, 1) op_kernel(int iters, long long* cycles, uint32_t* sink, float seed) {
    float a[8];
    uint32_t u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x + j, u[j] = threadIdx.x * 77u + j;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (kOp == 0) a[j] = fmaxf(a[j], a[(j + 1) & 7] * 0.5f) ;                    // FMUL + FMNMX
            if (kOp == 1) a[j] = a[j] * 1.0001f;                                          // FMUL only
            if (kOp == 2) {                                                               // F2FP.BF16.PACK_AB (+ a MOV-free chain)
                __nv_bfloat162 h = __floats2bfloat162_rn(a[j], a[(j + 1) & 7]);
                a[j] = __uint_as_float(*reinterpret_cast<uint32_t*>(&h));
            }
            if (kOp == 3) u[j] = (u[j] + 0x7fffu + ((u[(j + 1) & 7] >> 16) & 1u));        // integer RNE increment
            if (kOp == 4) u[j] = __byte_perm(u[j], u[(j + 1) & 7], 0x7632);              // PRMT
            if (kOp == 5) a[j] = fmaxf(a[j], a[(j + 1) & 7]);                             // FMNMX only
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s ^= __float_as_uint(a[j]) ^ u[j];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
    long long* d_cycles;
    uint32_t* d_sink;
    cudaMalloc(&d_cycles, 8);
    cudaMalloc(&d_sink, 512 * 4);
    const int iters = 4096;
    const char* names[6] = {"FMUL + FMNMX", "FMUL", "F2FP.BF16.PACK_AB", "int RNE increment (SHF/LOP3 + IADD3)", "PRMT", "FMNMX"};
    for (int op = 0; op < 6; ++op)
        for (int threads : {128, 256, 512}) {
            for (int it = 0; it < 2; ++it) {
                switch (op) {
                    case 0: op_kernel<0><<<1, threads>>>(iters, d_cycles, d_sink, 1.f); break;
                    case 1: op_kernel<1><<<1, threads>>>(iters, d_cycles, d_sink, 1.f); break;
                    case 2: op_kernel<2><<<1, threads>>>(iters, d_cycles, d_sink, 1.f); break;
                    case 3: op_kernel<3><<<1, threads>>>(iters, d_cycles, d_sink, 1.f); break;
                    case 4: op_kernel<4><<<1, threads>>>(iters, d_cycles, d_sink, 1.f); break;
                    case 5: op_kernel<5><<<1, threads>>>(iters, d_cycles, d_sink, 1.f); break;
                }
                cudaDeviceSynchronize();
            }
            long long c = 0;
            cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
            const double per = (double)c / (iters * 8.0) / (threads / 128.0);  // cycles per source-level op per warp per scheduler
            printf("%-40s %3d threads: %.2f cycles per op per warp-slot (%.1f thread-ops/cycle/SM)\n", names[op], threads, per,
                   (double)threads * iters * 8 / c);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
