// Microbenchmark: TMEM read rate of tcgen05.ld.32x32b.x16 on one SM -- the ceiling of the policy kernels' epilogues.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read tmem_read.cu && ./tmem_read
// Each of W warps (W = 4: one per TMEM lane quarter; W = 8: two per quarter) reads `cols` accumulator columns of its 32 lanes
// in batches of 16, `reps` times; mode 0: loads only (results xor-folded), mode 1: loads + the epilogue's arithmetic
// (bias add, LeakyReLU, bf16 pack) on every value, mode 2: the arithmetic alone on register values.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kMode>
__global__ void __launch_bounds__(256, 1) tmem_read_kernel(int warps, int cols, int reps, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sptr(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = threadIdx.x;
    float facc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < warps) {
        const int half = warps == 8 ? (warp >> 2) : 0, n_halves = warps == 8 ? 2 : 1;
        for (int r = 0; r < reps; ++r) {
            for (int c = half * 16; c < cols; c += 16 * n_halves) {
                uint32_t v[16];
                if (kMode != 2) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                        : "r"(t + c));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = acc + j * 0x01010101u + c;
                }
                if (kMode == 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc ^= v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const float x0 = __uint_as_float(v[j]) + 0.25f, x1 = __uint_as_float(v[j + 1]) + 0.5f;
                        const float y0 = fmaxf(x0, 0.01f * x0), y1 = fmaxf(x1, 0.01f * x1);
                        __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
                        acc ^= *reinterpret_cast<uint32_t*>(&h);
                    }
                }
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + (uint32_t)facc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int main() {
    long long* d_cycles;
    uint32_t* d_sink;
    cudaMalloc(&d_cycles, 8);
    cudaMalloc(&d_sink, 256 * 4);
    const int cols = 256, reps = 64;
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {4, 8}) {
            for (int it = 0; it < 2; ++it) {
                if (mode == 0) tmem_read_kernel<0><<<1, 256>>>(warps, cols, reps, d_cycles, d_sink);
                if (mode == 1) tmem_read_kernel<1><<<1, 256>>>(warps, cols, reps, d_cycles, d_sink);
                if (mode == 2) tmem_read_kernel<2><<<1, 256>>>(warps, cols, reps, d_cycles, d_sink);
                cudaDeviceSynchronize();
            }
            long long c = 0;
            cudaMemcpy(&c, d_cycles, 8, cudaMemcpyDeviceToHost);
            const double per_tile = (double)c / reps;  // cycles per 128 x 256 fp32 accumulator
            printf("mode %d (%s), %d warps: %.0f cycles per 128x256 accumulator = %.1f B/cycle, %.1f cycles per 16-column batch round\n", mode,
                   mode == 0 ? "loads only" : mode == 1 ? "loads + convert" : "convert only", warps, per_tile, 128.0 * 256 * 4 / per_tile,
                   per_tile / 16);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
