// Micro-benchmark (bring-up evidence, not part of the library): per-SM throughput of 2-D TMA tensor-map loads of a
// row-major fp32 matrix [65536 x 968] as a function of the box shape, ring depth fixed to ~80 KB of loads in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_box profiles/microbench/tma_box_throughput.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e = (x);                                                         \
        if (e != cudaSuccess) {                                                      \
            printf("%s: %s\n", #x, cudaGetErrorString(e));                           \
            exit(1);                                                                 \
        }                                                                            \
    } while (0)

__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sptr(b)) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned long long* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(sptr(b)), "r"(parity)
                     : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* b) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     sptr(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(sptr(b))
                 : "memory");
}

constexpr int kMaxStages = 16;

// every CTA streams row tiles of `tile_rows` rows (tile t = blockIdx.x + k * gridDim.x), `n_cols` columns in boxes
__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap map, int n_rows, int n_cols, int box_w, int box_h, int tile_rows,
              int stages, int stage_bytes, unsigned long long* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long full[kMaxStages], empty[kMaxStages];
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) {
            mb_init(&full[i], 1);
            mb_init(&empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_tiles = n_rows / tile_rows;
    if (threadIdx.x == 0) {
        int s = 0;
        uint32_t p = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
            for (int x = 0; x < n_cols; x += box_w)
                for (int y = 0; y < tile_rows; y += box_h) {
                    mb_wait(&empty[s], (p & 1u) ^ 1u);
                    mb_expect_tx(&full[s], (uint32_t)stage_bytes);
                    tma_2d(smem + (size_t)s * stage_bytes, &map, x, t * tile_rows + y, &full[s]);
                    if (++s == stages) s = 0, ++p;
                }
    } else if (threadIdx.x == 32) {
        int s = 0;
        uint32_t p = 0;
        unsigned long long acc = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
            for (int x = 0; x < n_cols; x += box_w)
                for (int y = 0; y < tile_rows; y += box_h) {
                    mb_wait(&full[s], p & 1u);
                    acc += *reinterpret_cast<const unsigned*>(smem + (size_t)s * stage_bytes);
                    mb_arrive(&empty[s]);
                    if (++s == stages) s = 0, ++p;
                }
        if (acc == 0x1234567ull) *sink = acc;
    }
}

int main() {
    const int n_rows = 65536, stride = 968, n_cols_real = 965;
    float* d;
    CK(cudaMalloc(&d, (size_t)n_rows * stride * 4));
    CK(cudaMemset(d, 0, (size_t)n_rows * stride * 4));
    unsigned char* flush;
    CK(cudaMalloc(&flush, 256u << 20));
    unsigned long long* sink;
    CK(cudaMalloc(&sink, 8));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Cfg {
        int w, h, stages;
        CUtensorMapSwizzle sw;
        const char* name;
        CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    } cfgs[] = {
        {32, 128, 5, CU_TENSOR_MAP_SWIZZLE_128B, "32 x 128 rows, swizzle128, 5 stages, L2 promotion 256B", CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {32, 128, 5, CU_TENSOR_MAP_SWIZZLE_128B, "32 x 128 rows, swizzle128, 5 stages, L2 promotion none", CU_TENSOR_MAP_L2_PROMOTION_NONE},
        {32, 112, 5, CU_TENSOR_MAP_SWIZZLE_128B, "32 x 112 rows (tile 112), 5 stages, L2 promotion 256B", CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
        {32, 128, 5, CU_TENSOR_MAP_SWIZZLE_128B, "32 x 128 rows, swizzle128, 5 stages (policy v2)"},
        {32, 128, 10, CU_TENSOR_MAP_SWIZZLE_128B, "32 x 128 rows, swizzle128, 10 stages"},
        {64, 128, 3, CU_TENSOR_MAP_SWIZZLE_NONE, "64 x 128 rows, 3 stages (policy v1)"},
        {64, 128, 5, CU_TENSOR_MAP_SWIZZLE_NONE, "64 x 128 rows, 5 stages"},
        {64, 64, 5, CU_TENSOR_MAP_SWIZZLE_NONE, "64 x 64 rows, 5 stages"},
        {128, 32, 5, CU_TENSOR_MAP_SWIZZLE_NONE, "128 x 32 rows, 5 stages"},
        {256, 16, 5, CU_TENSOR_MAP_SWIZZLE_NONE, "256 x 16 rows, 5 stages"},
        {256, 16, 10, CU_TENSOR_MAP_SWIZZLE_NONE, "256 x 16 rows, 10 stages"},
        {256, 32, 5, CU_TENSOR_MAP_SWIZZLE_NONE, "256 x 32 rows, 5 stages"},
        {32, 32, 16, CU_TENSOR_MAP_SWIZZLE_128B, "32 x 32 rows, swizzle128, 16 stages"},
    };
    const bool warm = getenv("TMA_WARM") != nullptr;  // TMA_WARM=1: 8192 rows (31 MB, L2-resident), no flush, repeated
    const int n_rows_run = warm ? 8192 : n_rows;
    for (const Cfg& c : cfgs) {
        CUtensorMap map;
        const cuuint64_t gdim[2] = {(cuuint64_t)n_cols_real, (cuuint64_t)n_rows};
        const cuuint64_t gstride[1] = {(cuuint64_t)stride * 4ull};
        const cuuint32_t box[2] = {(cuuint32_t)c.w, (cuuint32_t)c.h};
        const cuuint32_t estr[2] = {1u, 1u};
        CUresult rc = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstride, box, estr,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, c.promo,
                                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) {
            printf("%s: encode failed %d\n", c.name, (int)rc);
            continue;
        }
        const int stage_bytes = c.w * c.h * 4;
        const int n_cols = ((992 + c.w - 1) / c.w) * c.w;  // like the policy kernel: columns [0, 992)
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            if (!warm) CK(cudaMemset(flush, rep, 256u << 20));
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a);
            stream_kernel<<<148, 64, c.stages * stage_bytes>>>(map, n_rows_run, n_cols, c.w, c.h, c.h == 112 ? 112 : 128, c.stages, stage_bytes, sink);
            cudaEventRecord(b);
            CK(cudaDeviceSynchronize());
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (rep > 0 && ms < best) best = ms;
        }
        const double bytes = (double)n_rows_run * n_cols_real * 4;
        printf("%-52s in flight %6.1f KB/SM : %7.1f us  %6.0f GB/s (real bytes)\n", c.name, c.stages * stage_bytes / 1024.0,
               best * 1e3, bytes / (best * 1e-3) / 1e9);
    }
    return 0;
}
