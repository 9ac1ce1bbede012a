// Shapes, packed-weight layout and PTX wrappers shared by the policy kernels (policy.cu: tile-serial v1;
// policy_ws.cu: warp-specialised v2).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace rover {

constexpr int kPolThreads = 256;
constexpr int kTileM = 128;
constexpr int kObsCols = 965;
constexpr int kEncIn = 961;       // obs[:, 3:964]
constexpr int kEncInOffset = 3;
constexpr int kChunkK = 64;
constexpr int kNumChunks = 16;    // observation columns [0, 1024) in chunks of 64; columns outside [3, 964) are zeroed
constexpr int kAPlane = kTileM * 16 + 16;  // A plane stride (bytes), +16 B skew: conflict-free 16 B stores per plane
constexpr int kNumLayers = 6;
constexpr int kStagesF = 4;  // fp32 observation stages (prefetch distance 3 chunks)
constexpr int kStagesW = 3;  // layer-0 weight-chunk stages (prefetch distance 2 chunks)

// padded layer shapes: K (multiple of 16), N (multiple of 16)
__host__ __device__ constexpr int layer_k(int l) { return l == 0 ? 1024 : l == 1 ? 80 : l == 2 ? 64 : l == 3 ? 256 : l == 4 ? 160 : 128; }
__host__ __device__ constexpr int layer_n(int l) { return l == 0 ? 80 : l == 1 ? 64 : l == 2 ? 256 : l == 3 ? 160 : l == 4 ? 128 : 16; }
__host__ __device__ constexpr int layer_k_real(int l) { return l == 0 ? 961 : l == 1 ? 80 : l == 2 ? 64 : l == 3 ? 256 : l == 4 ? 160 : 128; }
__host__ __device__ constexpr int layer_n_real(int l) { return l == 0 ? 80 : l == 1 ? 60 : l == 2 ? 256 : l == 3 ? 160 : l == 4 ? 128 : 2; }

// packed blob: [W0: 16 chunks x (8 planes x 80 rows x 16 B)] [W1..W5: (K/8 planes x N rows x 16 B)] [biases fp32]
constexpr int kW0ChunkBytes = 8 * 80 * 16;  // 10,240
__host__ __device__ constexpr int weight_bytes(int l) { return l == 0 ? kNumChunks * kW0ChunkBytes : (layer_k(l) / 8) * layer_n(l) * 16; }
__host__ __device__ constexpr int weight_offset(int l) {
    int off = 0;
    for (int i = 0; i < l; ++i) off += weight_bytes(i);
    return off;
}
constexpr int kBiasOffset = weight_offset(kNumLayers);
__host__ __device__ constexpr int bias_offset(int l) {
    int off = kBiasOffset;
    for (int i = 0; i < l; ++i) off += layer_n(i) * 4;
    return off;
}
constexpr int kPackedBytes = bias_offset(kNumLayers);

// ---------------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t sptr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(unsigned long long* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(sptr(b)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sptr(dst)),
                 "l"(src), "r"(bytes), "r"(sptr(b))
                 : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* b) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(sptr(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(sptr(b))
        : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: LBO = byte stride between the two 8-element K chunks of one
// MMA (= plane stride), SBO = byte stride between 8-row groups (= 128 B), version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor, kind::f16: D = F32, A = B = BF16, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sptr(b)) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.01f * x; }

}  // namespace rover
