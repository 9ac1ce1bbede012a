// The MLP of the policy / value network on the encoder output: 64 -> 256 -> 160 -> 128 -> 2 (or 1), LeakyReLU, tanh.
//
// Second half of GaussianNeuralNetwork.compute (rover_envs/envs/navigation/learning/skrl/models.py:97-102; value network:
// :151-162) when the first half -- the heightmap encoder -- ran inside the height scan (fused_scan_policy.cu).  Its input
// is the encoder output [N, 64] bf16 = [e(60), obs[:, 0:4]] per environment, 128 B rows: a 128-row TMA box with
// SWIZZLE_128B lands as the canonical K-major A operand of the first MMA (K = 64), exactly like the bf16 observation mode
// of policy_ws.cu.  Layers and arithmetic are those of policy_ws.cu's layer group (same packed weights W2..W5, same
// epilogues, bf16 operands / fp32 accumulation on tcgen05), so the outputs match the unfused kernels bit for bit.
// One persistent CTA per SM: warp 4 prefetches the next tile's rows while warps 0-3 run the four layers of the current
// one.  The input is 128 B per environment, so the launch is bound by the layer chain's latency (~5 us per tile), not by
// memory: 16384 environments = 128 tiles = one tile per SM.
#include "policy_common.cuh"

namespace rover {

constexpr int kMlpThreads = 160;
constexpr int kMlpInStages = 2;
constexpr int kMlpInBytes = kTileM * 128;  // 128 rows x 64 bf16
constexpr int kMlpWBuf = 40 * 1024;

struct MlpSmem {
    unsigned char a_in[kMlpInStages][kMlpInBytes];  // TMA destinations (1024-byte aligned)
    unsigned char act[32 * kWsPlane];               // A operand of layers 3..5 (K <= 256): 64 KB
    unsigned char w[kMlpWBuf];
    float bias[256 + 160 + 128 + 16];
    unsigned long long in_full[kMlpInStages], in_empty[kMlpInStages];
    unsigned long long w_full, acc_done;
    uint32_t tmem_base;
};
static_assert(sizeof(MlpSmem) + 1024 <= 227 * 1024, "MlpSmem exceeds the shared memory of one SM");
static_assert(weight_bytes(2) <= kMlpWBuf && weight_bytes(3) == 2 * kMlpWBuf && weight_bytes(4) <= kMlpWBuf &&
                  weight_bytes(5) <= kMlpWBuf,
              "weight buffer plan");

__global__ void __launch_bounds__(kMlpThreads, 1)
policy_mlp_kernel(const __grid_constant__ CUtensorMap enc_map, int n_envs, const unsigned char* __restrict__ packed,
                  float* __restrict__ out, int value_head) {
    extern __shared__ unsigned char smem_dyn[];
    MlpSmem& sm = *reinterpret_cast<MlpSmem*>(smem_dyn + ((1024u - (sptr(smem_dyn) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (n_envs + kTileM - 1) / kTileM;

    if (tid == 4 * 32) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&enc_map)) : "memory");
        for (int i = 0; i < kMlpInStages; ++i) {
            mb_init(&sm.in_full[i], 1);
            mb_init(&sm.in_empty[i], 1);  // tcgen05.commit of the MMAs that read the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int)blockIdx.x < n_tiles) {
            mb_expect_tx(&sm.in_full[0], kMlpInBytes);
            tma_2d(sm.a_in[0], &enc_map, 0, blockIdx.x * kTileM, &sm.in_full[0]);
        }
    }
    if (tid == 0) {
        mb_init(&sm.w_full, 1);
        mb_init(&sm.acc_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sptr(&sm.tmem_base)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 4) {
        // =============================================================== TMA producer: one tile ahead
        if (lane == 0) {
            int i = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
                if (i == 0) continue;  // issued in the prologue
                const int s = i % kMlpInStages;
                if (i >= kMlpInStages) mb_wait(&sm.in_empty[s], (uint32_t)(i / kMlpInStages - 1) & 1u);
                mb_expect_tx(&sm.in_full[s], kMlpInBytes);
                tma_2d(sm.a_in[s], &enc_map, 0, tile * kTileM, &sm.in_full[s]);
            }
        }
    } else {
        // =============================================================== layer group (warps 0-3): layers 2..5
        const int row = tid;  // TMEM lane == tile row
        const uint32_t t_lane = (uint32_t)(warp * 32) << 16;
        const uint32_t acc = tmem;
        const uint32_t a0 = sptr(sm.act), b0 = sptr(sm.w);
        uint32_t ph_w = 0, ph_acc = 0;
        auto load_w = [&](int byte_offset, int bytes) {
            mb_expect_tx(&sm.w_full, (uint32_t)bytes);
            bulk_g2s(sm.w, packed + byte_offset, (uint32_t)bytes, &sm.w_full);
        };
        auto issue_mma = [&](int plane0, int k_steps, int nn, bool accumulate) {
            tc_fence_after();
            mb_wait(&sm.w_full, ph_w);
            const uint32_t idesc = make_idesc(nn);
            for (int j = 0; j < k_steps; ++j)
                umma(acc, make_desc(a0 + (plane0 + 2 * j) * kWsPlane, kWsPlane), make_desc(b0 + j * 2 * nn * 16, nn * 16), idesc,
                     accumulate || j != 0);
            umma_commit(&sm.acc_done);
        };
        auto wait_acc = [&]() {
            mb_wait(&sm.acc_done, ph_acc);
            ph_acc ^= 1u;
            tc_fence_after();
        };
        auto publish_act = [&]() {
            tc_fence_before();
            fence_async_smem();
            layer_group_sync();
        };
        if (tid == 0) load_w(weight_offset(2), weight_bytes(2));
        {
            constexpr int kFirst = (bias_offset(2) - kBiasOffset) / 4, kCount = 256 + 160 + 128 + 16;
            for (int k = tid; k < kCount; k += 128)
                sm.bias[k] = __ldg(reinterpret_cast<const float*>(packed + kBiasOffset) + kFirst + k);
            layer_group_sync();
        }
        const float* bias2 = sm.bias;
        const float* bias3 = bias2 + 256;
        const float* bias4 = bias3 + 160;
        const float* bias5 = bias4 + 128;
        int i = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
            const int s = i % kMlpInStages;
            const int grow = tile * kTileM + row;
            const bool live = grow < n_envs;
            // ---- layer 2: A = the encoder rows as they landed (K-major SWIZZLE_128B, K = 64), N = 256
            if (tid == 0) {
                mb_wait(&sm.in_full[s], (uint32_t)(i / kMlpInStages) & 1u);
                tc_fence_after();
                mb_wait(&sm.w_full, ph_w);
                const uint32_t idesc = make_idesc(layer_n(2));
                const uint32_t ain = sptr(sm.a_in[s]);
#pragma unroll
                for (int j = 0; j < 64 / 16; ++j)  // 16 bf16 = 32 B further along K inside the 128 B rows
                    umma(acc, make_desc_sw128(ain + j * 32), make_desc(b0 + j * 2 * layer_n(2) * 16, layer_n(2) * 16), idesc, j != 0);
                umma_commit(&sm.in_empty[s]);
                umma_commit(&sm.acc_done);
            }
            ph_w ^= 1u;
            wait_acc();
            if (tid == 0) load_w(weight_offset(3), kMlpWBuf);  // W3, K planes 0..15
            // ---- layer 2 epilogue -> A3 (256 columns); layer 3 in two K halves through the 40 KB weight buffer
            epilogue_to_act(acc + t_lane, layer_n(2), bias2, sm.act, row, nullptr);
            publish_act();
            if (tid == 0) issue_mma(0, 8, layer_n(3), false);
            ph_w ^= 1u;
            wait_acc();
            if (tid == 0) {
                load_w(weight_offset(3) + kMlpWBuf, kMlpWBuf);  // W3, K planes 16..31
                issue_mma(16, 8, layer_n(3), true);
            }
            ph_w ^= 1u;
            wait_acc();
            if (tid == 0) load_w(weight_offset(4), weight_bytes(4));
            // ---- layer 3 epilogue -> A4
            epilogue_to_act(acc + t_lane, layer_n(3), bias3, sm.act, row, nullptr);
            publish_act();
            if (tid == 0) issue_mma(0, layer_k(4) / 16, layer_n(4), false);
            ph_w ^= 1u;
            wait_acc();
            if (tid == 0) load_w(weight_offset(5), weight_bytes(5));
            // ---- layer 4 epilogue -> A5
            epilogue_to_act(acc + t_lane, layer_n(4), bias4, sm.act, row, nullptr);
            publish_act();
            if (tid == 0) issue_mma(0, layer_k(5) / 16, layer_n(5), false);
            ph_w ^= 1u;
            wait_acc();
            if (tid == 0 && tile + (int)gridDim.x < n_tiles) load_w(weight_offset(2), weight_bytes(2));  // next tile's W2
            // ---- last layer: mean = tanh(D5 + b5), two real columns (value head: one linear output)
            {
                float v[8];
                tmem_ld8(acc + t_lane, v);
                if (live && value_head) {
                    out[grow] = v[0] + bias5[0];
                } else if (live) {
                    float2 m;
                    m.x = tanhf(v[0] + bias5[0]);
                    m.y = tanhf(v[1] + bias5[1]);
                    *reinterpret_cast<float2*>(out + 2 * (size_t)grow) = m;
                }
            }
            tc_fence_before();
            layer_group_sync();  // every TMEM read of this tile is done before the next tile's first MMA overwrites acc
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    }
}

typedef CUresult (*MlpEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace rover

extern "C" int rover_policy_mlp_forward(const uint16_t* enc_bf16, int32_t n_envs, const void* packed, float* out,
                                        int32_t value_head, void* stream) {
    using namespace rover;
    ROVER_CHECK(n_envs >= 0, "rover_policy_mlp_forward: negative n_envs");
    if (n_envs == 0) return 0;
    ROVER_CHECK(enc_bf16 && packed && out, "rover_policy_mlp_forward: NULL argument");
    ROVER_CHECK((reinterpret_cast<uintptr_t>(enc_bf16) & 127) == 0, "rover_policy_mlp_forward: encoder rows must be 128-byte aligned");
    ROVER_CHECK((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "rover_policy_mlp_forward: packed blob not 128B aligned");
    static MlpEncodeTiledFn encode = nullptr;
    static int n_sms = 0;
    constexpr int kSmemBytes = (int)sizeof(MlpSmem) + 1024;
    if (!encode) {
        int dev = 0;
        ROVER_CUDA(cudaGetDevice(&dev));
        ROVER_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
        ROVER_CUDA(cudaFuncSetAttribute(policy_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        ROVER_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        ROVER_CHECK(fn && q == cudaDriverEntryPointSuccess, "rover_policy_mlp_forward: cuTensorMapEncodeTiled unavailable");
        encode = reinterpret_cast<MlpEncodeTiledFn>(fn);
    }
    alignas(64) CUtensorMap map;
    const cuuint64_t gdim[2] = {64ull, (cuuint64_t)n_envs};
    const cuuint64_t gstride[1] = {128ull};
    const cuuint32_t box[2] = {64u, (cuuint32_t)kTileM};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(enc_bf16), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ROVER_CHECK(rc == CUDA_SUCCESS, "rover_policy_mlp_forward: cuTensorMapEncodeTiled failed (%d)", (int)rc);
    const int n_tiles = (n_envs + kTileM - 1) / kTileM;
    const int grid = n_tiles < n_sms ? n_tiles : n_sms;
    policy_mlp_kernel<<<grid, kMlpThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(
        map, n_envs, static_cast<const unsigned char*>(packed), out, value_head ? 1 : 0);
    return check_launch("policy_mlp_kernel");
}
