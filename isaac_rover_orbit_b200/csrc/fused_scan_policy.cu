// Height scan fused with the policy forward: the observation never leaves the SM.
//
// Replaces, in ONE launch, the pair the reference runs back to back every step (reference root relative):
//   height_scan_rover                rover_envs/envs/navigation/mdp/observations.py:35-45  (+ ORBIT RayCaster / warp raycast)
//   GaussianNeuralNetwork.compute    rover_envs/envs/navigation/learning/skrl/models.py:24-36, 89-102
// (BASELINE.json configs[3]: "policy forward ... fused with the observation kernel").  Unfused, the 961 heights of an
// environment make an HBM round trip -- 3844 B written by the scan (+ 1930 B for the bf16 mirror), 3860 / 1930 B read
// back by the policy -- and the stand-alone forward is HBM-bound at 83 FLOP/B.  Here the scan's consumer warps write
// each height as bf16 straight into the shared-memory operand of the layer-0 MMA.
//
// Orientation.  The scan is environment-major (one environment = 961 rays at a time), so a 128-environment A tile as in
// policy_ws.cu (247 KB of bf16) cannot be collected on chip.  Every layer therefore runs TRANSPOSED on tcgen05:
//     D_l^T [features (M = 128 lanes) x envs (N = 16)] = W_l [features x K] * A_l^T [K x envs]
// with the weights as the A operand (K-major, streamed from L2 through a ring of 10 KB pieces by bulk copies) and a batch
// of 16 environments as the B operand (K-major rows of 16 B core matrices: the observation operand the scan fills, then the
// bf16 activations the epilogues write).  M is always 128: layers with fewer output rows let the MMA read past their
// rows into whatever follows in the ring (finite garbage into TMEM lanes nobody reads).
//
// Roles (one persistent CTA per SM, 20 warps = 5 warpgroups).  The CTA is launched with 96 registers per thread (640 x 96 =
// 61,440: setmaxnreg moves registers inside THAT pool, not inside the SM's file) and re-balances them per warpgroup: 120
// for the three consumer groups, 64 for the producer / streamer / issuer group, 48 for the epilogue group (60,416 in all):
//   warp 0       scan producer: poses -> window -> 3-D tensor-map TMA load per environment (4-stage ring), as variant 5
//   warp 1       weight streamer: the 37 weight pieces of a batch, in order, through a 3-stage ring
//   warp 2       MMA issuer: layer 0 when a batch of 16 environments is complete, then layers 1-5 as the epilogues publish
//   warps 4-15   scan consumers (variant 5's packed-fp32 ray pairs; 256-ray chunks handed out dynamically, in order)
//   warps 16-19  epilogue group: TMEM lane = output feature; + bias, LeakyReLU, bf16 -> the next layer's operand; tanh.
// The policy work of batch b runs while the consumers scan batch b+1 (two observation operands); only the last batch's
// layers (a few microseconds) are exposed at the end of the launch.
#include "policy_common.cuh"
#include "scan_paired.cuh"

namespace rover {

constexpr int kFuConsumerWarps = 12;
constexpr int kFuWarps = 20;
constexpr int kFuThreads = 32 * kFuWarps;
constexpr int kFuStages = 4, kFuFullBars = 8;  // scan window ring; two `full` barriers per stage (phase aliasing, see below)
constexpr int kFuBatch = 16;                   // environments per batch = N of every MMA
constexpr int kFuRays = 961, kFuChunks = 4;    // 961 rays = 3 x 256 + 193
constexpr int kFuK0 = 976;                     // layer-0 K: observation columns [0, 976); weights are zero outside [3, 964)
constexpr int kFuObsBytes = (kFuK0 / 8) * kOperandLbo;   // 33,184
constexpr int kFuActBytes = 32 * kOperandLbo;            // activations, K <= 256
constexpr int kFuWStage = 10240, kFuWStages = 3, kFuWSlack = 1024;
constexpr int kFuMaxPieces = 40;
constexpr int kFuTmemCols = 128;
constexpr int kFuBiasFloats = 80 + 64 + 256 + 160 + 128 + 16;
constexpr int kFuRegsLaunch = 96, kFuRegsConsumer = 120, kFuRegsUtil = 64, kFuRegsEpilogue = 48;
static_assert(128 * (3 * kFuRegsConsumer + kFuRegsUtil + kFuRegsEpilogue) <= kFuThreads * kFuRegsLaunch,
              "setmaxnreg: the re-balanced budgets must fit the pool the CTA was launched with");
// Phase aliasing (scan ring): environment E signals `full` barrier E % 8 with parity (E / 8) & 1; a waiter is fooled only if
// E - 8 has not completed yet.  Chunks are handed out in order and at most 12 are outstanding, so when a warp holds a
// chunk of E, at least 20 of the 32 chunks of E-8 .. E-1 are finished, hence one of E-4 .. E-1 -- whose load was issued
// only after E-8 .. E-5 had been consumed (4 stages, in-order producer).
static_assert(kFuFullBars == 2 * kFuStages && 8 * kFuChunks - kFuConsumerWarps > 4 * kFuChunks, "phase aliasing argument");

struct FuPiece {
    uint32_t gofs;       // byte offset inside the packed image
    uint16_t bytes;      // size of the piece
    uint16_t lbo;        // plane stride of the A operand = padded rows * 16
    uint8_t layer, nk;   // nk MMAs of K = 16
    uint8_t kstep0;      // first K step (units of 16) inside the layer's B operand
    uint8_t tcol;        // TMEM column of the accumulator
    uint8_t first;       // first piece of its accumulator (first MMA overwrites)
    uint8_t last;        // last piece of the layer: commit to the epilogue
    uint8_t row0;        // first output feature of the piece's block (0 or 128)
    uint8_t pad;
};
struct FuPlan {
    FuPiece p[kFuMaxPieces];
    int n, bias_ofs, total_bytes, pad;
};

// The packed image: every piece is the A operand as the MMA reads it -- [K / 8 planes][rows_padded][8 bf16] -- followed by
// the biases (fp32, padded layer widths 80, 64, 256, 160, 128, 16).
struct FuLayerBlock {
    int layer, row0, rows_real, rows_pad, k_len, k_piece, tcol;
};
inline FuPlan make_fused_plan() {
    FuPlan plan{};
    const FuLayerBlock blocks[] = {
        {0, 0, 80, 80, kFuK0, 64, 0},    {1, 0, 60, 64, 80, 80, 16},    {2, 0, 128, 128, 64, 32, 32},
        {2, 128, 128, 128, 64, 32, 48},  {3, 0, 128, 128, 256, 32, 64}, {3, 128, 32, 32, 256, 128, 80},
        {4, 0, 128, 128, 160, 32, 96},   {5, 0, 2, 8, 128, 128, 112},
    };
    int n = 0, ofs = 0;
    const int n_blocks = (int)(sizeof(blocks) / sizeof(blocks[0]));
    for (int b = 0; b < n_blocks; ++b) {
        const FuLayerBlock& B = blocks[b];
        for (int k = 0; k < B.k_len; k += B.k_piece) {
            const int len = (B.k_len - k) < B.k_piece ? (B.k_len - k) : B.k_piece;
            FuPiece& P = plan.p[n++];
            P.gofs = (uint32_t)ofs;
            P.bytes = (uint16_t)((len / 8) * B.rows_pad * 16);
            P.lbo = (uint16_t)(B.rows_pad * 16);
            P.layer = (uint8_t)B.layer;
            P.nk = (uint8_t)(len / 16);
            P.kstep0 = (uint8_t)(k / 16);
            P.tcol = (uint8_t)B.tcol;
            P.first = k == 0;
            P.last = (k + len == B.k_len) && (b + 1 == n_blocks || blocks[b + 1].layer != B.layer);
            P.row0 = (uint8_t)B.row0;
            ofs += P.bytes;
        }
    }
    plan.n = n;
    plan.bias_ofs = ofs;
    plan.total_bytes = ofs + kFuBiasFloats * 4;
    return plan;
}

struct FuPackArgs {
    const float* w[6];
    const float* b[6];
    int in_dim[6], out_dim[6];
};

// one thread per bf16 element of a piece
__global__ void fused_pack_kernel(const __grid_constant__ FuPackArgs a, const __grid_constant__ FuPlan plan,
                                  unsigned char* __restrict__ packed) {
    for (int pi = blockIdx.y; pi < plan.n; pi += gridDim.y) {
        const FuPiece P = plan.p[pi];
        const int rows_pad = P.lbo / 16;
        const int total = (P.bytes / 2);
        const int row0 = P.row0;
        const int l = P.layer;
        const int k_real = layer_k_real(l), n_real = a.out_dim[l];
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(packed + P.gofs);
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
            const int j = idx & 7;
            const int r = (idx >> 3) % rows_pad;
            const int plane = (idx >> 3) / rows_pad;
            const int k = P.kstep0 * 16 + plane * 8 + j;
            int src_k = k;
            if (l == 0) src_k = k - kEncInOffset;           // layer-0 K index = observation column
            if (l == 2) src_k = (k < 60) ? k + 4 : k - 60;  // operand order [e(60), obs[:, 0:4]] -> reference [x(4), e(60)]
            const int n = row0 + r;
            float v = 0.f;
            if (n < n_real && src_k >= 0 && src_k < k_real) v = a.w[l][(size_t)n * a.in_dim[l] + src_k];
            dst[idx] = __float2bfloat16_rn(v);
        }
    }
    if (blockIdx.y == 0) {
        float* bdst = reinterpret_cast<float*>(packed + plan.bias_ofs);
        int ofs = 0;
        for (int l = 0; l < 6; ++l) {
            for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < layer_n(l); n += gridDim.x * blockDim.x)
                bdst[ofs + n] = n < a.out_dim[l] ? a.b[l][n] : 0.f;
            ofs += layer_n(l);
        }
    }
}

struct __align__(128) FuSmem {
    PairStage stage[kFuStages];
    float vx[kPairMaxRays], vy[kPairMaxRays];
    LinePair2 xpair[kPairMaxLines], ypair[kPairMaxLines];
    unsigned char obs[2][kFuObsBytes];   // observation operands of two batches: [k / 8][16 env rows x 16 B (+ 16 B skew)]
    unsigned char act[kFuActBytes];      // activation operand of the layer being fed
    unsigned char w[kFuWStages * kFuWStage + kFuWSlack];
    float bias[kFuBiasFloats];
    unsigned long long full_bar[kFuFullBars], empty_bar[kFuStages];
    unsigned long long obs_full[2], obs_empty[2];
    unsigned long long w_full[kFuWStages], w_empty[kFuWStages];
    unsigned long long acc_full, act_ready, batch_done;
    uint32_t tmem_base;
    int next_chunk;
    float vz0;
    __nv_bfloat16 head[kFuBatch][4];  // observation head of the batch in the layer pipeline (kept across its layer 0 -> 1)
};
static_assert(sizeof(FuSmem) <= 227 * 1024, "FuSmem exceeds the shared memory of one SM");
__device__ __forceinline__ float sm_vz(const FuSmem& sm, int) { return sm.vz0; }

__device__ __forceinline__ void fu_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sptr(b)) : "memory");
}
__device__ __forceinline__ void fu_epi_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void fu_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int kRegs>
__device__ __forceinline__ void reg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void reg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}
// element (env slot e, k) of a K-major operand with kOperandLbo plane stride
__device__ __forceinline__ unsigned char* operand_at(unsigned char* base, int e, int k) {
    return base + (k >> 3) * kOperandLbo + (e >> 3) * 128 + (e & 7) * 16 + (k & 7) * 2;
}

// Epilogue of one accumulator block: feature `feat` (this thread's TMEM lane + row offset of the block), 16 envs.
// LeakyReLU(d + bias) -> act[k = feat] as bf16; `inject` (the observation head) replaces the value where given.
__device__ __forceinline__ void fu_epilogue_block(uint32_t taddr, const float bias, unsigned char* act, int feat, bool store,
                                                  const __nv_bfloat16* inject /* [e * 4], or nullptr */) {
    uint32_t r[16];
    fu_tmem_ld16(taddr, r);
    if (store) {
#pragma unroll
        for (int e = 0; e < kFuBatch; ++e) {
            __nv_bfloat16 v = __float2bfloat16_rn(leaky(__uint_as_float(r[e]) + bias));
            if (inject != nullptr) v = inject[4 * e];
            *reinterpret_cast<__nv_bfloat16*>(operand_at(act, e, feat)) = v;
        }
    }
}

template <bool kWriteObs>
__global__ void __launch_bounds__(kFuThreads, 1)
fused_scan_policy_kernel(const float* __restrict__ pos_w, const float* __restrict__ quat_w, int n_envs,
                         const float* __restrict__ ray_local, const __grid_constant__ ScanGridDev g,
                         const __grid_constant__ PlaneCellsDev pc, const __grid_constant__ CUtensorMap tmap,
                         float pattern_radius, float max_d, float base_offset, float* __restrict__ obs, int obs_stride,
                         const unsigned char* __restrict__ packed, const __grid_constant__ FuPlan plan,
                         float* __restrict__ mean, int value_head) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FuSmem& sm = *reinterpret_cast<FuSmem*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_iter = (n_envs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // envs of this CTA
    const int n_batches = (n_iter + kFuBatch - 1) / kFuBatch;
    const bool lines_in_smem = (pc.nx <= kPairMaxLines) && (pc.ny <= kPairMaxLines);
    constexpr int n_rays = kFuRays;

    // ------------------------------------------------------------------------------------------------ prologue
    ProducerEnv cur, nxt;
    if (warp == 0) {
        producer_load(cur, lane, n_iter, pos_w, quat_w);
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
            for (int s = 0; s < kFuFullBars; ++s) bar_init(&sm.full_bar[s], 2);  // TMA bytes + header published
            for (int s = 0; s < kFuStages; ++s) bar_init(&sm.empty_bar[s], (uint32_t)kFuChunks);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        producer_window(cur, pc, pattern_radius);
        if (lane < n_iter && lane < kFuStages) {
            bar_arrive_expect_tx(&sm.full_bar[lane], kPairStageBytes);
            tma_load_window_planar(sm.stage[lane].p, &tmap, cur.ic0, cur.jr0, &sm.full_bar[lane]);
        }
        producer_frame(cur);
    } else if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                mb_init(&sm.obs_full[i], kFuBatch * kFuChunks);  // one arrival per finished chunk
                mb_init(&sm.obs_empty[i], 1);                    // the epilogue group, after layer 0 of the batch
            }
            for (int i = 0; i < kFuWStages; ++i) {
                mb_init(&sm.w_full[i], 1);
                mb_init(&sm.w_empty[i], 1);  // tcgen05.commit
            }
            mb_init(&sm.acc_full, 1);
            mb_init(&sm.act_ready, 1);
            mb_init(&sm.batch_done, 1);
            sm.next_chunk = 0;
            sm.vz0 = __ldg(ray_local + 2);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sptr(&sm.tmem_base)),
                     "r"((uint32_t)kFuTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else if (warp >= 4) {
        // pattern (slot order) + grid-line pairs, every global load issued before the first shared store; then the
        // observation operands are zeroed (columns >= 965 and unused env rows must hold finite values: weights are zero
        // there, but NaN * 0 = NaN)
        constexpr int kFill = 32 * (kFuWarps - 4);
        constexpr int kPatLoads = (2 * kPairMaxRays + kFill - 1) / kFill;
        constexpr int kLineLoads = (kPairMaxLines + kFill - 1) / kFill;
        const int ct = threadIdx.x - 128;
        float* pat_flat = sm.vx;  // vx, vy contiguous
        float pat[kPatLoads], xl[kLineLoads], xh[kLineLoads], yl[kLineLoads], yh[kLineLoads];
#pragma unroll
        for (int k = 0; k < kPatLoads; ++k) {
            const int i = ct + k * kFill;
            const int comp = i / kPairMaxRays, r = ray_of_slot(i - comp * kPairMaxRays);
            pat[k] = (i < 2 * kPairMaxRays && r < n_rays) ? __ldg(ray_local + 3 * r + comp) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kLineLoads; ++k) {
            const int i = ct + k * kFill;
            const bool okx = lines_in_smem && i < pc.nx, oky = lines_in_smem && i < pc.ny;
            xl[k] = okx ? __ldg(pc.xs + i) : 0.f;
            xh[k] = okx ? __ldg(pc.xs + i + 1) : 0.f;
            yl[k] = oky ? __ldg(pc.ys + i) : 0.f;
            yh[k] = oky ? __ldg(pc.ys + i + 1) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kPatLoads; ++k) {
            const int i = ct + k * kFill;
            if (i < 2 * kPairMaxRays) pat_flat[i] = pat[k];
        }
#pragma unroll
        for (int k = 0; k < kLineLoads; ++k) {
            const int i = ct + k * kFill;
            if (lines_in_smem && i < pc.nx) sm.xpair[i] = {xl[k], xh[k]};
            if (lines_in_smem && i < pc.ny) sm.ypair[i] = {yl[k], yh[k]};
        }
        uint4* z = reinterpret_cast<uint4*>(&sm.obs[0][0]);
        for (int i = ct; i < 2 * kFuObsBytes / 16; i += kFill) z[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = ct; i < kFuBiasFloats; i += kFill)
            sm.bias[i] = __ldg(reinterpret_cast<const float*>(packed + plan.bias_ofs) + i);
        fence_async_smem();  // the zeroed operands are read by the tensor core (async proxy)
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp < 4) {
        reg_dec<kFuRegsUtil>();
        if (warp == 0) {
            // =============================== scan producer (as variant 5; 4 stages) ===============================
            for (int base = 0; base < n_iter; base += 32) {
                if (base > 0) cur = nxt;
                producer_verdict(cur, sm, pc, pattern_radius, lines_in_smem);
                const int n_here = min(32, n_iter - base);
                for (int k = 0; k < n_here; ++k) {
                    if (k == kFuStages && base + 32 < n_iter) {  // next batch of 32 poses, computed while the ring is full
                        producer_load(nxt, base + 32 + lane, n_iter, pos_w, quat_w);
                        producer_window(nxt, pc, pattern_radius);
                        producer_frame(nxt);
                    }
                    if (lane == k) {
                        const int it = base + k;
                        const int s = it % kFuStages;
                        unsigned long long* full = &sm.full_bar[it % kFuFullBars];
                        PairStage& st = sm.stage[s];
                        if (it >= kFuStages) {  // (the first ring pass was loaded in the prologue)
                            bar_wait(&sm.empty_bar[s], ((uint32_t)(it / kFuStages) & 1u) ^ 1u);  // stage drained
                            if (cur.ok) {
                                bar_arrive_expect_tx(full, kPairStageBytes);
                                tma_load_window_planar(st.p, &tmap, cur.ic0, cur.jr0, full);
                            } else {
                                bar_arrive(full);
                            }
                        }
                        st.hdr = {cur.cw, cur.sz, cur.px, cur.py, cur.pz, cur.ic0, cur.jr0, cur.ncols, cur.nrows,
                                  cur.ok ? 1 : 0, 0, 0};
                        bar_arrive(full);  // header published (release)
                    }
                    __syncwarp();  // environments are issued strictly in order
                }
            }
        } else if (warp == 1 && lane == 0) {
            // =============================== weight streamer ===============================
            uint32_t cnt = 0;
            for (int b = 0; b < n_batches; ++b) {
                for (int pi = 0; pi < plan.n; ++pi, ++cnt) {
                    const uint32_t s = cnt % kFuWStages;
                    if (cnt >= (uint32_t)kFuWStages) mb_wait(&sm.w_empty[s], ((cnt / kFuWStages) - 1u) & 1u);
                    mb_expect_tx(&sm.w_full[s], plan.p[pi].bytes);
                    bulk_g2s(sm.w + s * kFuWStage, packed + plan.p[pi].gofs, plan.p[pi].bytes, &sm.w_full[s]);
                }
            }
        } else if (warp == 2 && lane == 0) {
            // =============================== MMA issuer ===============================
            const uint32_t idesc = make_idesc(kFuBatch);
            uint32_t cnt = 0, act_phase = 0;
            for (int b = 0; b < n_batches; ++b) {
                const int buf = b & 1;
                if (b == n_batches - 1) {  // a short last batch: stand in for the chunks of the environments that do not exist
                    const int missing = (n_batches * kFuBatch - n_iter) * kFuChunks;
                    for (int i = 0; i < missing; ++i) fu_arrive(&sm.obs_full[buf]);
                }
                if (b > 0) mb_wait(&sm.batch_done, (uint32_t)(b - 1) & 1u);  // the epilogues of batch b-1 have read TMEM
                mb_wait(&sm.obs_full[buf], (uint32_t)(b >> 1) & 1u);
                tc_fence_after();
                int layer = 0;
                for (int pi = 0; pi < plan.n; ++pi, ++cnt) {
                    const FuPiece P = plan.p[pi];
                    if (P.layer != layer) {  // the epilogue of the previous layer has written this layer's operand
                        layer = P.layer;
                        mb_wait(&sm.act_ready, act_phase);
                        act_phase ^= 1u;
                        tc_fence_after();
                    }
                    const uint32_t s = cnt % kFuWStages;
                    mb_wait(&sm.w_full[s], (cnt / kFuWStages) & 1u);
                    tc_fence_after();
                    const uint32_t a0 = sptr(sm.w + s * kFuWStage);
                    const uint32_t b0 = (layer == 0 ? sptr(sm.obs[buf]) : sptr(sm.act)) + (uint32_t)P.kstep0 * 2u * kOperandLbo;
                    for (int j = 0; j < P.nk; ++j)
                        umma(tmem + P.tcol, make_desc(a0 + j * 2 * P.lbo, P.lbo), make_desc(b0 + j * 2 * kOperandLbo, kOperandLbo),
                             idesc, !(P.first && j == 0));
                    umma_commit(&sm.w_empty[s]);  // the piece's ring stage is free once these MMAs retire
                    if (P.last) umma_commit(&sm.acc_full);
                }
            }
        }
    } else if (warp < 4 + kFuConsumerWarps) {
        reg_inc<kFuRegsConsumer>();
        // =============================== scan consumers ===============================
        while (true) {
            int chunk_id = 0;
            if (lane == 0) chunk_id = atomicAdd(&sm.next_chunk, 1);
            chunk_id = __shfl_sync(0xffffffffu, chunk_id, 0);
            const int it = chunk_id / kFuChunks, c = chunk_id - it * kFuChunks;
            if (it >= n_iter) break;
            const int env = (int)blockIdx.x + it * (int)gridDim.x;
            const int s = it % kFuStages;
            const PairStage& st = sm.stage[s];
            const int batch = it / kFuBatch, slot = it - batch * kFuBatch, buf = batch & 1;
            if (batch >= 2) bar_wait(&sm.obs_empty[buf], (uint32_t)((batch >> 1) - 1) & 1u);  // operand drained (batch - 2)
            bar_wait(&sm.full_bar[it % kFuFullBars], (uint32_t)(it / kFuFullBars) & 1u);
            const PairHeader h = st.hdr;
            float* __restrict__ obs_row = obs + (size_t)env * obs_stride;
            float* __restrict__ out_row = obs_row + kOperandHead;
            unsigned char* __restrict__ operand_row = sm.obs[buf] + (slot >> 3) * 128 + (slot & 7) * 16;
            const int r_begin = c * kPairChunk, r_end = min(r_begin + kPairChunk, n_rays);
            if (c == 0 && lane < kOperandHead)  // the head of the observation (written by the post-step kernel): k = lane
                *reinterpret_cast<__nv_bfloat16*>(operand_row + lane * 2) = __float2bfloat16_rn(obs_row[lane]);
            if (h.mode == 1) {
                PairCtx cx;
                const float sz2 = __fmul_rn(h.sz, 2.f);
                const float wx0 = sm.xpair[h.ic0].lo, wy0 = sm.ypair[h.jr0].lo;
                cx.CW = dup(h.cw), cx.SZ = dup(h.sz), cx.NSZ = dup(-h.sz), cx.S2 = dup(sz2), cx.NS2 = dup(-sz2);
                cx.PX = dup(h.px), cx.PY = dup(h.py), cx.PZ = dup(h.pz);
                cx.NWX0 = dup(-wx0), cx.NWY0 = dup(-wy0), cx.IDX = dup(pc.inv_dx), cx.IDY = dup(pc.inv_dy);
                cx.MAGIC = dup(kFloorMagic), cx.BASE = dup(base_offset);
                cx.NEG0 = dup(__uint_as_float(0x80000000u | (unsigned)(n_envs >> 31)));  // -0.0, opaque to the compiler
                cx.pz = h.pz, cx.max_d = max_d;
                cx.ZFLAT = dup(__fadd_rn(sm.vz0, h.pz));
                cx.cmax = (uint32_t)(h.ncols - 1), cx.rmax = (uint32_t)(h.nrows - 1);
                cx.xoff = (uint32_t)(reinterpret_cast<const unsigned char*>(sm.xpair + h.ic0) - smem_raw);
                cx.yoff = (uint32_t)(reinterpret_cast<const unsigned char*>(sm.ypair + h.jr0) - smem_raw);
                cx.eoff = (uint32_t)(reinterpret_cast<const unsigned char*>(st.p) - smem_raw);
                constexpr int kSink = kWriteObs ? kSinkOperandAndGlobal : kSinkOperand;
                for (int b0 = r_begin; b0 < r_end; b0 += kPairBatch) {
                    const int r = b0 + lane;  // rays r, r + 64 (slot 0) and r + 32, r + 96 (slot 1)
                    float* __restrict__ o = out_row + r;
                    unsigned defer;
                    if (b0 + kPairBatch <= n_rays) {
                        defer = resolve_pair<true, true, false, kSink>(smem_raw, sm, cx, b0 + 2 * lane, r, n_rays, o, nullptr,
                                                                       operand_row);
                        defer |= resolve_pair<true, true, false, kSink>(smem_raw, sm, cx, b0 + 64 + 2 * lane, r + 32, n_rays,
                                                                        o + 32, nullptr, operand_row) << 2;
                    } else {
                        defer = resolve_pair<false, true, false, kSink>(smem_raw, sm, cx, b0 + 2 * lane, r, n_rays, o, nullptr,
                                                                        operand_row);
                        defer |= resolve_pair<false, true, false, kSink>(smem_raw, sm, cx, b0 + 64 + 2 * lane, r + 32, n_rays,
                                                                         o + 32, nullptr, operand_row) << 2;
                    }
                    if (defer != 0u) {  // rare: cell guess off by one, ray on the closed far border / outside, general cell
                        for (int u = 0; u < 4; ++u) {
                            if (!((defer >> u) & 1u)) continue;
                            const int rr = r + (u >> 1) * 32 + (u & 1) * 64;
                            const int sl = b0 + (u >> 1) * 64 + 2 * lane + (u & 1);
                            const float vx = sm.vx[sl], vy = sm.vy[sl];
                            const float tx = -__fmul_rn(sz2, vy), ty = __fmul_rn(sz2, vx);
                            const float X = __fadd_rn(__fadd_rn(__fadd_rn(vx, __fmul_rn(h.cw, tx)), -__fmul_rn(h.sz, ty)), h.px);
                            const float Y = __fadd_rn(__fadd_rn(__fadd_rn(vy, __fmul_rn(h.cw, ty)), __fmul_rn(h.sz, tx)), h.py);
                            float hh;
                            pair_resolve_deferred_ray(&sm, &st, g, pc.inv_dx, pc.inv_dy, X, Y, __fadd_rn(sm.vz0, h.pz), h.pz,
                                                      max_d, base_offset, &hh, nullptr);
                            operand_store(operand_row, rr, hh);
                            if (kWriteObs) out_row[rr] = hh;
                        }
                    }
                }
            } else {
                // window not staged (too large / not covered / too many grid lines): resolve from global memory into the
                // fp32 row (always present), then mirror into the operand
                pair_resolve_chunk_from_global(&sm, lane, r_begin, r_end, h, g, pc, max_d, base_offset, out_row, nullptr);
                __syncwarp();
                for (int r = r_begin + lane; r < r_end; r += 32) operand_store(operand_row, r, out_row[r]);
            }
            fence_async_smem();  // operand stores (generic proxy) -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) {
                bar_arrive(&sm.empty_bar[s]);       // this chunk no longer reads the window stage
                fu_arrive(&sm.obs_full[buf]);       // ... and its part of the operand is written
            }
        }
    } else {
        reg_dec<kFuRegsEpilogue>();
        // =============================== epilogue group (TMEM lane = output feature) ===============================
        const int t = threadIdx.x - 32 * (4 + kFuConsumerWarps);
        const uint32_t t_lane = (uint32_t)((warp & 3) * 32) << 16;
        const float* b0p = sm.bias;
        const float* b1p = b0p + 80;
        const float* b2p = b1p + 64;
        const float* b3p = b2p + 256;
        const float* b4p = b3p + 160;
        const float* b5p = b4p + 128;
        uint32_t acc_phase = 0;
        auto wait_acc = [&]() {
            mb_wait(&sm.acc_full, acc_phase);
            acc_phase ^= 1u;
            tc_fence_after();
        };
        auto publish = [&](unsigned long long* also) {  // operand stores visible to the tensor core; TMEM reads done
            tc_fence_before();
            fence_async_smem();
            fu_epi_sync();
            if (t == 0) {
                if (also != nullptr) fu_arrive(also);
                fu_arrive(&sm.act_ready);
            }
        };
        for (int b = 0; b < n_batches; ++b) {
            const int buf = b & 1;
            // ---- layer 0: D0 -> A1 (80 features); the observation head (k = 0..3 of the operand) is kept for layer 2
            wait_acc();
            if (t >= 60 && t < 64) {  // (threads 60..63 write rows k = 60..63 of A2 below: they carry the head across)
#pragma unroll
                for (int e = 0; e < kFuBatch; ++e)
                    sm.head[e][t - 60] = *reinterpret_cast<const __nv_bfloat16*>(operand_at(sm.obs[buf], e, t - 60));
            }
            fu_epilogue_block(tmem + t_lane + 0, t < 80 ? b0p[t] : 0.f, sm.act, t, t < 80, nullptr);
            publish(&sm.obs_empty[buf]);  // the observation operand of this batch may be refilled (batch b + 2)
            // ---- layer 1: -> A2 = [e(60), obs[:, 0:4]]
            wait_acc();
            fu_epilogue_block(tmem + t_lane + 16, t < 60 ? b1p[t] : 0.f, sm.act, t, t < 64,
                              (t >= 60 && t < 64) ? &sm.head[0][t - 60] : nullptr);
            publish(nullptr);
            // ---- layer 2: two blocks of 128 features -> A3 (256)
            wait_acc();
            fu_epilogue_block(tmem + t_lane + 32, b2p[t], sm.act, t, true, nullptr);
            fu_epilogue_block(tmem + t_lane + 48, b2p[128 + t], sm.act, 128 + t, true, nullptr);
            publish(nullptr);
            // ---- layer 3: 128 + 32 features -> A4 (160)
            wait_acc();
            fu_epilogue_block(tmem + t_lane + 64, b3p[t], sm.act, t, true, nullptr);
            fu_epilogue_block(tmem + t_lane + 80, t < 32 ? b3p[128 + t] : 0.f, sm.act, 128 + t, t < 32, nullptr);
            publish(nullptr);
            // ---- layer 4: -> A5 (128)
            wait_acc();
            fu_epilogue_block(tmem + t_lane + 96, b4p[t], sm.act, t, true, nullptr);
            publish(nullptr);
            // ---- layer 5: mean = tanh(D5 + b5) (policy, 2 rows) or the linear value (1 row)
            wait_acc();
            {
                uint32_t r[16];
                fu_tmem_ld16(tmem + t_lane + 112, r);
                const int n_out = value_head ? 1 : 2;
                if (t < n_out) {
                    const float bb = b5p[t];
#pragma unroll
                    for (int e = 0; e < kFuBatch; ++e) {
                        const int it = b * kFuBatch + e;
                        if (it < n_iter) {
                            const size_t env = (size_t)blockIdx.x + (size_t)it * gridDim.x;
                            const float v = __uint_as_float(r[e]) + bb;
                            mean[env * n_out + t] = value_head ? v : tanhf(v);
                        }
                    }
                }
            }
            tc_fence_before();
            fu_epi_sync();
            if (t == 0) fu_arrive(&sm.batch_done);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)kFuTmemCols) : "memory");
    }
}

}  // namespace rover

extern "C" int64_t rover_policy_pack_fused(const RoverPolicyWeights* weights, void* packed, void* stream) {
    using namespace rover;
    static const FuPlan plan = make_fused_plan();
    if (packed == nullptr) return plan.total_bytes;
    if (weights == nullptr) {
        fail("rover_policy_pack_fused: weights is NULL");
        return -1;
    }
    FuPackArgs a;
    for (int l = 0; l < 6; ++l) {
        const bool out_ok = weights->out_dim[l] == layer_n_real(l) || (l == 5 && weights->out_dim[l] == 1);  // value head
        if (!weights->w[l] || !weights->b[l] || weights->in_dim[l] < layer_k_real(l) || !out_ok) {
            fail("rover_policy_pack_fused: layer %d has shape [%d,%d], expected [%d,%d]", l, weights->out_dim[l],
                 weights->in_dim[l], layer_n_real(l), layer_k_real(l));
            return -1;
        }
        a.w[l] = weights->w[l];
        a.b[l] = weights->b[l];
        a.in_dim[l] = weights->in_dim[l];
        a.out_dim[l] = weights->out_dim[l];
    }
    fused_pack_kernel<<<dim3(8, plan.n), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, plan, static_cast<unsigned char*>(packed));
    return check_launch("fused_pack_kernel") ? -1 : plan.total_bytes;
}

extern "C" int rover_scan_policy_fused(const float* pos_w, const float* quat_w, int32_t n_envs,
                                       const float* ray_starts_local, int32_t n_rays, const float* pattern_box,
                                       const RoverScanGrid* grid, const RoverPlaneCells* cells, float max_distance,
                                       float base_offset, float* obs, int32_t obs_stride, int32_t write_obs,
                                       const void* packed_fused, float* out, int32_t value_head, void* stream) {
    using namespace rover;
    ROVER_CHECK(n_envs >= 0, "rover_scan_policy_fused: negative n_envs");
    if (n_envs == 0) return 0;
    ROVER_CHECK(pos_w && quat_w && ray_starts_local && pattern_box && grid && cells && obs && packed_fused && out,
                "rover_scan_policy_fused: NULL argument");
    ROVER_CHECK(n_rays == kFuRays, "rover_scan_policy_fused: the policy reads a 961-ray scan (31 x 31 grid), got %d rays", n_rays);
    ROVER_CHECK(obs_stride >= kOperandHead + kFuRays, "rover_scan_policy_fused: obs_stride %d < 965", obs_stride);
    ROVER_CHECK(cells->xs && cells->ys && cells->entries && cells->entries_planar && cells->nx > 0 && cells->ny > 0,
                "rover_scan_policy_fused: needs the plane-cell table with its planar copy");
    ROVER_CHECK((reinterpret_cast<uintptr_t>(packed_fused) & 15) == 0, "rover_scan_policy_fused: packed image not 16B aligned");
    static const FuPlan plan = make_fused_plan();
    static int n_sms = 0;
    if (n_sms == 0) {
        int dev = 0;
        ROVER_CUDA(cudaGetDevice(&dev));
        ROVER_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
        ROVER_CUDA(cudaFuncSetAttribute(fused_scan_policy_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(FuSmem)));
        ROVER_CUDA(cudaFuncSetAttribute(fused_scan_policy_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(FuSmem)));
    }
    ScanGridDev g;
    g.n_levels = grid->n_levels;
    g.span = grid->span;
    ROVER_CHECK(grid->n_levels >= 0 && grid->n_levels <= ROVER_MAX_LEVELS, "rover_scan_policy_fused: bad grid");
    for (int l = 0; l < grid->n_levels; ++l) {
        const RoverScanLevel& L = grid->level[l];
        g.level[l] = ScanLevelDev{L.ox, L.oy, L.cell, L.inv_cell, L.ncx, L.ncy, L.start_offset, 0};
    }
    g.cell_start = grid->cell_start;
    g.rec = reinterpret_cast<const float4*>(grid->records);
    alignas(64) CUtensorMap tmap;
    if (const int rc = encode_planar_tensor_map(&tmap, cells, kPairPitch, kPairWin, "rover_scan_policy_fused")) return rc;
    PlaneCellsDev pc{cells->xs, cells->ys, reinterpret_cast<const float4*>(cells->entries), cells->nx, cells->ny,
                     cells->inv_dx, cells->inv_dy};
    const float rx = fmaxf(fabsf(pattern_box[0]), fabsf(pattern_box[1])), ry = fmaxf(fabsf(pattern_box[2]), fabsf(pattern_box[3]));
    const float radius = sqrtf(rx * rx + ry * ry) * 1.0001f + 1.0e-3f;
    const int grid_dim = n_envs < n_sms ? n_envs : n_sms;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (write_obs)
        fused_scan_policy_kernel<true><<<grid_dim, kFuThreads, sizeof(FuSmem), s>>>(
            pos_w, quat_w, n_envs, ray_starts_local, g, pc, tmap, radius, max_distance, base_offset, obs, obs_stride,
            static_cast<const unsigned char*>(packed_fused), plan, out, value_head);
    else
        fused_scan_policy_kernel<false><<<grid_dim, kFuThreads, sizeof(FuSmem), s>>>(
            pos_w, quat_w, n_envs, ray_starts_local, g, pc, tmap, radius, max_distance, base_offset, obs, obs_stride,
            static_cast<const unsigned char*>(packed_fused), plan, out, value_head);
    return check_launch("fused_scan_policy_kernel");
}
