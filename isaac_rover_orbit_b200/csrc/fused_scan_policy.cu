// Height scan fused with the heightmap encoder of the policy: the observation never leaves the SM.
//
// Replaces, in one launch, the height scan and the part of the network that reads it (reference root relative):
//   height_scan_rover                rover_envs/envs/navigation/mdp/observations.py:35-45  (+ ORBIT RayCaster / warp raycast)
//   HeightmapEncoder (961 -> 80 -> 60, LeakyReLU)   rover_envs/envs/navigation/learning/skrl/models.py:24-36, called at :95-97
// (BASELINE.json configs[3]: "policy forward ... fused with the observation kernel").  Unfused, the 961 heights of an
// environment make a round trip through L2 / HBM -- 3844 B written by the scan (+ 1930 B for the bf16 mirror), 3860 /
// 1930 B read back by the policy.  Here the scan's consumer warps write every height as bf16 straight into the
// shared-memory operand of the layer-0 MMA; what leaves the SM is the encoder output, [e(60), obs[:, 0:4]] as 64 bf16 =
// 128 B per environment -- the input of the MLP (64 -> 256 -> 160 -> 128 -> 2), which runs as a second, small launch
// (policy_mlp.cu) on 128-environment tiles.
//
// Orientation.  The scan is environment-major (one environment = 961 rays at a time), so a 128-environment A tile as in
// policy_ws.cu (247 KB of bf16) cannot be collected on chip.  The two encoder layers therefore run TRANSPOSED on tcgen05:
//     D^T [features (M = 128 lanes) x envs (N = 16)] = W [features x K] * A^T [K x envs]
// with a batch of 16 environments as the B operand (K-major rows of 16 B core matrices in shared memory: the operand
// the scan fills) and the WEIGHTS as the A operand -- and W0, the only large matrix (80 x 961), is WEIGHT-STATIONARY IN
// TENSOR MEMORY: 128 lanes x 488 columns (2 bf16 per 32-bit cell) of the SM's 512 TMEM columns hold it for the whole
// launch (tcgen05.mma reads its A operand from TMEM), leaving 16 columns for the accumulator.  Nothing is streamed per
// batch: a first version that streamed all six layers' weights from L2 per 16-environment batch (325 KB) ran 2.3x
// SLOWER than the unfused pair -- 20 KB of weights per environment on top of the 22 KB table window, through a
// latency-bound 30 KB ring (profiles/r02_fused_scan_policy.md).  W1 (60 x 80) sits in shared memory.
//
// Roles (one persistent CTA per SM, 20 warps = 5 warpgroups).  The CTA is launched with 96 registers per thread (640 x 96 =
// 61,440: setmaxnreg moves registers inside THAT pool, not inside the SM's file) and re-balances them per warpgroup: 120
// for the three consumer groups, 64 for the producer / issuer group, 56 for the epilogue group (61,440 in all):
//   warp 0       scan producer: poses -> window -> 3-D tensor-map TMA load per environment (4-stage ring), as variant 5
//   warp 1       one bulk copy of W1 at start; TMEM allocation
//   warp 2       MMA issuer: layer 0 (61 MMAs, A from TMEM) when a batch of 16 environments is complete, layer 1 after
//                the epilogue has written its operand
//   warps 4-15   scan consumers (variant 5's packed-fp32 ray pairs; 256-ray chunks handed out dynamically, in order)
//   warps 16-19  epilogue group: loads W0 into TMEM while the first batch is being scanned; then TMEM lane = output
//                feature: + bias, LeakyReLU, bf16 -> layer 1's operand / the encoder output.
// The encoder work of batch b runs while the consumers scan batch b+1 (two observation operands); only the last batch's
// two layers (~2 us) are exposed at the end of the launch.
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
constexpr int kFuActBytes = (80 / 8) * kOperandLbo;      // layer 1's operand, K = 80
constexpr int kFuW0Cols = kFuK0 / 2;           // 488 TMEM columns: W0 as the A operand, 2 bf16 per 32-bit cell
constexpr int kFuAccCol = kFuW0Cols;           // accumulator (16 columns), used by layer 0 then layer 1 of a batch
constexpr int kFuTmemCols = 512;
constexpr int kFuEncCols = 64;                 // encoder output row: e(60), obs[:, 0:4]
constexpr int kFuBiasFloats = 80 + 64;
// packed image: [W0^T: kFuW0Cols x 128 lanes u32 (word j of lane t at j * 128 + t)] [W1 as an A operand: 10 planes x 64
// rows x 16 B] [bias0 (80), bias1 (64) fp32]
constexpr int kFuW0Bytes = kFuW0Cols * 128 * 4;
constexpr int kFuW1Bytes = 10 * 64 * 16, kFuW1Lbo = 64 * 16, kFuW1Slack = 1024;  // (M = 128 reads 64 rows past W1's)
constexpr int kFuW1Ofs = kFuW0Bytes, kFuBiasOfs = kFuW0Bytes + kFuW1Bytes;
constexpr int kFuPackedBytes = kFuBiasOfs + kFuBiasFloats * 4;
constexpr int kFuRegsLaunch = 96, kFuRegsConsumer = 120, kFuRegsUtil = 64, kFuRegsEpilogue = 56;
static_assert(128 * (3 * kFuRegsConsumer + kFuRegsUtil + kFuRegsEpilogue) <= kFuThreads * kFuRegsLaunch,
              "setmaxnreg: the re-balanced budgets must fit the pool the CTA was launched with");
// Phase aliasing (scan ring): environment E signals `full` barrier E % 8 with parity (E / 8) & 1; a waiter is fooled only if
// E - 8 has not completed yet.  Chunks are handed out in order and at most 12 are outstanding, so when a warp holds a
// chunk of E, at least 20 of the 32 chunks of E-8 .. E-1 are finished, hence one of E-4 .. E-1 -- whose load was issued
// only after E-8 .. E-5 had been consumed (4 stages, in-order producer).
static_assert(kFuFullBars == 2 * kFuStages && 8 * kFuChunks - kFuConsumerWarps > 4 * kFuChunks, "phase aliasing argument");

struct FuPackArgs {
    const float* w0;
    const float* b0;
    const float* w1;
    const float* b1;
    int in0, in1;  // row strides of the nn.Linear weights
};

__global__ void fused_pack_kernel(const __grid_constant__ FuPackArgs a, unsigned char* __restrict__ packed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    uint32_t* w0t = reinterpret_cast<uint32_t*>(packed);
    for (int idx = tid; idx < kFuW0Cols * 128; idx += nthr) {
        const int j = idx >> 7, t = idx & 127;  // TMEM column j (K = 2j, 2j + 1 = observation columns), lane t = feature
        float v[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int src_k = 2 * j + h - kEncInOffset;  // layer-0 K index = observation column
            if (t < 80 && src_k >= 0 && src_k < kEncIn) v[h] = a.w0[(size_t)t * a.in0 + src_k];
        }
        w0t[idx] = pack_bf16(v[0], v[1]);
    }
    __nv_bfloat16* w1 = reinterpret_cast<__nv_bfloat16*>(packed + kFuW1Ofs);
    for (int idx = tid; idx < kFuW1Bytes / 2; idx += nthr) {
        const int jj = idx & 7, r = (idx >> 3) & 63, plane = idx >> 9;
        const int k = plane * 8 + jj;
        w1[idx] = __float2bfloat16_rn((r < 60 && k < 80) ? a.w1[(size_t)r * a.in1 + k] : 0.f);
    }
    float* bias = reinterpret_cast<float*>(packed + kFuBiasOfs);
    for (int n = tid; n < kFuBiasFloats; n += nthr) bias[n] = n < 80 ? a.b0[n] : (n - 80 < 60 ? a.b1[n - 80] : 0.f);
}

struct __align__(128) FuSmem {
    PairStage stage[kFuStages];
    float vx[kPairMaxRays], vy[kPairMaxRays];
    LinePair2 xpair[kPairMaxLines], ypair[kPairMaxLines];
    unsigned char obs[2][kFuObsBytes];   // observation operands of two batches: [k / 8][16 env rows x 16 B (+ 16 B skew)]
    unsigned char act[kFuActBytes];      // layer 1's operand
    unsigned char w1[kFuW1Bytes + kFuW1Slack];
    float bias[kFuBiasFloats];
    unsigned long long full_bar[kFuFullBars], empty_bar[kFuStages];
    unsigned long long obs_full[2], obs_empty[2];
    unsigned long long w1_full, w0_ready, acc_full, act_ready, batch_done;
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
// mbarrier wait of the roles that are NOT on the scan's critical path (issuer, epilogue group): poll, then sleep.  A
// batch takes ~8 us to fill, their work ~2 us; a tight try_wait loop in 5 extra warps took issue slots from the scan
// consumers sharing their schedulers (ncu, round 2: 46.0 M warp-instructions at 16384 envs against 32.3 M for the scan).
__device__ __forceinline__ void fu_wait_relaxed(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(sptr(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(256);
    }
}
__device__ __forceinline__ void fu_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void fu_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]: the A operand read from tensor memory (lane = row, 2 bf16 per 32-bit column)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
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

template <bool kWriteObs>
__global__ void __launch_bounds__(kFuThreads, 1)
fused_scan_encoder_kernel(const float* __restrict__ pos_w, const float* __restrict__ quat_w, int n_envs,
                          const float* __restrict__ ray_local, const __grid_constant__ ScanGridDev g,
                          const __grid_constant__ PlaneCellsDev pc, const __grid_constant__ CUtensorMap tmap,
                          float pattern_radius, float max_d, float base_offset, float* __restrict__ obs, int obs_stride,
                          const unsigned char* __restrict__ packed, __nv_bfloat16* __restrict__ enc) {
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
            mb_init(&sm.w1_full, 1);
            mb_init(&sm.w0_ready, 1);
            mb_init(&sm.acc_full, 1);
            mb_init(&sm.act_ready, 1);
            mb_init(&sm.batch_done, 1);
            sm.next_chunk = 0;
            sm.vz0 = __ldg(ray_local + 2);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mb_expect_tx(&sm.w1_full, kFuW1Bytes);
            bulk_g2s(sm.w1, packed + kFuW1Ofs, kFuW1Bytes, &sm.w1_full);
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
            sm.bias[i] = __ldg(reinterpret_cast<const float*>(packed + kFuBiasOfs) + i);
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
        } else if (warp == 2 && lane == 0) {
            // =============================== MMA issuer ===============================
            const uint32_t idesc = make_idesc(kFuBatch);
            const uint32_t acc = tmem + kFuAccCol;
            fu_wait_relaxed(&sm.w1_full, 0u);
            fu_wait_relaxed(&sm.w0_ready, 0u);  // W0 sits in TMEM columns [0, 488)
            tc_fence_after();
            for (int b = 0; b < n_batches; ++b) {
                const int buf = b & 1;
                if (b == n_batches - 1) {  // a short last batch: stand in for the chunks of the environments that do not exist
                    const int missing = (n_batches * kFuBatch - n_iter) * kFuChunks;
                    for (int i = 0; i < missing; ++i) fu_arrive(&sm.obs_full[buf]);
                }
                if (b > 0) fu_wait_relaxed(&sm.batch_done, (uint32_t)(b - 1) & 1u);  // the epilogues of batch b-1 have read TMEM
                fu_wait_relaxed(&sm.obs_full[buf], (uint32_t)(b >> 1) & 1u);
                tc_fence_after();
                // ---- layer 0: D = W0 (TMEM) x obs^T (shared), K = 976 in 61 steps of 16 (8 TMEM columns each)
                const uint32_t b0 = sptr(sm.obs[buf]);
                for (int ks = 0; ks < kFuK0 / 16; ++ks)
                    umma_ts(acc, tmem + ks * 8, make_desc(b0 + ks * 2 * kOperandLbo, kOperandLbo), idesc, ks != 0);
                umma_commit(&sm.acc_full);
                // ---- layer 1: D = W1 (shared) x A1^T (shared), K = 80; the accumulator columns are re-used
                fu_wait_relaxed(&sm.act_ready, (uint32_t)b & 1u);
                tc_fence_after();
                const uint32_t a1 = sptr(sm.w1), b1 = sptr(sm.act);
                for (int j = 0; j < 80 / 16; ++j)
                    umma(acc, make_desc(a1 + j * 2 * kFuW1Lbo, kFuW1Lbo), make_desc(b1 + j * 2 * kOperandLbo, kOperandLbo), idesc,
                         j != 0);
                umma_commit(&sm.acc_full);
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
                const float sz2 = __fmul_rn(h.sz, 2.f);
                const PairCtx cx = make_pair_ctx(h, sm, st, smem_raw, pc.inv_dx, pc.inv_dy, base_offset, max_d, sm.vz0,
                                                 n_envs >> 31);
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
        // ---- W0 -> tensor memory, while the consumers scan the first batch: lane t = feature t, column j = K pair j.
        //      The image is column-major (word j of lane t at j * 128 + t): a warp's load is one 128-byte line.
        {
            const uint32_t* __restrict__ w0t = reinterpret_cast<const uint32_t*>(packed) + t;
            static_assert(kFuW0Cols % 8 == 0, "W0 is stored 8 columns at a time");
#pragma unroll 4
            for (int c0 = 0; c0 < kFuW0Cols; c0 += 8) {
                uint32_t r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = __ldg(w0t + (size_t)(c0 + i) * 128);
                fu_tmem_st8(tmem + t_lane + c0, r);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            fu_epi_sync();
            if (t == 0) fu_arrive(&sm.w0_ready);
        }
        const float bias0 = t < 80 ? sm.bias[t] : 0.f;
        const float bias1 = t < 60 ? sm.bias[80 + t] : 0.f;
        uint32_t acc_phase = 0;
        const uint32_t acc = tmem + t_lane + kFuAccCol;
        for (int b = 0; b < n_batches; ++b) {
            const int buf = b & 1;
            // ---- layer 0: D0 -> A1 (80 features); the observation head (k = 0..3 of the operand) is kept for the output
            fu_wait_relaxed(&sm.acc_full, acc_phase);
            acc_phase ^= 1u;
            tc_fence_after();
            if (t >= 60 && t < 64) {  // (threads 60..63 write columns 60..63 of the encoder output below)
#pragma unroll
                for (int e = 0; e < kFuBatch; ++e)
                    sm.head[e][t - 60] = *reinterpret_cast<const __nv_bfloat16*>(operand_at(sm.obs[buf], e, t - 60));
            }
            {
                uint32_t r[16];
                fu_tmem_ld16(acc, r);
                if (t < 80) {
#pragma unroll
                    for (int e = 0; e < kFuBatch; ++e)
                        *reinterpret_cast<__nv_bfloat16*>(operand_at(sm.act, e, t)) =
                            __float2bfloat16_rn(leaky(__uint_as_float(r[e]) + bias0));
                }
            }
            tc_fence_before();   // TMEM reads done before layer 1 overwrites the accumulator
            fence_async_smem();  // operand stores (generic proxy) -> visible to the tensor core (async proxy)
            fu_epi_sync();
            if (t == 0) {
                fu_arrive(&sm.obs_empty[buf]);  // the observation operand of this batch may be refilled (batch b + 2)
                fu_arrive(&sm.act_ready);
            }
            // ---- layer 1: -> encoder output row [e(60), obs[:, 0:4]] (bf16), the A operand of the MLP's first layer
            fu_wait_relaxed(&sm.acc_full, acc_phase);
            acc_phase ^= 1u;
            tc_fence_after();
            {
                uint32_t r[16];
                fu_tmem_ld16(acc, r);
                if (t < kFuEncCols) {
#pragma unroll
                    for (int e = 0; e < kFuBatch; ++e) {
                        const int it = b * kFuBatch + e;
                        if (it < n_iter) {
                            const size_t env = (size_t)blockIdx.x + (size_t)it * gridDim.x;
                            const __nv_bfloat16 v = t < 60 ? __float2bfloat16_rn(leaky(__uint_as_float(r[e]) + bias1))
                                                           : sm.head[e][t - 60];
                            enc[env * kFuEncCols + t] = v;
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
    if (packed == nullptr) return kFuPackedBytes;
    if (weights == nullptr) {
        fail("rover_policy_pack_fused: weights is NULL");
        return -1;
    }
    for (int l = 0; l < 2; ++l) {
        if (!weights->w[l] || !weights->b[l] || weights->in_dim[l] < layer_k_real(l) || weights->out_dim[l] != layer_n_real(l)) {
            fail("rover_policy_pack_fused: layer %d has shape [%d,%d], expected [%d,%d]", l, weights->out_dim[l],
                 weights->in_dim[l], layer_n_real(l), layer_k_real(l));
            return -1;
        }
    }
    FuPackArgs a{weights->w[0], weights->b[0], weights->w[1], weights->b[1], weights->in_dim[0], weights->in_dim[1]};
    fused_pack_kernel<<<64, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, static_cast<unsigned char*>(packed));
    return check_launch("fused_pack_kernel") ? -1 : kFuPackedBytes;
}

extern "C" int rover_scan_encoder_fused(const float* pos_w, const float* quat_w, int32_t n_envs,
                                        const float* ray_starts_local, int32_t n_rays, const float* pattern_box,
                                        const RoverScanGrid* grid, const RoverPlaneCells* cells, float max_distance,
                                        float base_offset, float* obs, int32_t obs_stride, int32_t write_obs,
                                        const void* packed_fused, uint16_t* enc_bf16, void* stream) {
    using namespace rover;
    ROVER_CHECK(n_envs >= 0, "rover_scan_encoder_fused: negative n_envs");
    if (n_envs == 0) return 0;
    ROVER_CHECK(pos_w && quat_w && ray_starts_local && pattern_box && grid && cells && obs && packed_fused && enc_bf16,
                "rover_scan_encoder_fused: NULL argument");
    ROVER_CHECK(n_rays == kFuRays, "rover_scan_encoder_fused: the encoder reads a 961-ray scan (31 x 31 grid), got %d rays", n_rays);
    ROVER_CHECK(obs_stride >= kOperandHead + kFuRays, "rover_scan_encoder_fused: obs_stride %d < 965", obs_stride);
    ROVER_CHECK(cells->xs && cells->ys && cells->entries && cells->entries_planar && cells->nx > 0 && cells->ny > 0,
                "rover_scan_encoder_fused: needs the plane-cell table with its planar copy");
    ROVER_CHECK((reinterpret_cast<uintptr_t>(packed_fused) & 15) == 0, "rover_scan_encoder_fused: packed image not 16B aligned");
    static int n_sms = 0;
    if (n_sms == 0) {
        int dev = 0;
        ROVER_CUDA(cudaGetDevice(&dev));
        ROVER_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
        ROVER_CUDA(cudaFuncSetAttribute(fused_scan_encoder_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(FuSmem)));
        ROVER_CUDA(cudaFuncSetAttribute(fused_scan_encoder_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(FuSmem)));
    }
    ScanGridDev g;
    g.n_levels = grid->n_levels;
    g.span = grid->span;
    ROVER_CHECK(grid->n_levels >= 0 && grid->n_levels <= ROVER_MAX_LEVELS, "rover_scan_encoder_fused: bad grid");
    for (int l = 0; l < grid->n_levels; ++l) {
        const RoverScanLevel& L = grid->level[l];
        g.level[l] = ScanLevelDev{L.ox, L.oy, L.cell, L.inv_cell, L.ncx, L.ncy, L.start_offset, 0};
    }
    g.cell_start = grid->cell_start;
    g.rec = reinterpret_cast<const float4*>(grid->records);
    alignas(64) CUtensorMap tmap;
    if (const int rc = encode_planar_tensor_map(&tmap, cells, kPairPitch, kPairWin, "rover_scan_encoder_fused")) return rc;
    PlaneCellsDev pc{cells->xs, cells->ys, reinterpret_cast<const float4*>(cells->entries), cells->nx, cells->ny,
                     cells->inv_dx, cells->inv_dy};
    const float rx = fmaxf(fabsf(pattern_box[0]), fabsf(pattern_box[1])), ry = fmaxf(fabsf(pattern_box[2]), fabsf(pattern_box[3]));
    const float radius = sqrtf(rx * rx + ry * ry) * 1.0001f + 1.0e-3f;
    const int grid_dim = n_envs < n_sms ? n_envs : n_sms;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    __nv_bfloat16* enc = reinterpret_cast<__nv_bfloat16*>(enc_bf16);
    if (write_obs)
        fused_scan_encoder_kernel<true><<<grid_dim, kFuThreads, sizeof(FuSmem), s>>>(
            pos_w, quat_w, n_envs, ray_starts_local, g, pc, tmap, radius, max_distance, base_offset, obs, obs_stride,
            static_cast<const unsigned char*>(packed_fused), enc);
    else
        fused_scan_encoder_kernel<false><<<grid_dim, kFuThreads, sizeof(FuSmem), s>>>(
            pos_w, quat_w, n_envs, ray_starts_local, g, pc, tmap, radius, max_distance, base_offset, obs, obs_stride,
            static_cast<const unsigned char*>(packed_fused), enc);
    return check_launch("fused_scan_encoder_kernel");
}
