// Persistent, warp-specialised plane-cell height scan -- variant 4.
//
// Why (ncu on variant 3, profiles/r01_scan_*): one CTA per environment pays a serial prologue (pose load -> atan2 /
// sincos -> window -> bulk copies -> wait, ~3 us) that even 8 resident CTAs hide only partly (17 % barrier stalls),
// and the per-ray loads of the ray-pattern table miss a 23 KB L1 (36 % long-scoreboard stalls).
//
// Structure (the TMA producer / consumer ring of the Blackwell playbook, with SIMT math instead of MMA):
//   * grid = #SMs persistent CTAs (one per SM, 480 threads); CTA b owns environments b, b + grid, b + 2 grid, ...
//   * warp 0 = producer.  Every 32 environments its lanes compute 32 sensor frames at once (ORBIT yaw_quat in the
//     reference's fp32 order).  Per environment it derives the table window from the position alone
//     (pattern radius, no yaw needed), loads the window's grid-line pairs, publishes frame + window in the stage
//     header, arms the stage's `full` mbarrier with the byte count and issues ONE 2-D tensor-map TMA load (UTMALDG)
//     of the kPipeWin x kPipeWin-cell window.
//   * warps 1..14 = two consumer groups of 7 warps; group q takes the CTA's environments q, q+2, ...  A group waits
//     on `full`, resolves 961 rays (5 per thread, interleaved for ILP) from shared memory only (pattern table, line
//     pairs, table entries), stores the heights, and releases the stage through the `empty` mbarrier.
//     (15 warps -> 16-warp register allocation -> 128 registers/thread: no spills; 25 warps capped at 72 and spilled.)
//   * 8 stages of 22 KB: up to eight table windows are in flight per SM, so the HBM/L2 latency of a window
//     (~2 us cold) is hidden behind the ray phases of the environments ahead of it; pattern table and grid lines
//     are loaded once per SM.
// Rays in general cells, windows that do not fit or do not cover (non-uniform lattices) use the global-memory
// helpers, so the staging can never change a result.
#include "scan_pipe.cuh"

namespace rover {

constexpr int kPipeGroups = 2;          // consumer groups; group q resolves environments it = q, q+2, ... of the CTA
constexpr int kPipeConsumerWarps = 7;   // warps per consumer group (1 + 2*7 = 15 warps -> 16-warp allocation, 128 regs)
constexpr int kPipeRaysPerThread = 5;   // rays resolved together by one consumer thread (ILP)
constexpr int kPipeConsumers = 32 * kPipeConsumerWarps;
constexpr int kPipeThreads = 32 * (1 + kPipeGroups * kPipeConsumerWarps);
constexpr int kPipeStages = 8;          // ring depth (multiple of kPipeGroups: a stage always serves the same group)
constexpr int kPipeWin = 26;            // window cells per axis
constexpr int kPipeCtasPerSm = 1;
constexpr int kPipeMaxRays = 1024;      // pattern table held in shared memory (float4 per ray)
constexpr int kPipeMaxLines = 1024;     // grid lines per axis held in shared memory (else read from global)

struct __align__(128) PipeStage {
    float4 ent[kPipeWin * kPipeWin * 2];
    LinePair2 xp[kPipeWin];
    LinePair2 yp[kPipeWin];
    StageHeader hdr;
};

struct PipeSmem {
    PipeStage stage[kPipeStages];
    float4 pattern[kPipeMaxRays];
    float xs[kPipeMaxLines + 1];
    float ys[kPipeMaxLines + 1];
    unsigned long long full_bar[kPipeStages];
    unsigned long long empty_bar[kPipeStages];
};

static_assert(kPipeStages % kPipeGroups == 0, "a stage must always be consumed by the same group");
static_assert(sizeof(PipeSmem) <= 227 * 1024, "PipeSmem exceeds the shared memory of one SM");

// Rare paths, kept out of line so that they do not raise the register pressure of the consumer loop.
// (a) a ray whose cell guess missed, that lies on the closed far border / outside the grid, or sits in a general cell
__device__ __noinline__ void resolve_deferred_ray(const PipeStage* st, const ScanGridDev& g, float inv_dx, float inv_dy,
                                                  float X, float Y, float Z, float pz, float max_d, float base_offset,
                                                  float* __restrict__ out, float* __restrict__ hit3) {
    const int cmax = st->hdr.ncols - 1, rmax = st->hdr.nrows - 1;
    const float wx0 = st->xp[0].lo, wy0 = st->yp[0].lo, wx1 = st->xp[cmax].hi, wy1 = st->yp[rmax].hi;
    float zhit = -INFINITY;
    if (X >= wx0 && X <= wx1 && Y >= wy0 && Y <= wy1) {
        int ci = min(max(__float2int_rd((X - wx0) * inv_dx), 0), cmax);
        int cj = min(max(__float2int_rd((Y - wy0) * inv_dy), 0), rmax);
        while (ci > 0 && X < st->xp[ci].lo) --ci;
        while (ci < cmax && X >= st->xp[ci].hi) ++ci;
        while (cj > 0 && Y < st->yp[cj].lo) --cj;
        while (cj < rmax && Y >= st->yp[cj].hi) ++cj;
        const int e = 2 * (cj * kPipeWin + ci);
        const float4 q = st->ent[e + 1];
        zhit = (q.w == 0.f) ? eval_cell(st->ent[e], q, __fsub_rn(X, st->xp[ci].lo), __fsub_rn(Y, st->yp[cj].lo), Z, max_d)
                            : walk_home_grid(g, X, Y, Z, max_d);
    }
    store_result(pz, X, Y, Z, zhit, base_offset, out, hit3);
}

// (b) a whole environment whose window is not staged (too large, or not covered on a non-uniform lattice)
__device__ __noinline__ void resolve_env_from_global(const float4* pattern, int t, int n_rays, const StageHeader h,
                                                     const ScanGridDev& g, const PlaneCellsDev& pc, float max_d,
                                                     float base_offset, float* __restrict__ out_row,
                                                     float* __restrict__ hits_row) {
    const float sz2 = __fmul_rn(h.sz, 2.f);
    for (int r = t; r < n_rays; r += kPipeConsumers) {
        const float4 v = pattern[r];
        const float tx = -__fmul_rn(sz2, v.y), ty = __fmul_rn(sz2, v.x);
        const float X = __fadd_rn(__fadd_rn(__fadd_rn(v.x, __fmul_rn(h.cw, tx)), -__fmul_rn(h.sz, ty)), h.px);
        const float Y = __fadd_rn(__fadd_rn(__fadd_rn(v.y, __fmul_rn(h.cw, ty)), __fmul_rn(h.sz, tx)), h.py);
        const float Z = __fadd_rn(v.z, h.pz);
        store_result(h.pz, X, Y, Z, resolve_from_global(g, pc, X, Y, Z, max_d), base_offset, out_row + r,
                     hits_row ? hits_row + 3 * (size_t)r : nullptr);
    }
}

template <bool kHits>
__global__ void __launch_bounds__(kPipeThreads, kPipeCtasPerSm)
height_scan_pipelined_kernel(const float* __restrict__ pos_w, const float* __restrict__ quat_w, int n_envs,
                             const float* __restrict__ ray_local, int n_rays, const __grid_constant__ ScanGridDev g,
                             const __grid_constant__ PlaneCellsDev pc, const __grid_constant__ CUtensorMap tmap,
                             float pattern_radius, float max_d,
                             float base_offset, float* __restrict__ out, int out_stride, float* __restrict__ hits) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PipeSmem& sm = *reinterpret_cast<PipeSmem*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_iter = (n_envs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // envs of this CTA

    // ---- one-time setup: barriers + pattern table
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPipeStages; ++s) {
            bar_init(&sm.full_bar[s], 1);
            bar_init(&sm.empty_bar[s], kPipeConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the producer warp computes its first 32 sensor frames while the other warps fill the shared tables
    float fcw = 0.f, fsz = 0.f, fpx = 0.f, fpy = 0.f, fpz = 0.f;
    if (warp == 0) {
        const int e = (int)blockIdx.x + lane * (int)gridDim.x;
        if (lane < n_iter && e < n_envs) {
            const SensorFrame f = make_frame(pos_w + 3 * (size_t)e, quat_w + 4 * (size_t)e);
            fcw = f.cw, fsz = f.sz, fpx = f.px, fpy = f.py, fpz = f.pz;
        }
    }
    for (int r = threadIdx.x; r < n_rays; r += kPipeThreads)
        sm.pattern[r] = make_float4(__ldg(ray_local + 3 * r), __ldg(ray_local + 3 * r + 1), __ldg(ray_local + 3 * r + 2), 0.f);
    const bool lines_in_smem = (pc.nx <= kPipeMaxLines) && (pc.ny <= kPipeMaxLines);
    if (lines_in_smem) {
        for (int i = threadIdx.x; i <= pc.nx; i += kPipeThreads) sm.xs[i] = __ldg(pc.xs + i);
        for (int i = threadIdx.x; i <= pc.ny; i += kPipeThreads) sm.ys[i] = __ldg(pc.ys + i);
    }
    __syncthreads();

    const float gx_lo = __ldg(pc.xs), gx_hi = __ldg(pc.xs + pc.nx), gy_lo = __ldg(pc.ys), gy_hi = __ldg(pc.ys + pc.ny);

    if (warp == 0) {
        // =============================== producer ===============================
        for (int it = 0; it < n_iter; ++it) {
            if ((it & 31) == 0 && it > 0) {  // 32 sensor frames at once, one per lane (first batch: see above)
                const int e = (int)blockIdx.x + (it + lane) * (int)gridDim.x;
                if (e < n_envs) {
                    const SensorFrame f = make_frame(pos_w + 3 * (size_t)e, quat_w + 4 * (size_t)e);
                    fcw = f.cw, fsz = f.sz, fpx = f.px, fpy = f.py, fpz = f.pz;
                }
            }
            const int src = it & 31;
            const float cw = __shfl_sync(0xffffffffu, fcw, src), sz = __shfl_sync(0xffffffffu, fsz, src);
            const float px = __shfl_sync(0xffffffffu, fpx, src), py = __shfl_sync(0xffffffffu, fpy, src);
            const float pz = __shfl_sync(0xffffffffu, fpz, src);
            const int s = it % kPipeStages;
            const uint32_t ph = (uint32_t)(it / kPipeStages) & 1u;
            PipeStage& st = sm.stage[s];
            bar_wait(&sm.empty_bar[s], ph ^ 1u);  // consumers have drained this stage

            // window from the position alone: every ray origin lies within pattern_radius of (px, py)
            const float xmin = px - pattern_radius, xmax = px + pattern_radius;
            const float ymin = py - pattern_radius, ymax = py + pattern_radius;
            const int ic0 = min(max(guess_col(xmin, gx_lo, pc.inv_dx) - 1, 0), pc.nx - 1);
            const int jr0 = min(max(guess_col(ymin, gy_lo, pc.inv_dy) - 1, 0), pc.ny - 1);
            const int ic1 = min(max(guess_col(xmax, gx_lo, pc.inv_dx) + 1, 0), pc.nx - 1);
            const int jr1 = min(max(guess_col(ymax, gy_lo, pc.inv_dy) + 1, 0), pc.ny - 1);
            const int ncols = ic1 - ic0 + 1, nrows = jr1 - jr0 + 1;
            bool ok = (ncols <= kPipeWin) && (nrows <= kPipeWin);
            if (ok) {
                if (lines_in_smem) {
                    if (lane < ncols) st.xp[lane] = {sm.xs[ic0 + lane], sm.xs[ic0 + lane + 1]};
                    if (lane < nrows) st.yp[lane] = {sm.ys[jr0 + lane], sm.ys[jr0 + lane + 1]};
                } else {
                    if (lane < ncols) st.xp[lane] = {__ldg(pc.xs + ic0 + lane), __ldg(pc.xs + ic0 + lane + 1)};
                    if (lane < nrows) st.yp[lane] = {__ldg(pc.ys + jr0 + lane), __ldg(pc.ys + jr0 + lane + 1)};
                }
                __syncwarp();
                // the arithmetic guess is exact +-1 on a uniform lattice; verify coverage for any other lattice
                ok = (st.xp[0].lo <= fmaxf(xmin, gx_lo)) && (st.xp[ncols - 1].hi >= fminf(xmax, gx_hi)) &&
                     (st.yp[0].lo <= fmaxf(ymin, gy_lo)) && (st.yp[nrows - 1].hi >= fminf(ymax, gy_hi));
            }
            if (lane == 0) st.hdr = {cw, sz, px, py, pz, ncols, nrows, ok ? 1 : 0};
            __syncwarp();
            if (lane == 0) {
                if (ok) {
                    // the box is always kPipeWin x kPipeWin cells; cells beyond the table are zero-filled and never read
                    bar_arrive_expect_tx(&sm.full_bar[s], (uint32_t)(kPipeWin * kPipeWin * 32));
                    tma_load_window(st.ent, &tmap, ic0, jr0, &sm.full_bar[s]);
                } else {
                    bar_arrive(&sm.full_bar[s]);
                }
            }
        }
    } else {
        // =============================== consumers ===============================
        const int group = (warp - 1) / kPipeConsumerWarps;
        const int t = threadIdx.x - 32 - group * kPipeConsumers;
        for (int it = group; it < n_iter; it += kPipeGroups) {
            const int env = (int)blockIdx.x + it * (int)gridDim.x;
            const int s = it % kPipeStages;
            const uint32_t ph = (uint32_t)(it / kPipeStages) & 1u;
            PipeStage& st = sm.stage[s];
            bar_wait(&sm.full_bar[s], ph);
            const StageHeader h = st.hdr;
            float* __restrict__ out_row = out + (size_t)env * out_stride;
            float* __restrict__ hits_row = kHits ? hits + (size_t)env * n_rays * 3 : nullptr;
            const float sz2 = __fmul_rn(h.sz, 2.f);  // fl(fl(sz*v)*2) == fl(fl(2*sz)*v): scaling by 2 is exact
            if (h.mode == 1) {
                const int cmax = h.ncols - 1, rmax = h.nrows - 1;
                const float wx0 = st.xp[0].lo, wy0 = st.yp[0].lo;
                const float inv_dx = pc.inv_dx, inv_dy = pc.inv_dy;
                for (int r0 = 0; r0 < n_rays; r0 += kPipeRaysPerThread * kPipeConsumers) {
                    // branch-free body over 4 rays: every load uses a clamped (always valid) index, stores are
                    // predicated, anything unusual is deferred -> the four dependency chains interleave (ILP)
                    unsigned defer = 0;
#pragma unroll
                    for (int u = 0; u < kPipeRaysPerThread; ++u) {
                        const int r = r0 + u * kPipeConsumers + t;
                        const bool valid = r < n_rays;
                        const float4 v = sm.pattern[valid ? r : 0];
                        // ORBIT quat_apply_yaw + pos, same roundings as ray_origin()
                        const float tx = -__fmul_rn(sz2, v.y), ty = __fmul_rn(sz2, v.x);
                        const float X = __fadd_rn(__fadd_rn(__fadd_rn(v.x, __fmul_rn(h.cw, tx)), -__fmul_rn(h.sz, ty)), h.px);
                        const float Y = __fadd_rn(__fadd_rn(__fadd_rn(v.y, __fmul_rn(h.cw, ty)), __fmul_rn(h.sz, tx)), h.py);
                        const float Z = __fadd_rn(v.z, h.pz);
                        const int ci = min(max(__float2int_rd((X - wx0) * inv_dx), 0), cmax);
                        const int cj = min(max(__float2int_rd((Y - wy0) * inv_dy), 0), rmax);
                        const LinePair2 xp = st.xp[ci], yp = st.yp[cj];
                        const int e = 2 * (cj * kPipeWin + ci);
                        const float4 q = st.ent[e + 1], p = st.ent[e];
                        const bool fast = (X >= xp.lo) & (X < xp.hi) & (Y >= yp.lo) & (Y < yp.hi) & (q.w == 0.f);
                        const float zhit = eval_cell(p, q, __fsub_rn(X, xp.lo), __fsub_rn(Y, yp.lo), Z, max_d);
                        if (valid && fast)
                            store_result(h.pz, X, Y, Z, zhit, base_offset, out_row + r, kHits ? hits_row + 3 * (size_t)r : nullptr);
                        defer |= (valid && !fast) ? (1u << u) : 0u;
                    }
                    // rare: cell guess off by one, ray on the closed far border or outside the grid, general cell
                    if (defer != 0u) {
                        for (int u = 0; u < kPipeRaysPerThread; ++u) {
                            if (!((defer >> u) & 1u)) continue;
                            const int r = r0 + u * kPipeConsumers + t;
                            const float4 v = sm.pattern[r];
                            const float tx = -__fmul_rn(sz2, v.y), ty = __fmul_rn(sz2, v.x);
                            const float X = __fadd_rn(__fadd_rn(__fadd_rn(v.x, __fmul_rn(h.cw, tx)), -__fmul_rn(h.sz, ty)), h.px);
                            const float Y = __fadd_rn(__fadd_rn(__fadd_rn(v.y, __fmul_rn(h.cw, ty)), __fmul_rn(h.sz, tx)), h.py);
                            resolve_deferred_ray(&st, g, inv_dx, inv_dy, X, Y, __fadd_rn(v.z, h.pz), h.pz, max_d, base_offset,
                                                 out_row + r, kHits ? hits_row + 3 * (size_t)r : nullptr);
                        }
                    }
                }
            } else {
                resolve_env_from_global(sm.pattern, t, n_rays, h, g, pc, max_d, base_offset, out_row, hits_row);
            }
            __syncwarp();
            if (lane == 0) bar_arrive(&sm.empty_bar[s]);  // this warp is done reading the stage
        }
    }
}

int launch_height_scan_pipelined(const float* pos_w, const float* quat_w, int n_envs, const float* ray_local,
                                 int n_rays, const ScanGridDev& g, const RoverPlaneCells* cells, float4 pattern_box,
                                 float max_d, float base_offset, float* out, int out_stride, float* hits,
                                 cudaStream_t stream) {
    ROVER_CHECK(n_rays <= kPipeMaxRays, "height_scan_pipelined: pattern larger than %d rays: use variant 2 or 3",
                kPipeMaxRays);
    static int n_sms = 0;
    static bool configured = false;
    if (!configured) {
        int dev = 0;
        ROVER_CUDA(cudaGetDevice(&dev));
        ROVER_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
        ROVER_CUDA(cudaFuncSetAttribute(height_scan_pipelined_kernel<false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PipeSmem)));
        ROVER_CUDA(cudaFuncSetAttribute(height_scan_pipelined_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PipeSmem)));
        configured = true;
    }
    alignas(64) CUtensorMap tmap;
    if (const int rc = encode_cells_tensor_map(&tmap, cells, kPipeWin, "height_scan_pipelined")) return rc;
    PlaneCellsDev pc{cells->xs, cells->ys, reinterpret_cast<const float4*>(cells->entries), cells->nx, cells->ny,
                     cells->inv_dx, cells->inv_dy};
    // every ray origin lies within this distance of the sensor position (the yaw rotation preserves norms)
    const float rx = fmaxf(fabsf(pattern_box.x), fabsf(pattern_box.y)), ry = fmaxf(fabsf(pattern_box.z), fabsf(pattern_box.w));
    const float radius = sqrtf(rx * rx + ry * ry) * 1.0001f + 1.0e-3f;
    const int grid = n_envs < n_sms * kPipeCtasPerSm ? n_envs : n_sms * kPipeCtasPerSm;
    if (hits)
        height_scan_pipelined_kernel<true><<<grid, kPipeThreads, sizeof(PipeSmem), stream>>>(
            pos_w, quat_w, n_envs, ray_local, n_rays, g, pc, tmap, radius, max_d, base_offset, out, out_stride, hits);
    else
        height_scan_pipelined_kernel<false><<<grid, kPipeThreads, sizeof(PipeSmem), stream>>>(
            pos_w, quat_w, n_envs, ray_local, n_rays, g, pc, tmap, radius, max_d, base_offset, out, out_stride, hits);
    return check_launch("height_scan_pipelined_kernel");
}

}  // namespace rover
