// Plane-cell height scan with the per-environment window staged through the bulk-copy (TMA) engine -- variant 3.
//
// ncu on variant 2 (height_scan_cells_kernel): nothing saturated (L1TEX 32 %, issue 40 %, DRAM 4 %), 27 % of the
// stall samples sit on the first use of the table entry -- every ray chases ~7 dependent global loads (grid-line
// look-ups, then the entry, which comes from HBM when L2 is cold).  The window of table entries an environment can
// touch is a dense 2-D tile of the row-major table (<= 32 rows of <= 32 cells x 32 B), so one CTA per environment
//   1. computes its sensor frame and the window (bounding box of the yaw-rotated pattern rectangle);
//   2. has warp 0 issue one cp.async.bulk (UBLKCP) per window row into shared memory, completion on an mbarrier,
//      while the other warps load the window's grid lines and compute their ray origins;
//   3. resolves every ray from shared memory only: locate in the window lines, 2 x LDS.128, 5 FMA.
// Rays in general cells (tag != 0) and CTAs whose window exceeds the shared tile take the global-memory paths,
// so results never depend on the staging.
#include "scan_common.cuh"

namespace rover {

constexpr int kTmaThreads = 128;  // small CTAs: >= 8 resident per SM so that prologues overlap ray phases
constexpr int kWinMax = 28;       // window cells per axis held in shared memory (28 x 28 x 32 B = 24.5 KB)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// index of the half-open interval [lines[i], lines[i+1]) containing v, i in [0, n-1]; lines may live in shared
// or global memory.  `guess` is an arithmetic first guess (exact or off by one on a uniform lattice); both bounds
// of the guessed interval are fetched together and the fix-up loops (exact for any monotone lines) rarely run.
// `lo_out` receives lines[i], the cell's frame origin.
__device__ __forceinline__ int locate_in(const float* lines, int n, int guess, float v, float& lo_out) {
    int i = min(max(guess, 0), n - 1);
    float a = lines[i];
    const float b = lines[i + 1];
    if (v < a || v >= b) {
        while (i > 0 && v < lines[i]) --i;
        while (i < n - 1 && v >= lines[i + 1]) ++i;
        a = lines[i];
    }
    lo_out = a;
    return i;
}

__device__ __forceinline__ int guess_cell(float v, float lo, float inv_d) {
    return (int)fminf(fmaxf(floorf((v - lo) * inv_d), -1.0e6f), 1.0e6f);
}

__device__ __forceinline__ float eval_entry(const float4 p, const float4 q, float lx, float ly, float Z, float max_d) {
    const float E = fmaf(q.x, lx, fmaf(q.y, ly, q.z));
    const float z = fmaf(p.w, fminf(E, 0.f), fmaf(p.x, lx, fmaf(p.y, ly, p.z)));
    const float t = Z - z;
    return (t >= 0.f && t < max_d) ? z : -INFINITY;
}

// home-grid walk for rays of general cells (same code as the direct kernel)
__device__ __noinline__ float cast_down_slow(const ScanGridDev& g, float X, float Y, float Z, float max_d) {
    float best = -INFINITY;
    for (int l = 0; l < g.n_levels; ++l) {
        const ScanLevelDev& L = g.level[l];
        const int i = cell_of(X, L.ox, L.inv_cell);
        const int j = cell_of(Y, L.oy, L.inv_cell);
        const int j0 = max(j - g.span, 0), j1 = min(j, L.ncy - 1);
        const int i0 = max(i - g.span, 0), i1 = min(i, L.ncx - 1);
        for (int jj = j0; jj <= j1; ++jj) {
            const float ly = __fsub_rn(Y, __fadd_rn(L.oy, __fmul_rn((float)jj, L.cell)));
            const int* __restrict__ row = g.cell_start + L.start_offset + jj * L.ncx;
            for (int ii = i0; ii <= i1; ++ii) {
                const float lx = __fsub_rn(X, __fadd_rn(L.ox, __fmul_rn((float)ii, L.cell)));
                const int b = __ldg(row + ii), e = __ldg(row + ii + 1);
                for (int r = b; r < e; ++r)
                    test_record(__ldg(g.rec + 3 * r), __ldg(g.rec + 3 * r + 1), __ldg(g.rec + 3 * r + 2), lx, ly, Z,
                                max_d, best);
            }
        }
    }
    return best;
}

// reference rounding chain: t -> hit.z = Z + t*(-1) -> (pos.z - hit.z) - offset ; a miss is -inf / +inf
__device__ __forceinline__ void write_result(const SensorFrame f, float X, float Y, float Z, float zhit,
                                             float base_offset, float* __restrict__ out, float* __restrict__ hit3) {
    float h = -INFINITY, hx = INFINITY, hy = INFINITY, hz = INFINITY;
    if (zhit != -INFINITY) {
        const float t = __fsub_rn(Z, zhit);
        hz = __fsub_rn(Z, t);
        hx = X;
        hy = Y;
        h = __fsub_rn(__fsub_rn(f.pz, hz), base_offset);
    }
    *out = h;
    if (hit3) {
        hit3[0] = hx;
        hit3[1] = hy;
        hit3[2] = hz;
    }
}

struct CellWindow {
    int ic0, jr0, ncols, nrows;
    float xmin, xmax, ymin, ymax;  // padded bounding box of the ray origins
};

// whole-environment fallback: the table is read from global memory (window larger than the shared tile, or the
// arithmetic window guess did not cover the pattern on a non-uniform lattice).  Rare, so kept out of line.
__device__ __noinline__ void scan_env_from_global(const SensorFrame f, const float* __restrict__ ray_local,
                                                   int n_rays, const ScanGridDev& g, const PlaneCellsDev& pc,
                                                   float max_d, float base_offset, float* __restrict__ out_row,
                                                   float* __restrict__ hits_row) {
    const float gx_lo = __ldg(pc.xs), gx_hi = __ldg(pc.xs + pc.nx), gy_lo = __ldg(pc.ys), gy_hi = __ldg(pc.ys + pc.ny);
    for (int r = threadIdx.x; r < n_rays; r += kTmaThreads) {
        float X, Y, Z;
        ray_origin(f, __ldg(ray_local + 3 * r), __ldg(ray_local + 3 * r + 1), __ldg(ray_local + 3 * r + 2), X, Y, Z);
        float zhit = -INFINITY;
        if ((X >= gx_lo) && (X <= gx_hi) && (Y >= gy_lo) && (Y <= gy_hi)) {
            float x0, y0;
            const int i = locate_in(pc.xs, pc.nx, guess_cell(X, gx_lo, pc.inv_dx), X, x0);
            const int j = locate_in(pc.ys, pc.ny, guess_cell(Y, gy_lo, pc.inv_dy), Y, y0);
            const float4* __restrict__ e = pc.ent + 2 * ((size_t)j * pc.nx + i);
            const float4 p = __ldg(e), q = __ldg(e + 1);
            zhit = (q.w == 0.f) ? eval_entry(p, q, __fsub_rn(X, x0), __fsub_rn(Y, y0), Z, max_d)
                                : cast_down_slow(g, X, Y, Z, max_d);
        }
        write_result(f, X, Y, Z, zhit, base_offset, out_row + r, hits_row ? hits_row + 3 * (size_t)r : nullptr);
    }
}

// Hot-loop helpers index the __shared__ arrays directly (shared::cta addressing, no generic pointers).
struct LinePair {
    float lo, hi;  // lines[i], lines[i+1]: one LDS.64 answers "is v in cell i?"
};

template <bool kHits>
__global__ void __launch_bounds__(kTmaThreads, 8)
height_scan_cells_tma_kernel(const float* __restrict__ pos_w, const float* __restrict__ quat_w,
                             const float* __restrict__ ray_local, int n_rays, const __grid_constant__ ScanGridDev g,
                             const __grid_constant__ PlaneCellsDev pc, float4 pattern_box, float max_d,
                             float base_offset, float* __restrict__ out, int out_stride, float* __restrict__ hits) {
    __shared__ __align__(128) float4 s_ent[kWinMax * kWinMax * 2];  // [row][col][2]
    __shared__ __align__(8) LinePair s_xp[kWinMax], s_yp[kWinMax];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ SensorFrame frame_s;
    __shared__ CellWindow win_s;
    __shared__ int s_staged;

    const int env = blockIdx.x;
    float* __restrict__ out_row = out + (size_t)env * out_stride;
    float* __restrict__ hits_row = kHits ? hits + (size_t)env * n_rays * 3 : nullptr;
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        const float gx_lo = __ldg(pc.xs), gy_lo = __ldg(pc.ys);
        const SensorFrame f = make_frame(pos_w + 3 * (size_t)env, quat_w + 4 * (size_t)env);
        frame_s = f;
        CellWindow w;
        w.xmin = INFINITY, w.xmax = -INFINITY, w.ymin = INFINITY, w.ymax = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // ray origins are affine in the local pattern coordinates
            float X, Y, Z;
            ray_origin(f, (c & 1) ? pattern_box.y : pattern_box.x, (c & 2) ? pattern_box.w : pattern_box.z, 0.f, X, Y, Z);
            w.xmin = fminf(w.xmin, X), w.xmax = fmaxf(w.xmax, X), w.ymin = fminf(w.ymin, Y), w.ymax = fmaxf(w.ymax, Y);
        }
        const float pad = 1.0e-3f;  // >> fp32 rounding of the affine map, << a cell
        w.xmin -= pad, w.xmax += pad, w.ymin -= pad, w.ymax += pad;
        // arithmetic guess +- one cell; verified against the real lines once they are in shared memory
        w.ic0 = min(max(guess_cell(w.xmin, gx_lo, pc.inv_dx) - 1, 0), pc.nx - 1);
        w.jr0 = min(max(guess_cell(w.ymin, gy_lo, pc.inv_dy) - 1, 0), pc.ny - 1);
        const int ic1 = min(max(guess_cell(w.xmax, gx_lo, pc.inv_dx) + 1, 0), pc.nx - 1);
        const int jr1 = min(max(guess_cell(w.ymax, gy_lo, pc.inv_dy) + 1, 0), pc.ny - 1);
        w.ncols = ic1 - w.ic0 + 1;
        w.nrows = jr1 - w.jr0 + 1;
        s_staged = (w.ncols <= kWinMax && w.nrows <= kWinMax) ? 1 : 0;
        win_s = w;
    }
    __syncthreads();
    const int ncols = win_s.ncols, nrows = win_s.nrows;
    const bool fits = s_staged != 0;

    if (fits) {
        if (threadIdx.x < 32) {
            // warp 0: arm the barrier with the byte count, then one bulk copy (UBLKCP) per window row
            const uint32_t row_bytes = (uint32_t)ncols * 32u;
            if (threadIdx.x == 0) mbar_expect_tx(&s_bar, row_bytes * (uint32_t)nrows);
            __syncwarp();
            const int ic0 = win_s.ic0, jr0 = win_s.jr0;
            for (int r = threadIdx.x; r < nrows; r += 32)
                bulk_g2s(s_ent + (size_t)r * ncols * 2, pc.ent + 2 * ((size_t)(jr0 + r) * pc.nx + ic0), row_bytes, &s_bar);
        } else if (threadIdx.x < 64) {
            const int c = threadIdx.x - 32;
            if (c < ncols) s_xp[c] = {__ldg(pc.xs + win_s.ic0 + c), __ldg(pc.xs + win_s.ic0 + c + 1)};
        } else if (threadIdx.x < 96) {
            const int r = threadIdx.x - 64;
            if (r < nrows) s_yp[r] = {__ldg(pc.ys + win_s.jr0 + r), __ldg(pc.ys + win_s.jr0 + r + 1)};
        }
    }
    __syncthreads();  // grid lines visible; the table tile may still be in flight

    const SensorFrame f = frame_s;
    // does the window cover every ray origin that lies inside the grid?  (always, on a uniform lattice)
    bool covered = fits;
    if (fits) {
        const float gx_lo = __ldg(pc.xs), gx_hi = __ldg(pc.xs + pc.nx), gy_lo = __ldg(pc.ys), gy_hi = __ldg(pc.ys + pc.ny);
        covered = (s_xp[0].lo <= fmaxf(win_s.xmin, gx_lo)) && (s_xp[ncols - 1].hi >= fminf(win_s.xmax, gx_hi)) &&
                  (s_yp[0].lo <= fmaxf(win_s.ymin, gy_lo)) && (s_yp[nrows - 1].hi >= fminf(win_s.ymax, gy_hi));
    }
    if (!covered) {
        if (fits) mbar_wait(&s_bar, 0);  // the tile must land before this CTA's shared memory is released
        scan_env_from_global(f, ray_local, n_rays, g, pc, max_d, base_offset, out_row, hits_row);
        return;
    }

    const float wx0 = s_xp[0].lo, wy0 = s_yp[0].lo;
    const float wx1 = s_xp[ncols - 1].hi, wy1 = s_yp[nrows - 1].hi;  // == grid border when the window is clamped
    const float sz2 = __fmul_rn(f.sz, 2.f);  // fl(fl(sz*v)*2) == fl(fl(2*sz)*v): scaling by 2 is exact
    const float inv_dx = pc.inv_dx, inv_dy = pc.inv_dy;
    const int cmax = ncols - 1, rmax = nrows - 1;
    mbar_wait(&s_bar, 0);

    unsigned slow_mask = 0;  // bit k: the k-th ray of this thread sits in a general cell (second pass below)
    int k = 0;
#pragma unroll 2
    for (int r = threadIdx.x; r < n_rays; r += kTmaThreads, ++k) {
        const float vx = __ldg(ray_local + 3 * r), vy = __ldg(ray_local + 3 * r + 1), vz = __ldg(ray_local + 3 * r + 2);
        // ORBIT quat_apply_yaw + pos (same roundings as ray_origin())
        const float tx = -__fmul_rn(sz2, vy), ty = __fmul_rn(sz2, vx);
        const float X = __fadd_rn(__fadd_rn(__fadd_rn(vx, __fmul_rn(f.cw, tx)), -__fmul_rn(f.sz, ty)), f.px);
        const float Y = __fadd_rn(__fadd_rn(__fadd_rn(vy, __fmul_rn(f.cw, ty)), __fmul_rn(f.sz, tx)), f.py);
        const float Z = __fadd_rn(vz, f.pz);
        int ci = min(max((int)floorf((X - wx0) * inv_dx), 0), cmax);
        int cj = min(max((int)floorf((Y - wy0) * inv_dy), 0), rmax);
        LinePair xp = s_xp[ci], yp = s_yp[cj];
        if (!(X >= xp.lo && X < xp.hi && Y >= yp.lo && Y < yp.hi)) {
            // rare: guess off by one, ray on the closed far border of the grid, or ray outside the grid
            if (!(X >= wx0 && X <= wx1 && Y >= wy0 && Y <= wy1)) {
                write_result(f, X, Y, Z, -INFINITY, base_offset, out_row + r, kHits ? hits_row + 3 * (size_t)r : nullptr);
                continue;
            }
            while (ci > 0 && X < s_xp[ci].lo) --ci;
            while (ci < cmax && X >= s_xp[ci].hi) ++ci;
            while (cj > 0 && Y < s_yp[cj].lo) --cj;
            while (cj < rmax && Y >= s_yp[cj].hi) ++cj;
            xp = s_xp[ci], yp = s_yp[cj];
        }
        const int e = 2 * (cj * ncols + ci);
        const float4 q = s_ent[e + 1];
        if (q.w != 0.f) {
            slow_mask |= 1u << (k & 31);
            continue;
        }
        const float4 p = s_ent[e];
        const float zhit = eval_entry(p, q, __fsub_rn(X, xp.lo), __fsub_rn(Y, yp.lo), Z, max_d);
        write_result(f, X, Y, Z, zhit, base_offset, out_row + r, kHits ? hits_row + 3 * (size_t)r : nullptr);
    }

    // second pass: rays of general cells walk the home grid (kept out of the hot loop: it needs many registers)
    if (slow_mask != 0u || k > 32) {
        k = 0;
        for (int r = threadIdx.x; r < n_rays; r += kTmaThreads, ++k) {
            if (k < 32 && !((slow_mask >> k) & 1u)) continue;
            float X, Y, Z;
            ray_origin(f, __ldg(ray_local + 3 * r), __ldg(ray_local + 3 * r + 1), __ldg(ray_local + 3 * r + 2), X, Y, Z);
            if (k >= 32) {  // more than 32 rays per thread: the mask wrapped, re-classify from the table
                if (!(X >= wx0 && X <= wx1 && Y >= wy0 && Y <= wy1)) continue;
                float x0, y0;
                const int i = locate_in(pc.xs, pc.nx, guess_cell(X, __ldg(pc.xs), pc.inv_dx), X, x0);
                const int j = locate_in(pc.ys, pc.ny, guess_cell(Y, __ldg(pc.ys), pc.inv_dy), Y, y0);
                if (__ldg(pc.ent + 2 * ((size_t)j * pc.nx + i) + 1).w == 0.f) continue;
            }
            const float zhit = cast_down_slow(g, X, Y, Z, max_d);
            write_result(f, X, Y, Z, zhit, base_offset, out_row + r, kHits ? hits_row + 3 * (size_t)r : nullptr);
        }
    }
}

int launch_height_scan_cells_tma(const float* pos_w, const float* quat_w, int n_envs, const float* ray_local,
                                 int n_rays, const ScanGridDev& g, const RoverPlaneCells* cells, float4 pattern_box,
                                 float max_d, float base_offset, float* out, int out_stride, float* hits,
                                 cudaStream_t stream) {
    PlaneCellsDev pc{cells->xs, cells->ys, reinterpret_cast<const float4*>(cells->entries), cells->nx, cells->ny,
                     cells->inv_dx, cells->inv_dy};
    if (hits)
        height_scan_cells_tma_kernel<true><<<n_envs, kTmaThreads, 0, stream>>>(
            pos_w, quat_w, ray_local, n_rays, g, pc, pattern_box, max_d, base_offset, out, out_stride, hits);
    else
        height_scan_cells_tma_kernel<false><<<n_envs, kTmaThreads, 0, stream>>>(
            pos_w, quat_w, ray_local, n_rays, g, pc, pattern_box, max_d, base_offset, out, out_stride, hits);
    return check_launch("height_scan_cells_tma_kernel");
}

}  // namespace rover
