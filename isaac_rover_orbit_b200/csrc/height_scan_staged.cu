// Shared-memory staged height-scan kernel (variant 1).
//
// Same contract as the direct kernel (height_scan.cu), different data movement.  ncu on the direct kernel shows
// the L1TEX tag stage at ~82 % (scattered 16-byte record loads, ~20 wavefronts per request) and ~550 instructions
// per ray, most of them per-cell address / frame arithmetic.  Here one CTA owns one environment and
//   1. derives the window of level-0 home cells its ray pattern can touch (bounding box of the yaw-rotated
//      pattern rectangle + SPAN home rows/columns), from the monotone fp32 cell function;
//   2. copies that window -- a few contiguous runs of the home-sorted record array, one per home row -- into
//      shared memory, re-basing every record from its home-cell frame to ONE window frame on the way
//      (C' = C - A*dx - B*dy: 8 FMAs per record, paid once per CTA instead of per ray and cell);
//   3. walks its 961 rays against shared memory only: one (begin,end) pair per home row, then 3 x LDS.128 +
//      8 FMA per candidate, no per-cell frames, no global loads on level 0.
// Coarser levels (triangles larger than the level-0 cell) hold few records and are read straight from global
// memory.  A window that does not fit (capacity or extent) makes the whole CTA take the direct path, so the result
// never depends on the staging succeeding.
//
// Crack freedom after re-basing: the re-based constant and the evaluation in the window frame carry at most
// 2^-24 * 6 * (|A| + |B|) * W of rounding error (W = window extent in metres), so that much is added to every
// C' as outward bias (about 2.5e-6 m of fattening at W = 5 m, below the fp32 ulp of a 200 m world coordinate).
#include "scan_common.cuh"

namespace rover {

constexpr int kStagedThreads = 256;
constexpr int kRecordCap = 1280;  // records per window held in shared memory (61,440 B)
constexpr int kTableCap = 3072;   // ints for the per-row offset table (12,288 B)
constexpr int kMaxRows = 64;
constexpr size_t kStagedSmem = (size_t)kRecordCap * 48 + (size_t)kTableCap * 4;

struct Window {
    int ic0, jr0;      // first home column / row of the window (level-0 cell indices)
    int ncols, nrows;  // window extent in home cells
    int staged;        // 1: shared-memory path, 0: direct path for this CTA
    float p0x, p0y;    // window frame origin = min corner of home cell (ic0, jr0)
    float bias_k;      // 2^-24 * 6 * W
};

__device__ __forceinline__ float cast_down_levels(const ScanGridDev& g, int first_level, float X, float Y, float Z,
                                                  float max_d, float best) {
    for (int l = first_level; l < g.n_levels; ++l) {
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

__global__ void __launch_bounds__(kStagedThreads, 3)
height_scan_staged_kernel(const float* __restrict__ pos_w, const float* __restrict__ quat_w,
                          const float* __restrict__ ray_local, int n_rays, const __grid_constant__ ScanGridDev g,
                          float4 pattern_box /* xmin, xmax, ymin, ymax of the local ray pattern */, float max_d,
                          float base_offset, float* __restrict__ out, int out_stride, float* __restrict__ hits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_rec = reinterpret_cast<float4*>(smem_raw);
    int* s_tab = reinterpret_cast<int*>(smem_raw + (size_t)kRecordCap * 48);  // [nrows][ncols + 1] -> smem record index
    __shared__ SensorFrame frame_s;
    __shared__ Window win_s;
    __shared__ int s_row_base[kMaxRows + 1];

    const int env = blockIdx.x;
    const ScanLevelDev& L0 = g.level[0];
    const int S = g.span;

    if (threadIdx.x == 0) {
        const SensorFrame f = make_frame(pos_w + 3 * (size_t)env, quat_w + 4 * (size_t)env);
        frame_s = f;
        // bounding box of the rotated pattern rectangle (ray origins are affine in the local coordinates)
        float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float X, Y, Z;
            ray_origin(f, (c & 1) ? pattern_box.y : pattern_box.x, (c & 2) ? pattern_box.w : pattern_box.z, 0.f, X, Y, Z);
            xmin = fminf(xmin, X), xmax = fmaxf(xmax, X), ymin = fminf(ymin, Y), ymax = fmaxf(ymax, Y);
        }
        const float pad = 1.0e-3f;  // >> fp32 rounding of the affine map, << a cell
        Window w;
        w.ic0 = max(cell_of(xmin - pad, L0.ox, L0.inv_cell) - S, 0);
        w.jr0 = max(cell_of(ymin - pad, L0.oy, L0.inv_cell) - S, 0);
        const int ic1 = min(cell_of(xmax + pad, L0.ox, L0.inv_cell), L0.ncx - 1);
        const int jr1 = min(cell_of(ymax + pad, L0.oy, L0.inv_cell), L0.ncy - 1);
        w.ncols = ic1 - w.ic0 + 1;
        w.nrows = jr1 - w.jr0 + 1;
        w.staged = 1;
        if (w.ncols <= 0 || w.nrows <= 0) {
            w.ncols = w.nrows = 0;  // pattern entirely outside the level-0 grid: nothing to stage
        } else if (w.nrows > kMaxRows || w.nrows * (w.ncols + 1) > kTableCap) {
            w.staged = 0;
        }
        w.p0x = __fadd_rn(L0.ox, __fmul_rn((float)w.ic0, L0.cell));
        w.p0y = __fadd_rn(L0.oy, __fmul_rn((float)w.jr0, L0.cell));
        w.bias_k = 5.9604645e-8f * 6.f * fmaxf((float)max(w.ncols, w.nrows) + 1.f, 1.f) * L0.cell;
        win_s = w;
    }
    __syncthreads();
    const SensorFrame f = frame_s;
    Window w = win_s;

    if (w.staged && w.nrows > 0) {
        const int* __restrict__ cs0 = g.cell_start + L0.start_offset;
        const int tw = w.ncols + 1;
        // ---- phase 1: global record indices of the window's home cells
        for (int t = threadIdx.x; t < w.nrows * tw; t += kStagedThreads) {
            const int r = t / tw, c = t - r * tw;
            s_tab[t] = __ldg(cs0 + (size_t)(w.jr0 + r) * L0.ncx + w.ic0 + c);
        }
        __syncthreads();
        // ---- phase 2: shared-memory base of every home row (exclusive scan of the row run lengths)
        if (threadIdx.x < 32) {
            int carry = 0;
            for (int r0 = 0; r0 < w.nrows; r0 += 32) {
                const int r = r0 + threadIdx.x;
                const int len = (r < w.nrows) ? (s_tab[r * tw + w.ncols] - s_tab[r * tw]) : 0;
                int inc = len;
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if ((int)threadIdx.x >= o) inc += v;
                }
                if (r < w.nrows) s_row_base[r] = carry + inc - len;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (threadIdx.x == 0) {
                s_row_base[w.nrows] = carry;
                if (carry > kRecordCap) win_s.staged = 0;
            }
        }
        __syncthreads();
        w.staged = win_s.staged;
    }

    if (w.staged && w.nrows > 0) {
        const int tw = w.ncols + 1;
        // ---- phase 3: copy + re-base the records, one home cell per thread slot
        for (int t = threadIdx.x; t < w.nrows * w.ncols; t += kStagedThreads) {
            const int r = t / w.ncols, c = t - r * w.ncols;
            const int gb = s_tab[r * tw + c], ge = s_tab[r * tw + c + 1];
            const int dst0 = s_row_base[r] + (gb - s_tab[r * tw]);
            // frame shift home cell -> window origin (both are fp32 cell corners; the difference is exact)
            const float dx = __fsub_rn(__fadd_rn(L0.ox, __fmul_rn((float)(w.ic0 + c), L0.cell)), w.p0x);
            const float dy = __fsub_rn(__fadd_rn(L0.oy, __fmul_rn((float)(w.jr0 + r), L0.cell)), w.p0y);
            for (int q = gb; q < ge; ++q) {
                float4 r0 = __ldg(g.rec + 3 * (size_t)q), r1 = __ldg(g.rec + 3 * (size_t)q + 1),
                       r2 = __ldg(g.rec + 3 * (size_t)q + 2);
                r0.z = fmaf(-r0.x, dx, fmaf(-r0.y, dy, r0.z)) + w.bias_k * (fabsf(r0.x) + fabsf(r0.y));
                r1.y = fmaf(-r0.w, dx, fmaf(-r1.x, dy, r1.y)) + w.bias_k * (fabsf(r0.w) + fabsf(r1.x));
                r2.x = fmaf(-r1.z, dx, fmaf(-r1.w, dy, r2.x)) + w.bias_k * (fabsf(r1.z) + fabsf(r1.w));
                r2.w = fmaf(-r2.y, dx, fmaf(-r2.z, dy, r2.w));
                float4* d = s_rec + 3 * (size_t)(dst0 + (q - gb));
                d[0] = r0;
                d[1] = r1;
                d[2] = r2;
            }
        }
        __syncthreads();
        // ---- phase 4: turn the table into shared-memory record indices (in place; each entry owned by one thread)
        for (int t = threadIdx.x; t < w.nrows * tw; t += kStagedThreads) {
            const int r = t / tw;
            // s_tab[r*tw] is the row's first global index; it is needed by every entry of the row, so rows are
            // converted back to front: entry c > 0 first, entry 0 last (done by the thread that owns it, after a sync)
            if (t - r * tw != 0) s_tab[t] = s_row_base[r] + (s_tab[t] - s_tab[r * tw]);
        }
        __syncthreads();
        for (int r = threadIdx.x; r < w.nrows; r += kStagedThreads) s_tab[r * tw] = s_row_base[r];
        __syncthreads();
    }

    // ---- rays
    for (int r = threadIdx.x; r < n_rays; r += kStagedThreads) {
        const float vx = __ldg(ray_local + 3 * r), vy = __ldg(ray_local + 3 * r + 1), vz = __ldg(ray_local + 3 * r + 2);
        float X, Y, Z;
        ray_origin(f, vx, vy, vz, X, Y, Z);
        float best = -INFINITY;
        if (w.staged) {
            if (w.nrows > 0) {
                const int tw = w.ncols + 1;
                const int ci = cell_of(X, L0.ox, L0.inv_cell) - w.ic0;
                const int cj = cell_of(Y, L0.oy, L0.inv_cell) - w.jr0;
                const float lx = __fsub_rn(X, w.p0x), ly = __fsub_rn(Y, w.p0y);
                const int c0 = max(ci - S, 0), c1 = min(ci, w.ncols - 1);
                const int r0 = max(cj - S, 0), r1 = min(cj, w.nrows - 1);
                if (c0 <= c1) {
                    for (int rr = r0; rr <= r1; ++rr) {
                        const int b = s_tab[rr * tw + c0], e = s_tab[rr * tw + c1 + 1];
                        for (int q = b; q < e; ++q)
                            test_record(s_rec[3 * q], s_rec[3 * q + 1], s_rec[3 * q + 2], lx, ly, Z, max_d, best);
                    }
                }
            }
            if (g.n_levels > 1) best = cast_down_levels(g, 1, X, Y, Z, max_d, best);
        } else {
            best = cast_down_levels(g, 0, X, Y, Z, max_d, best);
        }
        float h = -INFINITY, hx = INFINITY, hy = INFINITY, hz = INFINITY;
        if (best != -INFINITY) {
            const float t = __fsub_rn(Z, best);
            hz = __fsub_rn(Z, t);
            hx = X;
            hy = Y;
            h = __fsub_rn(__fsub_rn(f.pz, hz), base_offset);
        }
        out[(size_t)env * out_stride + r] = h;
        if (hits) {
            float* p = hits + ((size_t)env * n_rays + r) * 3;
            p[0] = hx;
            p[1] = hy;
            p[2] = hz;
        }
    }
}

int launch_height_scan_staged(const float* pos_w, const float* quat_w, int n_envs, const float* ray_local, int n_rays,
                              const ScanGridDev& g, float4 pattern_box, float max_d, float base_offset, float* out,
                              int out_stride, float* hits, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        ROVER_CUDA(cudaFuncSetAttribute(height_scan_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kStagedSmem));
        configured = true;
    }
    height_scan_staged_kernel<<<n_envs, kStagedThreads, kStagedSmem, stream>>>(
        pos_w, quat_w, ray_local, n_rays, g, pattern_box, max_d, base_offset, out, out_stride, hits);
    return check_launch("height_scan_staged_kernel");
}

}  // namespace rover
