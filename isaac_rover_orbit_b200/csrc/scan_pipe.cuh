// Pieces shared by the TMA-pipelined height-scan kernels (variants 4 and 5): mbarrier / TMA wrappers, the plane-cell
// evaluation, the reference's result rounding chain and the out-of-line global-memory fallbacks.
#pragma once
#include <cuda.h>  // CUtensorMap types only; the encoder is resolved at run time (no link dependency on libcuda)

#include "scan_common.cuh"

namespace rover {

struct LinePair2 {
    float lo, hi;
};

struct StageHeader {
    float cw, sz, px, py, pz;
    int ncols, nrows;
    int mode;  // 1: window staged in shared memory, 0: consumers read the table from global memory
};

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_addr(bar)) : "memory");
}
__device__ __forceinline__ void bar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(s_addr(bar)), "r"(parity)
            : "memory");
    }
}
// one 2-D TMA tile load (UTMALDG): box (set in the tensor map) = win rows x win cells of the table, starting at cell (col, row)
__device__ __forceinline__ void tma_load_window(void* dst, const CUtensorMap* tmap, int col, int row,
                                                unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(s_addr(dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(col * 8), "r"(row), "r"(s_addr(bar))
        : "memory");
}

__device__ __forceinline__ int guess_col(float v, float lo, float inv_d) {
    return (int)fminf(fmaxf(floorf((v - lo) * inv_d), -1.0e6f), 1.0e6f);
}

__device__ __forceinline__ float eval_cell(const float4 p, const float4 q, float lx, float ly, float Z, float max_d) {
    const float E = fmaf(q.x, lx, fmaf(q.y, ly, q.z));
    const float z = fmaf(p.w, fminf(E, 0.f), fmaf(p.x, lx, fmaf(p.y, ly, p.z)));
    const float t = Z - z;
    return (t >= 0.f && t < max_d) ? z : -INFINITY;
}

__device__ __forceinline__ void store_result(float pz, float X, float Y, float Z, float zhit, float base_offset,
                                             float* __restrict__ out, float* __restrict__ hit3) {
    float h = -INFINITY, hx = INFINITY, hy = INFINITY, hz = INFINITY;
    if (zhit != -INFINITY) {
        const float t = __fsub_rn(Z, zhit);  // reference rounding chain: t, hit.z = Z - t, (pos.z - hit.z) - offset
        hz = __fsub_rn(Z, t);
        hx = X;
        hy = Y;
        h = __fsub_rn(__fsub_rn(pz, hz), base_offset);
    }
    *out = h;
    if (hit3) {
        hit3[0] = hx;
        hit3[1] = hy;
        hit3[2] = hz;
    }
}

// home-grid walk (general cells)
static __device__ __noinline__ float walk_home_grid(const ScanGridDev& g, float X, float Y, float Z, float max_d) {
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

// one ray resolved from the table in global memory (fallback environments and general cells)
static __device__ __noinline__ float resolve_from_global(const ScanGridDev& g, const PlaneCellsDev& pc, float X, float Y,
                                                   float Z, float max_d) {
    const float gx_lo = __ldg(pc.xs), gx_hi = __ldg(pc.xs + pc.nx), gy_lo = __ldg(pc.ys), gy_hi = __ldg(pc.ys + pc.ny);
    if (!((X >= gx_lo) && (X <= gx_hi) && (Y >= gy_lo) && (Y <= gy_hi))) return -INFINITY;
    int i = min(max(guess_col(X, gx_lo, pc.inv_dx), 0), pc.nx - 1);
    int j = min(max(guess_col(Y, gy_lo, pc.inv_dy), 0), pc.ny - 1);
    while (i > 0 && X < __ldg(pc.xs + i)) --i;
    while (i < pc.nx - 1 && X >= __ldg(pc.xs + i + 1)) ++i;
    while (j > 0 && Y < __ldg(pc.ys + j)) --j;
    while (j < pc.ny - 1 && Y >= __ldg(pc.ys + j + 1)) ++j;
    const float4* __restrict__ e = pc.ent + 2 * ((size_t)j * pc.nx + i);
    const float4 p = __ldg(e), q = __ldg(e + 1);
    if (q.w != 0.f) return walk_home_grid(g, X, Y, Z, max_d);
    return eval_cell(p, q, __fsub_rn(X, __ldg(pc.xs + i)), __fsub_rn(Y, __ldg(pc.ys + j)), Z, max_d);
}

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                           CUtensorMapFloatOOBfill);

// 2-D tensor map over the plane-cell table: rows of nx cells x 8 floats; box = win x win cells.  Returns 0 on success.
inline int encode_cells_tensor_map(CUtensorMap* tmap, const RoverPlaneCells* cells, int win, const char* who) {
    static TensorMapEncodeTiledFn encode = nullptr;
    if (encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        ROVER_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        ROVER_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess,
                    "%s: cuTensorMapEncodeTiled is not available in this driver", who);
        encode = reinterpret_cast<TensorMapEncodeTiledFn>(fn);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cells->nx * 8ull, (cuuint64_t)cells->ny};
    const cuuint64_t gstride[1] = {(cuuint64_t)cells->nx * 32ull};
    const cuuint32_t box[2] = {(cuuint32_t)win * 8u, (cuuint32_t)win};
    const cuuint32_t estride[2] = {1u, 1u};
    const CUresult rc = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(cells->entries), gdim, gstride,
                               box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ROVER_CHECK(rc == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed (%d)", who, (int)rc);
    return 0;
}

// 3-D tensor map over the planar copy of the table [2, ny, nx * 4 floats]; box = win_cols cells x win_rows rows x 2.
inline int encode_planar_tensor_map(CUtensorMap* tmap, const RoverPlaneCells* cells, int win_cols, int win_rows,
                                    const char* who) {
    static TensorMapEncodeTiledFn encode = nullptr;
    if (encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        ROVER_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        ROVER_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess,
                    "%s: cuTensorMapEncodeTiled is not available in this driver", who);
        encode = reinterpret_cast<TensorMapEncodeTiledFn>(fn);
    }
    ROVER_CHECK((reinterpret_cast<uintptr_t>(cells->entries_planar) & 15) == 0, "%s: entries_planar not 16B aligned", who);
    const cuuint64_t gdim[3] = {(cuuint64_t)cells->nx * 4ull, (cuuint64_t)cells->ny, 2ull};
    const cuuint64_t gstride[2] = {(cuuint64_t)cells->nx * 16ull, (cuuint64_t)cells->nx * 16ull * (cuuint64_t)cells->ny};
    const cuuint32_t box[3] = {(cuuint32_t)win_cols * 4u, (cuuint32_t)win_rows, 2u};
    const cuuint32_t estride[3] = {1u, 1u, 1u};
    const CUresult rc = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(cells->entries_planar), gdim,
                               gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ROVER_CHECK(rc == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled (planar) failed (%d)", who, (int)rc);
    return 0;
}

}  // namespace rover
