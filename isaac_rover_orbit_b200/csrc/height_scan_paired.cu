// Persistent, warp-specialised plane-cell height scan with packed-fp32 ray pairs -- variant 5.
//
// Why (ncu on variant 4, profiles/r01_scan_v4*.md + r01_scan_v9 source page): the kernel is issue-bound, not HBM-bound.
//   * 13 % of the warp time was the serial table fill in the prologue (3 dependent cold misses per thread),
//   * 26 % was consumers waiting on the first windows (the producer started only after that prologue),
//   * 5 x 224 thread slots per 961-ray environment left 14 % of the lanes predicated off, the two 7-warp groups ended
//     up to one environment apart, and the per-ray store path recomputed a 64-bit address and a double select.
//
// What changes (same producer / mbarrier ring / 2-D tensor-map TMA loads as variant 4):
//   * Work unit = chunk of 256 consecutive rays of one environment; the 14 consumer warps take chunks round-robin over
//     the CTA's whole run (chunk K -> warp K mod 14), so lanes are 94 % used at 961 rays and every warp gets the same
//     number of chunks; a stage's `empty` barrier counts one arrival per chunk.
//   * A lane resolves ray PAIRS (2k, 2k+1): the ray pattern sits in shared memory as three float arrays, so one LDS.64
//     yields the same coordinate of both rays, and the whole ORBIT rotation chain, the window-relative cell coordinate
//     and the reference's result rounding chain run on FADD2 / FMUL2 (add/mul.rn.f32x2 -- per-lane IEEE fp32, so the
//     roundings are exactly those of the scalar chain) -- half the issue slots of the scalar code.
//   * floor() to a cell index is one FADD2.RM against 1.5 * 2^23 (the integer lands in the mantissa) instead of two
//     F2I on the quarter-rate XU pipe; clamping is done on the raw bits.
//   * Grid-line pairs for the whole table live in shared memory once per SM (no per-stage copies by the producer);
//     the prologue issues all of its global loads before the first store (one cold-miss latency instead of three).
// Rays in general cells, cell guesses that miss, windows that do not fit or do not cover (non-uniform lattices) take
// the same out-of-line global-memory helpers as variant 4, so the staging can never change a result.
#include <cuda_bf16.h>

#include "scan_pipe.cuh"

#ifndef ROVER_SCAN_DBG
#define ROVER_SCAN_DBG 0  // bring-up experiments only: 1 = consumers skip the ray work, 2 = TMA loads only for the first ring pass
#endif

namespace rover {

#if ROVER_SCAN_DBG == 3  // timeline of CTA 0 and CTA 100 in SM cycles since the CTA started (bring-up builds only)
__device__ unsigned long long g_scan_dbg[2][512];
__device__ unsigned long long g_scan_cta[2][256];  // [0][b] = globaltimer at CTA start, [1][b] = at its last warp's end
__device__ __forceinline__ unsigned long long dbg_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define DBG_STAMP(slot)                                                                             \
    do {                                                                                            \
        if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 100) && (slot) < 512)                    \
            g_scan_dbg[blockIdx.x != 0][slot] = (unsigned long long)(clock64() - dbg_t0);           \
    } while (0)
#define DBG_STAMP_ANY(slot)                                                                         \
    do {                                                                                            \
        if ((blockIdx.x == 0 || blockIdx.x == 100) && (slot) < 512)                                 \
            g_scan_dbg[blockIdx.x != 0][slot] = (unsigned long long)(clock64() - dbg_t0);           \
    } while (0)
#else
#define DBG_STAMP(slot) \
    do {                \
    } while (0)
#define DBG_STAMP_ANY(slot) \
    do {                    \
    } while (0)
#endif

#ifndef ROVER_PAIR_WARPS
#define ROVER_PAIR_WARPS 14
#endif
constexpr int kPairConsumerWarps = ROVER_PAIR_WARPS;  // 1 + 14 warps -> 16-warp register allocation -> 128 registers / thread
constexpr int kPairThreads = 32 * (1 + kPairConsumerWarps);
constexpr int kPairStages = 8;          // ring depth (data)
#ifndef ROVER_PAIR_EARLY
#define ROVER_PAIR_EARLY 8
#endif
constexpr int kPairEarly = ROVER_PAIR_EARLY;  // windows whose TMA load is issued in the prologue (<= kPairStages)
constexpr int kPairFullBars = 16;       // `full` barriers: two per stage, see the note on phase aliasing below
constexpr int kPairWin = 26;            // window cells per axis
constexpr int kPairPitch = 27;          // cells per staged row: odd, so that consecutive rows shift by one 16-byte bank
                                        // group and a quarter-warp's LDS.128 of 8 neighbouring cells is conflict free
constexpr int kPairPlane = kPairWin * kPairPitch;             // float4 per plane
constexpr uint32_t kPairStageBytes = 2u * kPairPlane * 16u;   // bytes one TMA load delivers
#ifndef ROVER_PAIR_CHUNK
#define ROVER_PAIR_CHUNK 256
#endif
constexpr int kPairBatch = 128;         // rays per pass of a warp: 2 pair slots x 32 lanes x 2 rays
constexpr int kPairChunk = ROVER_PAIR_CHUNK;  // rays per work unit (a whole number of batches)
static_assert(kPairChunk % kPairBatch == 0, "a chunk is a whole number of batches");
constexpr int kPairMaxRays = 1024;      // pattern held in shared memory (3 float arrays)
constexpr int kPairMaxLines = 1024;     // grid-line pairs per axis held in shared memory
constexpr float kFloorMagic = 12582912.0f;  // 1.5 * 2^23: fl_rm(v + magic) has floor(v) in its mantissa for |v| < 2^22
constexpr int kFloorMagicBits = 0x4B400000;

// Lane <-> ray mapping inside a batch of 128 rays [b0, b0 + 128): lane l, slot u (0/1) resolves the pair
//   A = b0 + 32u + l,  B = A + 64.
// One instruction therefore serves 32 CONSECUTIVE rays (0.1 m apart: a quarter-warp touches ~4 cells, not 8, and
// its stores fill whole sectors), while the two rays of a pair still come out of one LDS.64: the pattern arrays are
// held in slot order, slot(A) = b0 + 64u + 2l, slot(B) = slot(A) + 1.
__host__ __device__ __forceinline__ int ray_of_slot(int s) {
    const int j = s & 127;
    return (s & ~127) + 64 * (j & 1) + 32 * (j >> 6) + ((j & 63) >> 1);
}
__host__ __device__ __forceinline__ int slot_of_ray(int r) {
    const int j = r & 127;
    return (r & ~127) + 64 * ((j >> 5) & 1) + 2 * (j & 31) + (j >> 6);
}

struct PairHeader {
    float cw, sz, px, py;
    float pz;
    int ic0, jr0, ncols;
    int nrows, mode, pad0, pad1;  // mode 1: window staged in shared memory, 0: read the table from global memory
};

// Planes of the window: p[row][col] = (a, b, c, k), q[row][col] = (A, B, C, tag) of cell (jr0 + row, ic0 + col).
// (With the interleaved 32-byte entries of the other variants every LDS.128 of a quarter-warp could reach only the
// even 16-byte bank groups: 2.5x the minimum number of shared-memory wavefronts, measured with ncu.)
struct __align__(128) PairStage {
    float4 p[kPairPlane];
    float4 q[kPairPlane];
    PairHeader hdr;
};
static_assert(offsetof(PairStage, q) == kPairPlane * 16, "planes must be contiguous: one 3-D TMA box fills both");

struct PairSmem {
    PairStage stage[kPairStages];
    float vx[kPairMaxRays], vy[kPairMaxRays], vz[kPairMaxRays];
    LinePair2 xpair[kPairMaxLines], ypair[kPairMaxLines];
    unsigned long long full_bar[kPairFullBars];
    unsigned long long empty_bar[kPairStages];
};
// Phase aliasing: a consumer warp visits only every 3rd..4th environment, so it may reach environment E while the
// TMA load of E - 8 (same stage) is still in flight (loads complete out of order); with one `full` barrier per stage
// its parity wait would then be satisfied by the phase of E - 16.  Environment E therefore signals on barrier E % 16
// with parity (E / 16) & 1: the warp's previous chunk (environment >= E - 4) was issued after E - 8, which was issued
// after E - 16 had been consumed (the producer issues strictly in order), so the barrier is never more than one phase
// behind the waiter.
static_assert(kPairFullBars == 2 * kPairStages, "full barriers: two per stage");

static_assert(sizeof(PairSmem) <= 227 * 1024, "PairSmem exceeds the shared memory of one SM");

// one 3-D TMA tile load (UTMALDG): box (set in the tensor map) = kPairPitch cells x kPairWin rows x 2 planes of the
// planar table, starting at cell (col, row); lands as PairStage::p followed by PairStage::q
__device__ __forceinline__ void tma_load_window_planar(void* dst, const CUtensorMap* tmap, int col, int row,
                                                       unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(s_addr(dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(col * 4), "r"(row), "r"(0), "r"(s_addr(bar))
        : "memory");
}

// ---- packed fp32 pairs (one 64-bit register pair; lo = ray 2k, hi = ray 2k+1)
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 dup(float v) { return pk(v, v); }
__device__ __forceinline__ float lo_of(f32x2 v) {
    return __uint_as_float((unsigned)(v & 0xffffffffull));
}
__device__ __forceinline__ float hi_of(f32x2 v) {
    return __uint_as_float((unsigned)(v >> 32));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// a * b rounded once, as a product that ptxas cannot contract with a following add: ptxas (12.9) fuses
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false, which would drop the reference's intermediate
// rounding.  fma(a, b, -0.0) == fl(a * b) for every input (sign of zero included); `nz` holds -0.0 behind a value the
// compiler cannot fold, so the instruction stays an FFMA2 and the following FADD2 stays an add.
__device__ __forceinline__ f32x2 mul2_unfused(f32x2 a, f32x2 b, f32x2 nz) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}
__device__ __forceinline__ f32x2 add2_rm(f32x2 a, f32x2 b) {  // round toward -inf
    f32x2 r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// per-chunk constants of the consumer loop
struct PairCtx {
    f32x2 CW, SZ, NSZ, S2, NS2, PX, PY, PZ, NWX0, NWY0, IDX, IDY, MAGIC, BASE, NEG0, ZFLAT;
    float pz, max_d;
    uint32_t xoff, yoff, eoff;  // byte offsets into shared memory of the window's line pairs / p plane
    uint32_t cmax, rmax;        // last column / row of the window
};

// Rare path (a): a ray whose cell guess missed, that lies on the closed far border / outside the grid, or sits in a
// general cell.  Out of line so that it does not raise the register pressure of the consumer loop.
__device__ __noinline__ void pair_resolve_deferred_ray(const PairSmem* sm, const PairStage* st, const ScanGridDev& g,
                                                       float inv_dx, float inv_dy, float X, float Y, float Z, float pz,
                                                       float max_d, float base_offset, float* __restrict__ out,
                                                       __nv_bfloat16* __restrict__ out_bf) {
    const PairHeader& h = st->hdr;
    const int cmax = h.ncols - 1, rmax = h.nrows - 1;
    const LinePair2* xp = sm->xpair + h.ic0;
    const LinePair2* yp = sm->ypair + h.jr0;
    const float wx0 = xp[0].lo, wy0 = yp[0].lo, wx1 = xp[cmax].hi, wy1 = yp[rmax].hi;
    float zhit = -INFINITY;
    if (X >= wx0 && X <= wx1 && Y >= wy0 && Y <= wy1) {
        int ci = min(max(__float2int_rd((X - wx0) * inv_dx), 0), cmax);
        int cj = min(max(__float2int_rd((Y - wy0) * inv_dy), 0), rmax);
        while (ci > 0 && X < xp[ci].lo) --ci;
        while (ci < cmax && X >= xp[ci].hi) ++ci;
        while (cj > 0 && Y < yp[cj].lo) --cj;
        while (cj < rmax && Y >= yp[cj].hi) ++cj;
        const int e = cj * kPairPitch + ci;
        const float4 q = st->q[e];
        zhit = (q.w == 0.f) ? eval_cell(st->p[e], q, __fsub_rn(X, xp[ci].lo), __fsub_rn(Y, yp[cj].lo), Z, max_d)
                            : walk_home_grid(g, X, Y, Z, max_d);
    }
    store_result(pz, X, Y, Z, zhit, base_offset, out, nullptr);
    if (out_bf != nullptr) *out_bf = __float2bfloat16_rn(*out);
}

// Rare path (b): a chunk of an environment whose window is not staged (too large, not covered on a non-uniform
// lattice, or more grid lines than the shared table holds).
__device__ __noinline__ void pair_resolve_chunk_from_global(const PairSmem* sm, int lane, int r_begin, int r_end,
                                                            const PairHeader h, const ScanGridDev& g,
                                                            const PlaneCellsDev& pc, float max_d, float base_offset,
                                                            float* __restrict__ out_row,
                                                            __nv_bfloat16* __restrict__ bf_row) {
    const float sz2 = __fmul_rn(h.sz, 2.f);
    for (int r = r_begin + lane; r < r_end; r += 32) {
        const int sl = slot_of_ray(r);
        const float vx = sm->vx[sl], vy = sm->vy[sl];
        const float tx = -__fmul_rn(sz2, vy), ty = __fmul_rn(sz2, vx);
        const float X = __fadd_rn(__fadd_rn(__fadd_rn(vx, __fmul_rn(h.cw, tx)), -__fmul_rn(h.sz, ty)), h.px);
        const float Y = __fadd_rn(__fadd_rn(__fadd_rn(vy, __fmul_rn(h.cw, ty)), __fmul_rn(h.sz, tx)), h.py);
        const float Z = __fadd_rn(sm->vz[sl], h.pz);
        store_result(h.pz, X, Y, Z, resolve_from_global(g, pc, X, Y, Z, max_d), base_offset, out_row + r, nullptr);
        if (bf_row != nullptr) bf_row[r] = __float2bfloat16_rn(out_row[r]);
    }
}

// One pair slot: rays (r, r+1).  Everything that can go wrong is deferred, so the body is branch-free up to the stores.
// Returns a 2-bit mask of the rays that need the rare path.
// kFull: the whole batch lies inside the pattern (no per-ray validity predicates); kFlatZ: every ray starts at the
// same local z (grid patterns), so Z is a per-environment constant.
// kBf: the heights are also stored as bf16 (round to nearest even) at ob[0] / ob[64] -- the observation mirror that
// feeds rover_policy_forward_bf16.
template <bool kFull, bool kFlatZ, bool kBf>
__device__ __forceinline__ unsigned resolve_pair(const unsigned char* __restrict__ smem, const PairSmem& sm,
                                                 const PairCtx& c, int slot, int ray_a, int n_rays,
                                                 float* __restrict__ o, __nv_bfloat16* __restrict__ ob) {
    const bool valid0 = kFull || ray_a < n_rays, valid1 = kFull || ray_a + 64 < n_rays;
    const int idx = slot;  // slots beyond the pattern hold zeros
    const f32x2 VX = *reinterpret_cast<const f32x2*>(sm.vx + idx);
    const f32x2 VY = *reinterpret_cast<const f32x2*>(sm.vy + idx);
    // ORBIT quat_apply_yaw + pos, same roundings as ray_origin(): tx = -(2sz * vy), ty = 2sz * vx,
    // X = ((vx + cw*tx) + (-(sz*ty))) + px, Y = ((vy + cw*ty) + sz*tx) + py, Z = vz + pz.  Negations are exact.
    const f32x2 TX = mul2(c.NS2, VY), TY = mul2(c.S2, VX);
    const f32x2 X = add2(add2(add2(VX, mul2_unfused(c.CW, TX, c.NEG0)), mul2_unfused(c.NSZ, TY, c.NEG0)), c.PX);
    const f32x2 Y = add2(add2(add2(VY, mul2_unfused(c.CW, TY, c.NEG0)), mul2_unfused(c.SZ, TX, c.NEG0)), c.PY);
    const f32x2 Z = kFlatZ ? c.ZFLAT : add2(*reinterpret_cast<const f32x2*>(sm.vz + idx), c.PZ);
    // window-relative cell guess: floor((X - wx0) * inv_dx) left in the mantissa, clamped on the raw bits
    const f32x2 BX = add2_rm(mul2(add2(X, c.NWX0), c.IDX), c.MAGIC);
    const f32x2 BY = add2_rm(mul2(add2(Y, c.NWY0), c.IDY), c.MAGIC);
    float nz[2];
    unsigned defer = 0;
    bool keep[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float x = k ? hi_of(X) : lo_of(X), y = k ? hi_of(Y) : lo_of(Y);
        // one unsigned min clamps both ends: bits below the bias (negative index, NaN ...) wrap to a huge value
        const uint32_t bi = min(__float_as_uint(k ? hi_of(BX) : lo_of(BX)) - (uint32_t)kFloorMagicBits, c.cmax);
        const uint32_t bj = min(__float_as_uint(k ? hi_of(BY) : lo_of(BY)) - (uint32_t)kFloorMagicBits, c.rmax);
        const float2 xp = *reinterpret_cast<const float2*>(smem + (c.xoff + bi * 8u));
        const float2 yp = *reinterpret_cast<const float2*>(smem + (c.yoff + bj * 8u));
#if ROVER_SCAN_DBG == 5  // timing experiment: conflict-free entry loads (results are wrong)
        const uint32_t e = c.eoff + ((threadIdx.x & 31) * 16u) + ((bj + bi) & 1u) * 1024u;
        const float4 q = *reinterpret_cast<const float4*>(smem + e + 512);
        const float4 p = *reinterpret_cast<const float4*>(smem + e);
        const bool fast = (x >= xp.x) & (x < xp.y) & (y >= yp.x) & (y < yp.y);
#else
        const uint32_t e = c.eoff + bj * (uint32_t)(kPairPitch * 16) + bi * 16u;
        const float4 q = *reinterpret_cast<const float4*>(smem + e + kPairPlane * 16);
        const float4 p = *reinterpret_cast<const float4*>(smem + e);
        const bool fast = (x >= xp.x) & (x < xp.y) & (y >= yp.x) & (y < yp.y) & (q.w == 0.f);
#endif
        const float lx = __fsub_rn(x, xp.x), ly = __fsub_rn(y, yp.x);
        const float E = fmaf(q.x, lx, fmaf(q.y, ly, q.z));
        // -z, with z = fma(k, min(E, 0), fma(a, lx, fma(b, ly, c))) exactly as eval_cell (negation commutes with rn)
        nz[k] = fmaf(-p.w, fminf(E, 0.f), -fmaf(p.x, lx, fmaf(p.y, ly, p.z)));
        const bool valid = k ? valid1 : valid0;
        keep[k] = valid && fast;
        defer |= (valid && !fast) ? (1u << k) : 0u;
    }
    // reference rounding chain: t = Z - z, hit.z = Z - t, height = (pos.z - hit.z) - offset; miss -> -inf
    const f32x2 T = add2(Z, pk(nz[0], nz[1]));
    const f32x2 H = sub2(sub2(c.PZ, sub2(Z, T)), c.BASE);
    const float t0 = lo_of(T), t1 = hi_of(T);
    const float h0 = (t0 >= 0.f && t0 < c.max_d) ? lo_of(H) : -INFINITY;
    const float h1 = (t1 >= 0.f && t1 < c.max_d) ? hi_of(H) : -INFINITY;
    if (keep[0]) o[0] = h0;
    if (keep[1]) o[64] = h1;
    if (kBf) {
        if (keep[0]) ob[0] = __float2bfloat16_rn(h0);
        if (keep[1]) ob[64] = __float2bfloat16_rn(h1);
    }
    return defer;
}

// ---- producer side: one environment per lane
struct ProducerEnv {
    float pv[3], qv[4];  // raw pose
    float cw, sz, px, py, pz;
    int ic0, jr0, ncols, nrows;
    bool have, ok;
};

__device__ __forceinline__ void producer_load(ProducerEnv& e, int it, int n_iter, const float* __restrict__ pos_w,
                                              const float* __restrict__ quat_w) {
    e.have = it < n_iter;
    const size_t env = (size_t)blockIdx.x + (size_t)(e.have ? it : 0) * gridDim.x;
#pragma unroll
    for (int k = 0; k < 3; ++k) e.pv[k] = __ldg(pos_w + 3 * env + k);
#pragma unroll
    for (int k = 0; k < 4; ++k) e.qv[k] = __ldg(quat_w + 4 * env + k);
    e.px = e.pv[0], e.py = e.pv[1], e.pz = e.pv[2];
}

// window from the position alone: every ray origin lies within pattern_radius of (px, py)
__device__ __forceinline__ void producer_window(ProducerEnv& e, const PlaneCellsDev& pc, float radius) {
    const float gx_lo = __ldg(pc.xs), gy_lo = __ldg(pc.ys);
    const int ic1 = min(max(guess_col(e.px + radius, gx_lo, pc.inv_dx) + 1, 0), pc.nx - 1);
    const int jr1 = min(max(guess_col(e.py + radius, gy_lo, pc.inv_dy) + 1, 0), pc.ny - 1);
    e.ic0 = min(max(guess_col(e.px - radius, gx_lo, pc.inv_dx) - 1, 0), pc.nx - 1);
    e.jr0 = min(max(guess_col(e.py - radius, gy_lo, pc.inv_dy) - 1, 0), pc.ny - 1);
    e.ncols = ic1 - e.ic0 + 1;
    e.nrows = jr1 - e.jr0 + 1;
}

__device__ __forceinline__ void producer_frame(ProducerEnv& e) {
    const SensorFrame f = make_frame(e.pv, e.qv);
    e.cw = f.cw, e.sz = f.sz;
}

// does the window fit the stage and (for lattices where the arithmetic guess is not exact +-1) cover the pattern?
__device__ __forceinline__ void producer_verdict(ProducerEnv& e, const PairSmem& sm, const PlaneCellsDev& pc,
                                                 float radius, bool lines_in_smem) {
    bool ok = lines_in_smem && (e.ncols <= kPairWin) && (e.nrows <= kPairWin);
    if (ok) {
        const float gx_lo = sm.xpair[0].lo, gx_hi = sm.xpair[pc.nx - 1].hi;
        const float gy_lo = sm.ypair[0].lo, gy_hi = sm.ypair[pc.ny - 1].hi;
        ok = (sm.xpair[e.ic0].lo <= fmaxf(e.px - radius, gx_lo)) &&
             (sm.xpair[e.ic0 + e.ncols - 1].hi >= fminf(e.px + radius, gx_hi)) &&
             (sm.ypair[e.jr0].lo <= fmaxf(e.py - radius, gy_lo)) &&
             (sm.ypair[e.jr0 + e.nrows - 1].hi >= fminf(e.py + radius, gy_hi));
    }
    e.ok = ok;
}

// kBf16: rover_height_scan_obs -- `out` points at column head_cols of the fp32 observation rows; the kernel also
// writes the bf16 mirror obs_bf16[env, 0 : head_cols + n_rays] (head columns converted from the fp32 row, heights
// rounded from the values it stores).
template <bool kBf16>
__global__ void __launch_bounds__(kPairThreads, 1)
height_scan_paired_kernel(const float* __restrict__ pos_w, const float* __restrict__ quat_w, int n_envs,
                          const float* __restrict__ ray_local, int n_rays, const __grid_constant__ ScanGridDev g,
                          const __grid_constant__ PlaneCellsDev pc, const __grid_constant__ CUtensorMap tmap,
                          float pattern_radius, float max_d, float base_offset, float* __restrict__ out,
                          int out_stride, __nv_bfloat16* __restrict__ obs_bf16, int bf16_stride, int head_cols) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PairSmem& sm = *reinterpret_cast<PairSmem*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_iter = (n_envs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // envs of this CTA
    const int n_chunks = (n_rays + kPairChunk - 1) / kPairChunk;                          // chunks per env
    const bool lines_in_smem = (pc.nx <= kPairMaxLines) && (pc.ny <= kPairMaxLines);
#if ROVER_SCAN_DBG == 3
    const long long dbg_t0 = clock64();
    if (threadIdx.x == 0 && blockIdx.x < 256) {
        g_scan_cta[0][blockIdx.x] = dbg_globaltimer();
        g_scan_cta[1][blockIdx.x] = 0;
    }
#endif

    // ---- prologue.  Producer warp: barriers, then -- lane-parallel, one environment per lane -- positions, table
    // windows and the first ring pass of TMA loads (a window needs the position only), then the sensor frames.
    // Consumer warps: pattern + line tables, every global load issued before the first shared store.
    ProducerEnv cur, nxt;
    int my_flat_z = 1;
    if (warp == 0) {
        producer_load(cur, lane, n_iter, pos_w, quat_w);  // cold misses: in flight while the barriers are set up
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
            for (int s = 0; s < kPairFullBars; ++s) bar_init(&sm.full_bar[s], 2);  // TMA bytes + header published
            for (int s = 0; s < kPairStages; ++s) bar_init(&sm.empty_bar[s], (uint32_t)n_chunks);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        producer_window(cur, pc, pattern_radius);
        if (lane < n_iter && lane < kPairEarly && ROVER_SCAN_DBG != 4) {
            // the box is always kPairPitch x kPairWin cells; cells beyond the table are zero-filled and never read;
            // a window that later turns out not to fit / not to cover is loaded all the same and simply not used
            bar_arrive_expect_tx(&sm.full_bar[lane], kPairStageBytes);
            tma_load_window_planar(sm.stage[lane].p, &tmap, cur.ic0, cur.jr0, &sm.full_bar[lane]);
            DBG_STAMP_ANY(16 + lane);
        }
        producer_frame(cur);
        DBG_STAMP(1);
    } else {
        constexpr int kFill = 32 * kPairConsumerWarps;
        constexpr int kPatLoads = (3 * kPairMaxRays + kFill - 1) / kFill;
        constexpr int kLineLoads = (kPairMaxLines + kFill - 1) / kFill;
        const int ct = threadIdx.x - 32;
        static_assert(offsetof(PairSmem, vy) == offsetof(PairSmem, vx) + 4 * kPairMaxRays &&
                          offsetof(PairSmem, vz) == offsetof(PairSmem, vx) + 8 * kPairMaxRays,
                      "vx, vy, vz are filled as one array");
        float* pat_flat = sm.vx;
        float pat[kPatLoads], xl[kLineLoads], xh[kLineLoads], yl[kLineLoads], yh[kLineLoads];
        const float vz0 = __ldg(ray_local + 2);
#pragma unroll
        for (int k = 0; k < kPatLoads; ++k) {  // i = component * kPairMaxRays + slot (vx, vy, vz are contiguous)
            const int i = ct + k * kFill;
            const int comp = i / kPairMaxRays, r = ray_of_slot(i - comp * kPairMaxRays);
            pat[k] = (i < 3 * kPairMaxRays && r < n_rays) ? __ldg(ray_local + 3 * r + comp) : 0.f;
            if (comp == 2 && r < n_rays && pat[k] != vz0) my_flat_z = 0;
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
            if (i < 3 * kPairMaxRays) pat_flat[i] = pat[k];
        }
#pragma unroll
        for (int k = 0; k < kLineLoads; ++k) {
            const int i = ct + k * kFill;
            if (lines_in_smem && i < pc.nx) sm.xpair[i] = {xl[k], xh[k]};
            if (lines_in_smem && i < pc.ny) sm.ypair[i] = {yl[k], yh[k]};
        }
    }
    if (warp == 1) DBG_STAMP(3);
    const bool flat_z = __syncthreads_and(my_flat_z) != 0;  // also publishes the tables / barriers to every warp
    if (warp == 0) DBG_STAMP(0);

    if (warp == 0) {
        // =============================== producer ===============================
        // One lane owns one environment of the current batch of 32: it already holds frame, window and verdict, so the
        // per-environment critical path is wait(empty) -> header -> TMA issue.  (A serial loop that derived the window
        // per iteration took ~850 cycles alone and ~2000 beside busy consumers -- it, not the TMA, starved them.)
        for (int base = 0; base < n_iter; base += 32) {
            if (base > 0) cur = nxt;
            producer_verdict(cur, sm, pc, pattern_radius, lines_in_smem);
            const int n_here = min(32, n_iter - base);
            for (int k = 0; k < n_here; ++k) {
                if (k == kPairStages && base + 32 < n_iter) {  // next batch, computed while the ring is full
                    producer_load(nxt, base + 32 + lane, n_iter, pos_w, quat_w);
                    producer_window(nxt, pc, pattern_radius);
                    producer_frame(nxt);
                }
                if (lane == k) {
                    const int it = base + k;
                    const int s = it % kPairStages;
                    unsigned long long* full = &sm.full_bar[it % kPairFullBars];
                    PairStage& st = sm.stage[s];
                    const bool first_pass = it < kPairEarly && ROVER_SCAN_DBG != 4;  // loaded in the prologue
                    if (!first_pass) {
                        bar_wait(&sm.empty_bar[s], ((uint32_t)(it / kPairStages) & 1u) ^ 1u);  // stage drained
                        if (cur.ok && ROVER_SCAN_DBG != 2) {
                            bar_arrive_expect_tx(full, kPairStageBytes);
                            tma_load_window_planar(st.p, &tmap, cur.ic0, cur.jr0, full);
                        } else {
                            bar_arrive(full);
                        }
                        DBG_STAMP_ANY(16 + it);
                    }
                    st.hdr = {cur.cw, cur.sz, cur.px, cur.py, cur.pz, cur.ic0, cur.jr0, cur.ncols, cur.nrows,
                              cur.ok ? 1 : 0, 0, 0};
                    bar_arrive(full);  // header published (release)
                }
                // environments must be issued IN ORDER (the phase-aliasing argument above relies on it): without this
                // the lanes run ahead independently and a later environment whose stage drains first overtakes
                __syncwarp();
            }
        }
        DBG_STAMP(2);
    } else {
        // =============================== consumers ===============================
        const int w = warp - 1;
        const int step_it = kPairConsumerWarps / n_chunks, step_c = kPairConsumerWarps % n_chunks;
        int it = w / n_chunks, c = w % n_chunks;
#if ROVER_SCAN_DBG == 3
        int dbg_j = 0;
#endif
        while (it < n_iter) {
            const int env = (int)blockIdx.x + it * (int)gridDim.x;
            const int s = it % kPairStages;
            const PairStage& st = sm.stage[s];
#if ROVER_SCAN_DBG == 3
            const int dbg_slot = (w == 0 ? 64 : w == 13 ? 128 : w == 6 ? 192 : 100000) + 4 * dbg_j;
            ++dbg_j;
#endif
            DBG_STAMP(dbg_slot);
            bar_wait(&sm.full_bar[it % kPairFullBars], (uint32_t)(it / kPairFullBars) & 1u);
            DBG_STAMP(dbg_slot + 1);
            const PairHeader h = st.hdr;
            float* __restrict__ out_row = out + (size_t)env * out_stride;
            __nv_bfloat16* __restrict__ bf_row = kBf16 ? obs_bf16 + (size_t)env * bf16_stride + head_cols : nullptr;
            const int r_begin = c * kPairChunk, r_end = min(r_begin + kPairChunk, n_rays);
            if (kBf16 && c == 0 && lane < head_cols)  // the head of the observation (written by the post-step kernel)
                bf_row[lane - head_cols] = __float2bfloat16_rn(out_row[lane - head_cols]);
            if (ROVER_SCAN_DBG == 1) {
            } else if (h.mode == 1) {
                PairCtx cx;
                const float sz2 = __fmul_rn(h.sz, 2.f);  // fl(fl(sz*v)*2) == fl(fl(2*sz)*v): scaling by 2 is exact
                const float wx0 = sm.xpair[h.ic0].lo, wy0 = sm.ypair[h.jr0].lo;
                cx.CW = dup(h.cw), cx.SZ = dup(h.sz), cx.NSZ = dup(-h.sz), cx.S2 = dup(sz2), cx.NS2 = dup(-sz2);
                cx.PX = dup(h.px), cx.PY = dup(h.py), cx.PZ = dup(h.pz);
                cx.NWX0 = dup(-wx0), cx.NWY0 = dup(-wy0), cx.IDX = dup(pc.inv_dx), cx.IDY = dup(pc.inv_dy);
                cx.MAGIC = dup(kFloorMagic), cx.BASE = dup(base_offset);
                cx.NEG0 = dup(__uint_as_float(0x80000000u | (unsigned)(n_envs >> 31)));  // -0.0, opaque to the compiler
                cx.pz = h.pz, cx.max_d = max_d;
                cx.ZFLAT = dup(__fadd_rn(sm.vz[0], h.pz));
                cx.cmax = (uint32_t)(h.ncols - 1), cx.rmax = (uint32_t)(h.nrows - 1);
                cx.xoff = (uint32_t)(reinterpret_cast<const unsigned char*>(sm.xpair + h.ic0) - smem_raw);
                cx.yoff = (uint32_t)(reinterpret_cast<const unsigned char*>(sm.ypair + h.jr0) - smem_raw);
                cx.eoff = (uint32_t)(reinterpret_cast<const unsigned char*>(st.p) - smem_raw);
                for (int b0 = r_begin; b0 < r_end; b0 += kPairBatch) {
                    const int r = b0 + lane;  // rays r, r + 64 (slot 0) and r + 32, r + 96 (slot 1)
                    float* __restrict__ o = out_row + r;
                    __nv_bfloat16* __restrict__ ob = kBf16 ? bf_row + r : nullptr;
                    unsigned defer;
                    if (flat_z && b0 + kPairBatch <= n_rays) {
                        defer = resolve_pair<true, true, kBf16>(smem_raw, sm, cx, b0 + 2 * lane, r, n_rays, o, ob);
                        defer |= resolve_pair<true, true, kBf16>(smem_raw, sm, cx, b0 + 64 + 2 * lane, r + 32, n_rays, o + 32,
                                                                 ob + 32) << 2;
                    } else {
                        defer = resolve_pair<false, false, kBf16>(smem_raw, sm, cx, b0 + 2 * lane, r, n_rays, o, ob);
                        defer |= resolve_pair<false, false, kBf16>(smem_raw, sm, cx, b0 + 64 + 2 * lane, r + 32, n_rays, o + 32,
                                                                   ob + 32) << 2;
                    }
                    // rare: cell guess off by one, ray on the closed far border or outside the grid, general cell
                    if (defer != 0u) {
                        for (int u = 0; u < 4; ++u) {
                            if (!((defer >> u) & 1u)) continue;
                            const int rr = r + (u >> 1) * 32 + (u & 1) * 64;
                            const int sl = b0 + (u >> 1) * 64 + 2 * lane + (u & 1);
                            const float vx = sm.vx[sl], vy = sm.vy[sl];
                            const float tx = -__fmul_rn(sz2, vy), ty = __fmul_rn(sz2, vx);
                            const float X = __fadd_rn(__fadd_rn(__fadd_rn(vx, __fmul_rn(h.cw, tx)), -__fmul_rn(h.sz, ty)), h.px);
                            const float Y = __fadd_rn(__fadd_rn(__fadd_rn(vy, __fmul_rn(h.cw, ty)), __fmul_rn(h.sz, tx)), h.py);
                            pair_resolve_deferred_ray(&sm, &st, g, pc.inv_dx, pc.inv_dy, X, Y, __fadd_rn(sm.vz[sl], h.pz),
                                                      h.pz, max_d, base_offset, out_row + rr,
                                                      kBf16 ? bf_row + rr : nullptr);
                        }
                    }
                }
            } else {
                pair_resolve_chunk_from_global(&sm, lane, r_begin, r_end, h, g, pc, max_d, base_offset, out_row, bf_row);
            }
            __syncwarp();
            DBG_STAMP(dbg_slot + 2);
            if (lane == 0) bar_arrive(&sm.empty_bar[s]);  // this chunk no longer reads the stage
            it += step_it;
            c += step_c;
            if (c >= n_chunks) {
                c -= n_chunks;
                ++it;
            }
        }
    }
#if ROVER_SCAN_DBG == 3
    if (lane == 0 && blockIdx.x < 256) atomicMax(&g_scan_cta[1][blockIdx.x], dbg_globaltimer());
#endif
}

int launch_height_scan_pipelined(const float* pos_w, const float* quat_w, int n_envs, const float* ray_local,
                                 int n_rays, const ScanGridDev& g, const RoverPlaneCells* cells, float4 pattern_box,
                                 float max_d, float base_offset, float* out, int out_stride, float* hits,
                                 cudaStream_t stream);

// obs_bf16 != nullptr: `out` is column head_cols of the fp32 observation rows and the bf16 mirror is written as well
// (rover_height_scan_obs); the caller has checked that variant 5 can run (planar table, pattern size, no hit output).
int launch_height_scan_paired(const float* pos_w, const float* quat_w, int n_envs, const float* ray_local, int n_rays,
                              const ScanGridDev& g, const RoverPlaneCells* cells, float4 pattern_box, float max_d,
                              float base_offset, float* out, int out_stride, float* hits, uint16_t* obs_bf16,
                              int bf16_stride, int head_cols, cudaStream_t stream) {
    // hit positions are a debugging / test output: served by variant 4's kernel (same heights, same table); so is a
    // table without the planar copy
    if (obs_bf16 == nullptr && (hits != nullptr || cells->entries_planar == nullptr))
        return launch_height_scan_pipelined(pos_w, quat_w, n_envs, ray_local, n_rays, g, cells, pattern_box, max_d,
                                            base_offset, out, out_stride, hits, stream);
    ROVER_CHECK(n_rays >= 1 && n_rays <= kPairMaxRays, "height_scan_paired: pattern of %d rays (1..%d supported)", n_rays,
                kPairMaxRays);
    ROVER_CHECK(cells->entries_planar != nullptr && hits == nullptr, "height_scan_paired: bf16 mirror needs the planar table");
    ROVER_CHECK(head_cols >= 0 && head_cols <= 32, "height_scan_paired: head_cols %d out of range", head_cols);
    static int n_sms = 0;
    static bool configured = false;
    if (!configured) {
        int dev = 0;
        ROVER_CUDA(cudaGetDevice(&dev));
        ROVER_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
        ROVER_CUDA(cudaFuncSetAttribute(height_scan_paired_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(PairSmem)));
        ROVER_CUDA(cudaFuncSetAttribute(height_scan_paired_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(PairSmem)));
        configured = true;
    }
    alignas(64) CUtensorMap tmap;
    if (const int rc = encode_planar_tensor_map(&tmap, cells, kPairPitch, kPairWin, "height_scan_paired")) return rc;
    PlaneCellsDev pc{cells->xs, cells->ys, reinterpret_cast<const float4*>(cells->entries), cells->nx, cells->ny,
                     cells->inv_dx, cells->inv_dy};
    // every ray origin lies within this distance of the sensor position (the yaw rotation preserves norms)
    const float rx = fmaxf(fabsf(pattern_box.x), fabsf(pattern_box.y)), ry = fmaxf(fabsf(pattern_box.z), fabsf(pattern_box.w));
    const float radius = sqrtf(rx * rx + ry * ry) * 1.0001f + 1.0e-3f;
    const int grid = n_envs < n_sms ? n_envs : n_sms;
    if (obs_bf16 != nullptr)
        height_scan_paired_kernel<true><<<grid, kPairThreads, sizeof(PairSmem), stream>>>(
            pos_w, quat_w, n_envs, ray_local, n_rays, g, pc, tmap, radius, max_d, base_offset, out, out_stride,
            reinterpret_cast<__nv_bfloat16*>(obs_bf16), bf16_stride, head_cols);
    else
        height_scan_paired_kernel<false><<<grid, kPairThreads, sizeof(PairSmem), stream>>>(
            pos_w, quat_w, n_envs, ray_local, n_rays, g, pc, tmap, radius, max_d, base_offset, out, out_stride, nullptr, 0, 0);
    return check_launch("height_scan_paired_kernel");
}

}  // namespace rover

#if ROVER_SCAN_DBG == 3
extern "C" int rover_debug_scan_timeline(unsigned long long* host_dst) {
    return (int)cudaMemcpyFromSymbol(host_dst, rover::g_scan_dbg, sizeof(rover::g_scan_dbg));
}
extern "C" int rover_debug_scan_ctas(unsigned long long* host_dst) {
    return (int)cudaMemcpyFromSymbol(host_dst, rover::g_scan_cta, sizeof(rover::g_scan_cta));
}
#endif
