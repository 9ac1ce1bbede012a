// Init-time terrain tables on the GPU: the mesh -> heightmap rasterisation and the steep-cell mask
// that seeds the rock detection.  One thread per face / per cell; results are bit-identical to the reference's
// sequential loops because max() is order-independent and the gradient is evaluated in the reference's precision.
#include "common.cuh"

namespace rover {

// max() on a float cell: sign-aware integer atomics keep -0.0 / +0.0 and negative heights exact
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
    if (!signbit(v))
        atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// terrain_utils.py:23-57.  Per face: the cells covered by the face's XY bounding box take max(cell, max z of face).
// Cell indices are int()-truncated (toward zero), only the upper index is clamped, negative indices wrap once like
// Python's; a face beyond that raises IndexError in the reference and is counted in `out_of_range` here.
__global__ void __launch_bounds__(256)
mesh_to_heightmap_kernel(const float* __restrict__ vertices, const int* __restrict__ faces, int n_faces, float min_x,
                         float min_y, float cell_x, float cell_y, int rows, int cols, float* __restrict__ heightmap,
                         int* __restrict__ out_of_range) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_faces) return;
    const int a = __ldg(faces + 3 * (size_t)f), b = __ldg(faces + 3 * (size_t)f + 1), c = __ldg(faces + 3 * (size_t)f + 2);
    const float ax = __ldg(vertices + 3 * (size_t)a), ay = __ldg(vertices + 3 * (size_t)a + 1), az = __ldg(vertices + 3 * (size_t)a + 2);
    const float bx = __ldg(vertices + 3 * (size_t)b), by = __ldg(vertices + 3 * (size_t)b + 1), bz = __ldg(vertices + 3 * (size_t)b + 2);
    const float cx = __ldg(vertices + 3 * (size_t)c), cy = __ldg(vertices + 3 * (size_t)c + 1), cz = __ldg(vertices + 3 * (size_t)c + 2);
    const float lo_x = fminf(ax, fminf(bx, cx)), hi_x = fmaxf(ax, fmaxf(bx, cx));
    const float lo_y = fminf(ay, fminf(by, cy)), hi_y = fmaxf(ay, fmaxf(by, cy));
    const float zmax = fmaxf(az, fmaxf(bz, cz));
    // int((p - min) / cell): fp32 subtract and divide, truncation toward zero (:43-46)
    const long long min_i = (long long)__fdiv_rn(__fsub_rn(lo_x, min_x), cell_x);
    const long long max_i = min((long long)__fdiv_rn(__fsub_rn(hi_x, min_x), cell_x), (long long)cols - 1);
    const long long min_j = (long long)__fdiv_rn(__fsub_rn(lo_y, min_y), cell_y);
    const long long max_j = min((long long)__fdiv_rn(__fsub_rn(hi_y, min_y), cell_y), (long long)rows - 1);
    if (max_i < min_i || max_j < min_j) return;
    if (min_i < -(long long)cols || min_j < -(long long)rows) {  // heightmap[j, i] with j or i below -size
        atomicAdd(out_of_range, 1);
        return;
    }
    for (long long j = min_j; j <= max_j; ++j) {
        const long long jj = j < 0 ? j + rows : j;
        for (long long i = min_i; i <= max_i; ++i) {
            const long long ii = i < 0 ? i + cols : i;
            atomic_max_f32(heightmap + jj * cols + ii, zmax);
        }
    }
}

// terrain_utils.py:265-279: Sobel gradients with wrap-around borders (scipy convolve2d, boundary="wrap"; the kernel
// is flipped by the convolution), magnitude > threshold.  fp32 heights, float64 arithmetic as numpy promotes it.
__global__ void __launch_bounds__(256)
steep_mask_kernel(const float* __restrict__ heightmap, int rows, int cols, double threshold, uint8_t* __restrict__ steep) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int xm = x == 0 ? cols - 1 : x - 1, xp = x == cols - 1 ? 0 : x + 1;
    const int ym = y == 0 ? rows - 1 : y - 1, yp = y == rows - 1 ? 0 : y + 1;
    auto h = [&](int r, int c) { return (double)__ldg(heightmap + (size_t)r * cols + c); };
    const double a00 = h(ym, xm), a01 = h(ym, x), a02 = h(ym, xp);
    const double a10 = h(y, xm), a12 = h(y, xp);
    const double a20 = h(yp, xm), a21 = h(yp, x), a22 = h(yp, xp);
    // convolution with [[-1,0,1],[-2,0,2],[-1,0,1]]: out[y,x] = sum k[i,j] * in[y+1-i, x+1-j]; small-integer products
    // of fp32 values are exact in float64 and so is the six-term sum for terrain-scale heights
    const double gx = __dadd_rn(__dadd_rn(__dsub_rn(a00, a02), __dmul_rn(2.0, __dsub_rn(a10, a12))), __dsub_rn(a20, a22));
    const double gy = __dadd_rn(__dadd_rn(__dsub_rn(a00, a20), __dmul_rn(2.0, __dsub_rn(a01, a21))), __dsub_rn(a02, a22));
    const double mag = sqrt(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)));
    steep[(size_t)y * cols + x] = mag > threshold ? 1 : 0;
}

// terrain_utils.py:281-311: the morphological clean-up of the steep mask -- OpenCV box morphology and scipy
// binary_fill_holes in the reference.  A k x k box is separable: one pass along x, one along y.  OpenCV semantics: the
// anchor sits at k / 2, so the window covers offsets [-(k / 2), k - 1 - k / 2] (asymmetric for the 42 x 42 safety margin);
// pixels outside the image never contribute (dilate) / never limit (erode).  Binary images: dilate = any, erode = all.
__global__ void __launch_bounds__(256)
morph_line_kernel(const uint8_t* __restrict__ src, int rows, int cols, int k, int vertical, int erode, uint8_t* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int lo = -(k / 2), hi = k - 1 - k / 2;
    int hit = erode ? 1 : 0;
    for (int o = lo; o <= hi; ++o) {
        const int xx = vertical ? x : x + o, yy = vertical ? y + o : y;
        if (xx < 0 || xx >= cols || yy < 0 || yy >= rows) continue;
        const int v = __ldg(src + (size_t)yy * cols + xx) != 0;
        hit = erode ? (hit & v) : (hit | v);
    }
    dst[(size_t)y * cols + x] = (uint8_t)hit;
}

// scipy.ndimage.binary_fill_holes: the complement of the background connected (4-neighbourhood) to the image border.
// reach = 1 on background pixels already known to connect to the border.  One launch relaxes 32 x 32 tiles to their
// local fixed point in shared memory (halo of one pixel) and raises *changed when a tile changed; the host repeats the
// launch until nothing changes (path lengths are counted in tiles, not pixels).
constexpr int kFillTile = 32;
__global__ void __launch_bounds__(kFillTile * kFillTile)
fill_reach_kernel(const uint8_t* __restrict__ mask, uint8_t* __restrict__ reach, int rows, int cols, int* __restrict__ changed) {
    __shared__ uint8_t r[kFillTile + 2][kFillTile + 2];
    __shared__ int tile_changed, any_changed;
    const int tx = threadIdx.x % kFillTile, ty = threadIdx.x / kFillTile;
    const int x = blockIdx.x * kFillTile + tx, y = blockIdx.y * kFillTile + ty;
    const bool in = x < cols && y < rows;
    const bool open = in && __ldg(mask + (size_t)y * cols + x) == 0;
    auto load = [&](int yy, int xx) -> uint8_t {
        return (xx >= 0 && xx < cols && yy >= 0 && yy < rows) ? reach[(size_t)yy * cols + xx] : (uint8_t)0;
    };
    r[ty + 1][tx + 1] = in ? reach[(size_t)y * cols + x] : (uint8_t)0;
    if (ty == 0) r[0][tx + 1] = load(y - 1, x);
    if (ty == kFillTile - 1) r[kFillTile + 1][tx + 1] = load(y + 1, x);
    if (tx == 0) r[ty + 1][0] = load(y, x - 1);
    if (tx == kFillTile - 1) r[ty + 1][kFillTile + 1] = load(y, x + 1);
    if (threadIdx.x == 0) any_changed = 0;
    __syncthreads();
    const uint8_t before = r[ty + 1][tx + 1];
    for (int it = 0; it < 2 * kFillTile * kFillTile; ++it) {
        if (threadIdx.x == 0) tile_changed = 0;
        __syncthreads();
        if (open && !r[ty + 1][tx + 1] && (r[ty][tx + 1] | r[ty + 2][tx + 1] | r[ty + 1][tx] | r[ty + 1][tx + 2])) {
            r[ty + 1][tx + 1] = 1;
            tile_changed = 1;
        }
        __syncthreads();
        if (!tile_changed) break;
        __syncthreads();
    }
    if (in && r[ty + 1][tx + 1] != before) {
        reach[(size_t)y * cols + x] = 1;
        any_changed = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0 && any_changed) *changed = 1;
}

// reach seed: background pixels on the image border; and the final composition filled = !reach
__global__ void fill_seed_kernel(const uint8_t* __restrict__ mask, uint8_t* __restrict__ reach, int rows, int cols) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * cols) return;
    const int y = (int)(i / cols), x = (int)(i % cols);
    const bool border = x == 0 || y == 0 || x == cols - 1 || y == rows - 1;
    reach[i] = (border && mask[i] == 0) ? 1 : 0;
}
__global__ void fill_finish_kernel(const uint8_t* __restrict__ reach, uint8_t* __restrict__ out, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = reach[i] ? 0 : 1;
}

}  // namespace rover

using namespace rover;

extern "C" int rover_mesh_to_heightmap(const float* vertices, const int32_t* faces, int32_t n_faces, float min_x,
                                       float min_y, float cell_x, float cell_y, int32_t rows, int32_t cols,
                                       float* heightmap, int32_t* out_of_range, void* stream) {
    ROVER_CHECK(vertices && faces && heightmap && out_of_range, "rover_mesh_to_heightmap: NULL pointer");
    ROVER_CHECK(n_faces >= 0 && rows > 0 && cols > 0, "rover_mesh_to_heightmap: bad sizes (%d faces, %d x %d)", n_faces, rows, cols);
    ROVER_CHECK(cell_x > 0.f && cell_y > 0.f, "rover_mesh_to_heightmap: cell size must be positive");
    if (n_faces == 0) return 0;
    mesh_to_heightmap_kernel<<<(n_faces + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        vertices, faces, n_faces, min_x, min_y, cell_x, cell_y, rows, cols, heightmap, out_of_range);
    return check_launch("mesh_to_heightmap_kernel");
}

extern "C" int rover_steep_mask(const float* heightmap, int32_t rows, int32_t cols, double threshold, uint8_t* steep,
                                void* stream) {
    ROVER_CHECK(heightmap && steep, "rover_steep_mask: NULL pointer");
    ROVER_CHECK(rows > 0 && cols > 0 && rows <= 65535, "rover_steep_mask: bad shape %d x %d", rows, cols);
    steep_mask_kernel<<<dim3((cols + 255) / 256, rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(heightmap, rows, cols,
                                                                                                  threshold, steep);
    return check_launch("steep_mask_kernel");
}

extern "C" int rover_morph_box(const uint8_t* src, int32_t rows, int32_t cols, int32_t k, int32_t erode, uint8_t* tmp,
                               uint8_t* dst, void* stream) {
    using namespace rover;
    ROVER_CHECK(src && tmp && dst && rows > 0 && cols > 0 && k >= 1, "rover_morph_box: bad arguments");
    ROVER_CHECK(src != dst && src != tmp && tmp != dst, "rover_morph_box: src, tmp and dst must be distinct buffers");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const dim3 grid((cols + 255) / 256, rows);
    morph_line_kernel<<<grid, 256, 0, s>>>(src, rows, cols, k, 0, erode, tmp);
    morph_line_kernel<<<grid, 256, 0, s>>>(tmp, rows, cols, k, 1, erode, dst);
    return check_launch("morph_line_kernel");
}

extern "C" int rover_fill_holes(const uint8_t* mask, int32_t rows, int32_t cols, uint8_t* reach /* scratch */,
                                int32_t* changed /* device scratch, 1 int */, uint8_t* out, void* stream) {
    using namespace rover;
    ROVER_CHECK(mask && reach && changed && out && rows > 0 && cols > 0, "rover_fill_holes: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n = (size_t)rows * cols;
    fill_seed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mask, reach, rows, cols);
    const dim3 grid((cols + kFillTile - 1) / kFillTile, (rows + kFillTile - 1) / kFillTile);
    // init-time code: the host polls the convergence flag every 4 relaxation launches (a tile relaxes to its local fixed
    // point per launch, so the launch count grows with the longest background path measured in tiles)
    for (int round = 0; round < 100000; ++round) {
        ROVER_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), s));
        for (int rep = 0; rep < 4; ++rep) fill_reach_kernel<<<grid, kFillTile * kFillTile, 0, s>>>(mask, reach, rows, cols, changed);
        int h = 0;
        ROVER_CUDA(cudaMemcpyAsync(&h, changed, sizeof(int), cudaMemcpyDeviceToHost, s));
        ROVER_CUDA(cudaStreamSynchronize(s));
        if (!h) break;
    }
    fill_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reach, out, n);
    return check_launch("fill_holes");
}
