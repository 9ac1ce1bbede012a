"""Terrain side of the hot path: synthetic Mars-like mesh + the init-time lookup tables.

Host-side (numpy / torch / cv2 / scipy) like the reference's own init code; these run once per
environment construction and feed the per-step CUDA kernels with immutable tables:

* ``mesh_to_heightmap``      -- reference ``HeightmapManager.mesh_to_heightmap``
                                (envs/navigation/utils/terrains/terrain_utils.py:23-57), vectorised;
* ``find_rocks_in_heightmap``-- reference ``TerrainManager.find_rocks_in_heightmap`` (:265-311);
* ``random_rover_spawns``    -- reference ``TerrainManager.random_rover_spawns`` (:330-385);
* ``TerrainTables``          -- what ``TerrainManager.__init__`` (:92-127) leaves behind:
                                heightmap, safe rock mask, spawn table, (min_x, min_y) offset.

The reference loads its mesh from USD (``/World/terrain/hidden_terrain``); the USD blobs are not
shipped, so ``make_synthetic_terrain`` builds the 200 m x 200 m / 2 M-triangle stand-in that
SURVEY.md section 8(d) specifies (regular 0.2 m grid, fBm + rock bumps, seeded).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

HEIGHTMAP_RESOLUTION = 0.05  # terrain_utils.py:108
GRADIENT_THRESHOLD = 0.3  # terrain_utils.py:109
BORDER_MARGIN = 1.0  # terrain_utils.py:25
SPAWN_BORDER_OFFSET = 20.0  # terrain_utils.py:334
SPAWN_SEED = 41  # terrain_utils.py:124


# ----------------------------------------------------------------------------------------
# synthetic terrain
# ----------------------------------------------------------------------------------------
def _value_noise(n: int, cells: int, rng: np.random.Generator) -> np.ndarray:
    """Smooth value noise on an n x n grid from a (cells+1)^2 random lattice (smoothstep blend)."""
    lat = rng.standard_normal((cells + 2, cells + 2)).astype(np.float64)
    u = np.linspace(0.0, cells, n)
    i = np.minimum(u.astype(np.int64), cells - 1)
    f = u - i
    f = f * f * (3.0 - 2.0 * f)
    a = lat[i][:, i]
    b = lat[i][:, i + 1]
    c = lat[i + 1][:, i]
    d = lat[i + 1][:, i + 1]
    fx = f[None, :]
    fy = f[:, None]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def make_synthetic_terrain(size_m: float = 200.0, grid_res: float = 0.2, seed: int = 0, n_rocks: int | None = None,
                           octaves: int = 4, amplitude: float = 1.0):
    """Regular-grid triangle mesh over ``[0, size_m]^2``.

    Returns ``(vertices [V,3] f32, faces [F,3] i32)``; ``V = (size/res+1)^2``, ``F = 2 (size/res)^2``.
    Height = ``octaves`` of value noise (wavelength 40 m halving, amplitude halving) + smooth rock bumps
    (radius 0.3-1.5 m, height 0.3-1.0 m) so that the rock mask is non-trivial.
    """
    rng = np.random.default_rng(seed)
    n = int(round(size_m / grid_res)) + 1
    z = np.zeros((n, n), dtype=np.float64)
    amp = amplitude
    wavelength = 40.0
    for _ in range(octaves):
        cells = max(int(round(size_m / wavelength)), 1)
        z += amp * _value_noise(n, cells, rng)
        amp *= 0.5
        wavelength *= 0.5
    if n_rocks is None:
        n_rocks = int(round(1200 * (size_m / 200.0) ** 2))
    xs = np.arange(n) * grid_res
    for _ in range(n_rocks):
        cx, cy = rng.uniform(0.0, size_m, size=2)
        rad = rng.uniform(0.3, 1.5)
        hgt = rng.uniform(0.3, 1.0)
        i0, i1 = np.searchsorted(xs, [cx - rad, cx + rad])
        j0, j1 = np.searchsorted(xs, [cy - rad, cy + rad])
        if i1 <= i0 or j1 <= j0:
            continue
        dx = xs[i0:i1][None, :] - cx
        dy = xs[j0:j1][:, None] - cy
        q = 1.0 - (dx * dx + dy * dy) / (rad * rad)
        z[j0:j1, i0:i1] += hgt * np.where(q > 0, np.sqrt(np.maximum(q, 0.0)), 0.0)
    gx, gy = np.meshgrid(xs, xs, indexing="xy")
    vertices = np.stack([gx.ravel(), gy.ravel(), z.ravel()], axis=1).astype(np.float32)
    m = n - 1
    idx = (np.arange(m)[:, None] * n + np.arange(m)[None, :]).ravel()
    lower = np.stack([idx, idx + 1, idx + n], axis=1)
    upper = np.stack([idx + 1, idx + n + 1, idx + n], axis=1)
    faces = np.empty((2 * m * m, 3), dtype=np.int32)
    faces[0::2] = lower
    faces[1::2] = upper
    return vertices, faces


# ----------------------------------------------------------------------------------------
# init-time tables
# ----------------------------------------------------------------------------------------
def mesh_to_heightmap(vertices: np.ndarray, faces: np.ndarray, resolution: float = HEIGHTMAP_RESOLUTION,
                      device: str | torch.device = "cpu"):
    """Vectorised restatement of terrain_utils.py:23-57.

    Per face: the heightmap cells covered by the face's XY bounding box take ``max(cell, max z of face)``.
    Quirks kept because they define the tables the per-step lookups read: bounds are shrunk by the 1 m
    border; the array is allocated ``(nx, ny)`` but indexed ``[j, i]``; cell indices are ``int()``-truncated
    (toward zero) and only the upper index is clamped, so faces in the lower border wrap to the far side
    through Python's negative indexing.

    ``device`` = a CUDA device: the rasterisation runs in ``rover_mesh_to_heightmap`` (one thread per face, atomic max;
    csrc/terrain_build.cu) -- bit-identical, no fallback if the library is missing.  ``device="cpu"``: the host
    builder (vectorised torch scatter), which the CPU tests pin against the reference's golden tables.

    Returns ``(heightmap f32 [nx, ny], min_x, min_y, max_x, max_y)``.
    """
    v = np.asarray(vertices, dtype=np.float32)
    f = np.asarray(faces).astype(np.int64)
    margin = np.float32(BORDER_MARGIN)
    res = np.float32(resolution)
    mn = v.min(axis=0) + margin
    mx = v.max(axis=0) - margin
    min_x, min_y, max_x, max_y = mn[0], mn[1], mx[0], mx[1]
    gsx = (max_x - min_x) / res
    gsy = (max_y - min_y) / res
    shape = (int(gsx + 1), int(gsy + 1))
    csx = (max_x - min_x) / gsx
    csy = (max_y - min_y) / gsy
    dev = torch.device(device)
    if dev.type == "cuda":
        hm = _mesh_to_heightmap_cuda(v, f, min_x, min_y, csx, csy, shape, dev)
        return hm.cpu().numpy(), float(min_x), float(min_y), float(max_x), float(max_y)
    tri = v[f]  # [F, 3, 3]
    lo = tri.min(axis=1)
    hi = tri.max(axis=1)
    zmax = hi[:, 2]
    min_i = np.trunc((lo[:, 0] - min_x) / csx).astype(np.int64)
    max_i = np.minimum(np.trunc((hi[:, 0] - min_x) / csx).astype(np.int64), shape[1] - 1)
    min_j = np.trunc((lo[:, 1] - min_y) / csy).astype(np.int64)
    max_j = np.minimum(np.trunc((hi[:, 1] - min_y) / csy).astype(np.int64), shape[0] - 1)
    di = max_i - min_i + 1
    dj = max_j - min_j + 1
    keep = (di > 0) & (dj > 0)
    hm = torch.full((shape[0] * shape[1],), -99.0, dtype=torch.float32, device=dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    tmin_i, tmin_j, tdi, tdj, tz, tkeep = t(min_i), t(min_j), t(di), t(dj), t(zmax), t(keep)
    for b in range(int(dj[keep].max()) if keep.any() else 0):
        for a in range(int(di[keep].max())):
            sel = tkeep & (tdi > a) & (tdj > b)
            if not bool(sel.any()):
                continue
            ii = tmin_i[sel] + a
            jj = tmin_j[sel] + b
            if bool(((ii < -shape[1]) | (jj < -shape[0]) | (jj >= shape[0])).any()):
                raise IndexError("face outside the heightmap beyond the wrap range (the reference raises here too)")
            ii = torch.where(ii < 0, ii + shape[1], ii)  # Python negative indexing
            jj = torch.where(jj < 0, jj + shape[0], jj)
            hm.scatter_reduce_(0, jj * shape[1] + ii, tz[sel], reduce="amax", include_self=True)
    return hm.view(shape).cpu().numpy(), float(min_x), float(min_y), float(max_x), float(max_y)


def _mesh_to_heightmap_cuda(v: np.ndarray, f: np.ndarray, min_x, min_y, csx, csy, shape, dev) -> torch.Tensor:
    """The rasterisation of ``mesh_to_heightmap`` on the GPU (``rover_mesh_to_heightmap``); returns the device tensor."""
    from . import torch_ops  # noqa: F401  (registers torch.ops.rover_b200)

    with torch.cuda.device(dev):
        tv = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(dev)
        tf = torch.from_numpy(np.ascontiguousarray(f, dtype=np.int32)).to(dev)
        hm = torch.full(shape, -99.0, dtype=torch.float32, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.ops.rover_b200.mesh_to_heightmap(tv, tf, float(min_x), float(min_y), float(csx), float(csy), hm, bad)
        if int(bad.item()):
            raise IndexError("face outside the heightmap beyond the wrap range (the reference raises here too)")
    return hm


def steep_mask(heightmap: np.ndarray, threshold: float = GRADIENT_THRESHOLD, device="cpu") -> np.ndarray:
    """terrain_utils.py:265-279: Sobel gradient magnitude (wrap boundary, float64) > threshold -> bool ``[H, W]``.
    On a CUDA device the stencil runs in ``rover_steep_mask`` (csrc/terrain_build.cu)."""
    dev = torch.device(device)
    if dev.type == "cuda":
        from . import torch_ops  # noqa: F401

        with torch.cuda.device(dev):
            hm = torch.from_numpy(np.ascontiguousarray(heightmap, dtype=np.float32)).to(dev)
            return torch.ops.rover_b200.steep_mask(hm, float(threshold)).cpu().numpy().astype(bool)
    from scipy.signal import convolve2d

    kx = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]])
    gx = convolve2d(heightmap, kx, mode="same", boundary="wrap")
    gy = convolve2d(heightmap, kx.T, mode="same", boundary="wrap")
    return np.sqrt(gx**2 + gy**2) > threshold


def find_rocks_in_heightmap(heightmap: np.ndarray, threshold: float = GRADIENT_THRESHOLD, device="cpu"):
    """terrain_utils.py:265-311: Sobel gradient magnitude (wrap boundary) > threshold, then
    close 3x3 -> fill holes -> open 7x7 -> dilate 11x11 (= rock mask) -> dilate 42x42 (= safe mask).
    On a CUDA ``device`` the whole chain runs on the GPU (``csrc/terrain_build.cu``: the gradient stencil, separable box
    morphology with OpenCV's anchor convention, a tiled flood fill for ``binary_fill_holes``); on the host it is the
    reference's own OpenCV / scipy calls.  Same masks bit for bit (tests/test_gpu_terrain_build.py)."""
    dev = torch.device(device)
    if dev.type == "cuda":
        from . import torch_ops  # noqa: F401  (registers torch.ops.rover_b200)

        R = torch.ops.rover_b200
        with torch.cuda.device(dev):
            hm = torch.from_numpy(np.ascontiguousarray(heightmap, dtype=np.float32)).to(dev)
            m = R.steep_mask(hm, float(threshold))
            m = R.morph_box(R.morph_box(m, 3, False), 3, True)      # MORPH_CLOSE = dilate, erode
            m = R.fill_holes(m)
            m = R.morph_box(R.morph_box(m, 7, True), 7, False)      # MORPH_OPEN = erode, dilate
            rock = R.morph_box(m, 11, False)
            safe = R.morph_box(rock, 42, False)
            return rock.cpu().numpy(), safe.cpu().numpy()
    import cv2
    from scipy import ndimage

    steep = steep_mask(heightmap, threshold, device)
    ones = lambda k: np.ones((k, k), np.uint8)  # noqa: E731
    mask = cv2.morphologyEx(steep.astype(np.uint8), cv2.MORPH_CLOSE, ones(3))
    mask = ndimage.binary_fill_holes(mask).astype(np.uint8)
    mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, ones(7))
    rock = cv2.dilate(mask, ones(11), iterations=1)
    safe = cv2.dilate(rock, ones(42), iterations=1)
    return rock, safe


def random_rover_spawns(rock_mask: np.ndarray, heightmap: np.ndarray, min_x: float, min_y: float, n_spawns: int,
                        resolution: float = HEIGHTMAP_RESOLUTION, border_offset: float = SPAWN_BORDER_OFFSET,
                        seed: int | None = SPAWN_SEED) -> np.ndarray:
    """terrain_utils.py:330-385: legacy ``np.random.seed(seed)``; per spawn draw ``x, y = randint(lo, hi)``
    until ``rock_mask[y, x] == 0``; row = ``(x*res + min_x, y*res + min_y, heightmap[y, x])`` in fp32.
    The reference draws one candidate per Python iteration; here candidates are drawn in blocks -- the legacy
    ``RandomState.randint`` stream is the same whether it is consumed by scalar calls or as an ``(M, 2)`` array
    (checked in tests/test_oracle_golden.py against the reference's table) -- and the first ``n_spawns`` accepted
    ones, in order, are kept.  The global generator is left exactly where the reference's loop would leave it."""
    height, width = rock_mask.shape
    lo = int(border_offset / resolution)
    hi = int(min(height, width) - lo)
    if not (hi < width and hi < height):
        raise AssertionError(f"max_xy ({hi}) must be less than width/height ({width}, {height})")
    if hi <= lo:
        raise ValueError("terrain too small for the spawn border offset")
    if seed is not None:
        np.random.seed(seed)
    state0 = np.random.get_state()
    xs, ys, consumed, accepted = [], [], 0, 0
    while accepted < n_spawns:
        block = max(2 * (n_spawns - accepted), 1024)
        cand = np.random.randint(lo, hi, size=(block, 2))
        ok = rock_mask[cand[:, 1], cand[:, 0]] == 0
        idx = np.nonzero(ok)[0][: n_spawns - accepted]
        xs.append(cand[idx, 0])
        ys.append(cand[idx, 1])
        accepted += len(idx)
        if accepted >= n_spawns:
            consumed += int(idx[-1]) + 1 if len(idx) else block
        else:
            consumed += block
            if not ok.any() and consumed > 50_000_000:
                raise RuntimeError("no valid spawn location found (mask covers the spawn area)")
    # leave the legacy generator where the reference's scalar loop would: `consumed` (x, y) pairs after the seed
    np.random.set_state(state0)
    if consumed:
        np.random.randint(lo, hi, size=(consumed, 2))
    x, y = np.concatenate(xs), np.concatenate(ys)
    out = np.zeros((n_spawns, 3), dtype=np.float32)
    out[:, 0], out[:, 1], out[:, 2] = x, y, heightmap[y, x]
    out[:, 0] = out[:, 0] * resolution + min_x
    out[:, 1] = out[:, 1] * resolution + min_y
    return out


@dataclass
class TerrainTables:
    """Immutable per-terrain tables, as torch tensors on the target device."""

    vertices: torch.Tensor  # [V, 3] f32
    faces: torch.Tensor  # [F, 3] i32
    heightmap: torch.Tensor  # [H, W] f32
    safe_mask: torch.Tensor  # [H, W] u8
    rock_mask: torch.Tensor  # [H, W] u8
    spawn_table: torch.Tensor  # [2N, 3] f32
    offset_xy: torch.Tensor  # [2] f32 (min_x, min_y)
    resolution: float = HEIGHTMAP_RESOLUTION

    def to(self, device) -> "TerrainTables":
        kw = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        return TerrainTables(**kw)


def build_terrain_tables(vertices: np.ndarray, faces: np.ndarray, num_envs: int, device="cpu",
                         border_offset: float = SPAWN_BORDER_OFFSET, build_device=None) -> TerrainTables:
    """What ``TerrainManager.__init__`` (terrain_utils.py:92-127) computes, for ``num_envs`` envs."""
    hm, min_x, min_y, _, _ = mesh_to_heightmap(vertices, faces, device=build_device or "cpu")
    rock, safe = find_rocks_in_heightmap(hm, device=build_device or "cpu")
    spawns = random_rover_spawns(safe, hm, min_x, min_y, n_spawns=2 * num_envs, border_offset=border_offset)
    t = torch.from_numpy
    return TerrainTables(
        vertices=t(np.ascontiguousarray(vertices, dtype=np.float32)),
        faces=t(np.ascontiguousarray(faces, dtype=np.int32)),
        heightmap=t(hm), safe_mask=t(safe.astype(np.uint8)), rock_mask=t(rock.astype(np.uint8)),
        spawn_table=t(spawns), offset_xy=torch.tensor([min_x, min_y], dtype=torch.float32),
    ).to(device)
