"""Init-time terrain tables built on the GPU against the reference's golden tables and against the
host builder: bit-exact (the heightmap is a max over faces, the steep mask a thresholded float64 stencil)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from isaac_rover_orbit_b200 import terrain as TR

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "terrain_command.npz")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def golden_mesh():
    g = np.load(GOLDEN)
    v, f = TR.make_synthetic_terrain(float(g["terrain_size_m"]), float(g["terrain_grid_res"]), int(g["terrain_seed"]))
    return g, v, f


def test_heightmap_kernel_equals_reference_golden(cuda_device, golden_mesh):
    g, v, f = golden_mesh
    hm, min_x, min_y, _, _ = TR.mesh_to_heightmap(v, f, device=cuda_device)
    assert hm.dtype == np.float32
    assert _sha(hm) == str(g["ref_heightmap_sha"])  # sha of the table the unmodified reference produced
    assert np.array_equal(hm[::8, ::8], g["ref_heightmap_dec"])
    assert np.allclose([min_x, min_y], g["ref_min_xy"])


def test_rock_and_safe_masks_with_gpu_gradient_equal_golden(cuda_device, golden_mesh):
    g, v, f = golden_mesh
    hm, *_ = TR.mesh_to_heightmap(v, f, device=cuda_device)
    steep_gpu = TR.steep_mask(hm, device=cuda_device)
    assert np.array_equal(steep_gpu, TR.steep_mask(hm, device="cpu"))
    assert 0 < steep_gpu.sum() < steep_gpu.size
    rock, safe = TR.find_rocks_in_heightmap(hm, device=cuda_device)
    assert _sha(rock.astype(np.uint8)) == str(g["ref_rock_sha"])
    assert _sha(safe.astype(np.uint8)) == str(g["ref_safe_sha"])


def test_tables_built_on_gpu_equal_host_tables(cuda_device, golden_mesh):
    _, v, f = golden_mesh
    a = TR.build_terrain_tables(v, f, 64, build_device=cuda_device)
    b = TR.build_terrain_tables(v, f, 64, build_device="cpu")
    for name in ("heightmap", "safe_mask", "rock_mask", "spawn_table", "offset_xy"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name


def test_heightmap_kernel_quirks_wrap_clamp_and_negative_heights(cuda_device):
    """Faces in the lower border wrap to the far side (negative Python indices), the upper index is clamped, heights
    below zero and -0.0 survive the atomic max; host builder and kernel agree bit for bit."""
    rng = np.random.default_rng(5)
    n = 41
    xs, ys = np.meshgrid(np.linspace(0, 8, n, dtype=np.float32), np.linspace(0, 6, n, dtype=np.float32))
    z = rng.uniform(-2.0, 0.5, size=xs.shape).astype(np.float32)
    z[::7, ::5] = -0.0
    v = np.stack([xs.ravel(), ys.ravel(), z.ravel()], axis=1)
    idx = np.arange(n * n).reshape(n, n)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    f = np.concatenate([np.stack([a, b, d], 1), np.stack([a, d, c], 1)]).astype(np.int32)
    host, *_ = TR.mesh_to_heightmap(v, f, device="cpu")
    dev, *_ = TR.mesh_to_heightmap(v, f, device=cuda_device)
    assert np.array_equal(host.view(np.uint32), dev.view(np.uint32))
    assert (dev[-20:, :] > -99).any() and (dev[:, -20:] > -99).any()  # the wrapped border rows / columns got data


def test_heightmap_kernel_reports_faces_beyond_the_wrap_range(cuda_device):
    """A mesh narrower than ~3 m: after the 1 m border shrink the table has fewer cells than the border is wide, so the
    lower border's indices fall below -size -- the reference's heightmap[j, i] raises IndexError, and so do both builders."""
    v = np.array([[0, 0, 0], [2.5, 0, 0], [0, 2.5, 0], [2.5, 2.5, 0]], np.float32)
    f = np.array([[0, 1, 2], [1, 3, 2]], np.int32)
    with pytest.raises(IndexError):
        TR.mesh_to_heightmap(v, f, device="cpu")
    with pytest.raises(IndexError):
        TR.mesh_to_heightmap(v, f, device=cuda_device)
    hm, *_ = TR.mesh_to_heightmap(v * 4.0, f, device=cuda_device)  # 10 m: fine
    assert hm.max() == 0.0 and hm.min() == 0.0


def test_builder_entry_points_reject_bad_arguments(cuda_device):
    import ctypes as C

    from isaac_rover_orbit_b200 import _lib

    lib = _lib.load()
    assert lib.rover_mesh_to_heightmap(None, None, 1, 0.0, 0.0, 1.0, 1.0, 4, 4, None, None, None) != 0
    assert b"NULL" in lib.rover_last_error()
    t = torch.zeros(16, device=cuda_device)
    assert lib.rover_steep_mask(C.c_void_p(t.data_ptr()), 0, 4, 0.3, C.c_void_p(t.data_ptr()), None) != 0


@pytest.mark.parametrize("shape,seed", [((97, 131), 1), ((640, 515), 2), ((33, 1200), 3)])
def test_morphology_kernels_equal_opencv_and_scipy(cuda_device, shape, seed):
    """rover_morph_box = cv2.dilate / cv2.erode with a box of ones (anchor k / 2: the 42 x 42 box is asymmetric; outside
    pixels neutral), rover_fill_holes = scipy.ndimage.binary_fill_holes -- on blobs with holes, nested holes, holes
    touching the border and long corridors (many tile-to-tile relaxation rounds)."""
    import cv2
    from scipy import ndimage

    rng = np.random.default_rng(seed)
    img = (ndimage.gaussian_filter(rng.random(shape), 3.0) > 0.5).astype(np.uint8)
    img[shape[0] // 2, :] = 1      # a wall across the image ...
    img[shape[0] // 2, 5] = 0      # ... with one gap: the background below connects through it
    img[3:30, 3:30] = 1
    img[8:25, 8:25] = 0            # a hole ...
    img[12:20, 12:20] = 1          # ... with an island ...
    img[15:17, 15:17] = 0          # ... that has its own hole
    spiral = np.zeros((64, 64), np.uint8)  # a long corridor: the flood has to walk it tile by tile
    for k in range(0, 28, 4):
        spiral[k, k:64 - k] = 1
        spiral[k:64 - k, 63 - k] = 1
        spiral[63 - k, k:64 - k] = 1
        spiral[k + 4:64 - k, k] = 1
    if shape[0] > 100 and shape[1] > 100:
        img[-70:-6, -70:-6] = spiral
    R = torch.ops.rover_b200
    d = torch.from_numpy(img).to(cuda_device)
    ones = lambda k: np.ones((k, k), np.uint8)  # noqa: E731
    for k in (1, 3, 7, 11, 42):
        assert np.array_equal(R.morph_box(d, k, False).cpu().numpy(), cv2.dilate(img, ones(k), iterations=1)), ("dilate", k)
        assert np.array_equal(R.morph_box(d, k, True).cpu().numpy(), cv2.erode(img, ones(k), iterations=1)), ("erode", k)
    close = R.morph_box(R.morph_box(d, 3, False), 3, True).cpu().numpy()
    assert np.array_equal(close, cv2.morphologyEx(img, cv2.MORPH_CLOSE, ones(3)))
    opened = R.morph_box(R.morph_box(d, 7, True), 7, False).cpu().numpy()
    assert np.array_equal(opened, cv2.morphologyEx(img, cv2.MORPH_OPEN, ones(7)))
    filled = R.fill_holes(d).cpu().numpy()
    assert np.array_equal(filled, ndimage.binary_fill_holes(img).astype(np.uint8))
    assert filled.sum() > img.sum(), "the fixture must contain holes"
    assert R.fill_holes(torch.zeros_like(d)).sum() == 0 and bool((R.fill_holes(torch.ones_like(d)) == 1).all())


def test_plane_cells_built_on_the_device_equal_the_host_builder(cuda_device, golden_mesh):
    from isaac_rover_orbit_b200 import ops
    from isaac_rover_orbit_b200.plane_cells import build_plane_cells, build_plane_cells_torch

    _, v, f = golden_mesh
    f2 = np.delete(f, [5, 6, 777], axis=0)  # holes: cells with one triangle that does not cover them -> general
    f2[10] = f2[10][[0, 2, 1]]              # a clockwise triangle
    for faces in (f, f2):
        host, dev = build_plane_cells(v, faces), build_plane_cells_torch(v, faces, cuda_device)
        assert dev is not None and dev.entries.is_cuda
        assert torch.equal(host.xs, dev.xs.cpu()) and torch.equal(host.ys, dev.ys.cpu())
        assert torch.equal(host.entries, dev.entries.cpu())
        assert (host.inv_dx, host.inv_dy, host.n_general, host.n_empty) == (dev.inv_dx, dev.inv_dy, dev.n_general, dev.n_empty)
    # a mesh that is not a lattice is left to the host builder
    rng = np.random.default_rng(1)
    vv = rng.random((300, 3)).astype(np.float32) * 10
    ff = rng.integers(0, 300, (200, 3)).astype(np.int32)
    assert build_plane_cells_torch(vv, ff, cuda_device) is None
    h = ops.ScanGridHandle.from_mesh(v, f, cuda_device)
    assert not h.has_home_grid and h.cells.n_general == 0  # DEM terrain: the home grid is never built
    h2 = ops.ScanGridHandle.from_mesh(v, f2, cuda_device)
    assert h2.has_home_grid and h2.cells.n_general > 0
