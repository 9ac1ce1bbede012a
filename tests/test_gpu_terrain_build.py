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
