"""Init-time builder of the acceleration structure the height-scan kernel walks.

The reference hands its terrain mesh to ORBIT's ``RayCaster`` which converts it once into a warp BVH
(``wp.Mesh``; wired at rover_envs/envs/navigation/rover_env_cfg.py:78-86).  The height scan casts only
vertical rays (``attach_yaw_only=True`` + default direction (0,0,-1)), so the B200 structure is 2-D:
a multi-level uniform XY *home grid* of pre-processed triangle records.

* Every non-degenerate triangle is oriented CCW in XY and assigned to the finest level ``l`` (cell size
  ``c0 * 2^l``) on which its bounding box spans at most ``SPAN + 1`` cells per axis; its *home* is the
  cell of its bbox-min corner.  Records are sorted by (level, home row, home column), so the candidates
  of a ray in cell (i, j) are ``SPAN + 1`` contiguous runs -- rows ``j-SPAN..j``, columns ``i-SPAN..i``.
  No per-cell index lists, no duplication, and a patch of the map is a few contiguous byte ranges
  (what the shared-memory staging path bulk-copies).
* A record is 12 floats (3 x float4): three edge functions ``E_k = A_k*lx + B_k*ly + C_k`` and the plane
  ``z = a*lx + b*ly + c`` in the frame of the home cell's min corner (``lx = px - hx``).  Coefficients are
  computed in float64 from the float32 vertices.  ``C_k`` carries an outward bias equal to the float32
  evaluation-error bound, so that the shared edge of two triangles with different homes can never open a crack.
* The cell function ``floor((x - ox) * inv_c)`` is evaluated in float32 here exactly as on the device;
  it is monotone, so binning by the cells of the bbox corners is conservative with zero slop.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

MAX_LEVELS = 12
SPAN = 1
RECORD_FLOATS = 12


@dataclass
class ScanLevel:
    ox: float
    oy: float
    cell: float
    inv_cell: float
    ncx: int
    ncy: int
    start_offset: int  # offset of this level's (ncx*ncy+1) entries in ``cell_start``


@dataclass
class ScanGrid:
    levels: list
    cell_start: torch.Tensor  # uint32 stored as int32/int64? -> int32 tensor, record indices
    records: torch.Tensor  # [n_records, 12] f32
    span: int = SPAN
    n_triangles_in: int = 0
    n_dropped: int = 0

    @property
    def n_records(self) -> int:
        return int(self.records.shape[0])

    def to(self, device) -> "ScanGrid":
        return ScanGrid(self.levels, self.cell_start.to(device), self.records.to(device), self.span,
                        self.n_triangles_in, self.n_dropped)

    def nbytes(self) -> int:
        return self.cell_start.numel() * 4 + self.records.numel() * 4


def cell_index_f32(x: np.ndarray, origin: np.float32, inv_cell: np.float32) -> np.ndarray:
    """The device's cell function, bit-identical in float32: ``floor((x - o) * inv_c)``."""
    x = np.asarray(x, dtype=np.float32)
    return np.floor((x - np.float32(origin)) * np.float32(inv_cell)).astype(np.int64)


def default_cell_size(vertices: np.ndarray, faces: np.ndarray) -> float:
    """Level-0 cell size: slightly above the 90th-percentile triangle extent (never below 5 cm)."""
    tri = vertices[faces][:, :, :2]
    ext = (tri.max(axis=1) - tri.min(axis=1)).max(axis=1)
    if len(ext) == 0:
        return 1.0
    return float(max(np.quantile(ext, 0.9) * 1.001, 0.05))


def build_scan_grid(vertices, faces, cell_size: float | None = None, span: int = SPAN) -> ScanGrid:
    """Build the multi-level home grid on the host (numpy).  ``vertices`` [V,3] f32, ``faces`` [F,3] int."""
    v = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 3))
    f = np.ascontiguousarray(np.asarray(faces).astype(np.int64).reshape(-1, 3))
    n_in = len(f)
    if n_in and (f.min() < 0 or f.max() >= len(v)):
        raise ValueError("face index out of range")
    if n_in == 0:
        lvl = ScanLevel(0.0, 0.0, 1.0, 1.0, 1, 1, 0)
        return ScanGrid([lvl], torch.zeros(2, dtype=torch.int32), torch.zeros(0, RECORD_FLOATS), span, 0, 0)
    tri = v[f].astype(np.float64)  # [F,3,3]
    # orientation / degeneracy in XY
    e1 = tri[:, 1, :2] - tri[:, 0, :2]
    e2 = tri[:, 2, :2] - tri[:, 0, :2]
    area2 = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    finite = np.isfinite(tri).all(axis=(1, 2))
    keep = (area2 != 0.0) & finite
    flip = area2 < 0.0
    tri[flip] = tri[flip][:, [0, 2, 1], :]
    tri = tri[keep]
    n_drop = int(n_in - keep.sum())
    if len(tri) == 0:
        lvl = ScanLevel(0.0, 0.0, 1.0, 1.0, 1, 1, 0)
        return ScanGrid([lvl], torch.zeros(2, dtype=torch.int32), torch.zeros(0, RECORD_FLOATS), span, n_in, n_drop)

    lo = tri[:, :, :2].min(axis=1).astype(np.float32)
    hi = tri[:, :, :2].max(axis=1).astype(np.float32)
    c0 = np.float32(cell_size if cell_size is not None else default_cell_size(v, f[keep]))
    xmin, ymin = lo.min(axis=0)
    xmax, ymax = hi.max(axis=0)

    level_of = np.full(len(tri), -1, dtype=np.int64)
    geo = []
    for l in range(MAX_LEVELS):
        c = np.float32(c0 * np.float32(2.0**l))
        inv = np.float32(1.0) / c
        ox = np.float32(xmin - np.float32(0.5) * c)
        oy = np.float32(ymin - np.float32(0.5) * c)
        ncx = int(cell_index_f32(xmax, ox, inv)) + 1
        ncy = int(cell_index_f32(ymax, oy, inv)) + 1
        geo.append((ox, oy, c, inv, ncx, ncy))
        todo = level_of < 0
        if not todo.any():
            geo.pop()
            break
        i0 = cell_index_f32(lo[todo, 0], ox, inv)
        i1 = cell_index_f32(hi[todo, 0], ox, inv)
        j0 = cell_index_f32(lo[todo, 1], oy, inv)
        j1 = cell_index_f32(hi[todo, 1], oy, inv)
        fits = ((i1 - i0) <= span) & ((j1 - j0) <= span)
        if l == MAX_LEVELS - 1:
            fits[:] = True  # last level takes the rest; its cells are >= 2048 * c0 wide
        idx = np.nonzero(todo)[0][fits]
        level_of[idx] = l
    if (level_of < 0).any():
        raise RuntimeError("triangle larger than the coarsest level")
    if level_of.max() == MAX_LEVELS - 1:
        # the catch-all level must really satisfy the span rule, otherwise candidates would be missed
        ox, oy, c, inv, ncx, ncy = geo[-1]
        sel = level_of == MAX_LEVELS - 1
        if ((cell_index_f32(hi[sel, 0], ox, inv) - cell_index_f32(lo[sel, 0], ox, inv)) > span).any() or \
                ((cell_index_f32(hi[sel, 1], oy, inv) - cell_index_f32(lo[sel, 1], oy, inv)) > span).any():
            raise RuntimeError("mesh extent exceeds the multi-level grid range; raise cell_size")

    levels, starts, recs = [], [], []
    rec_base = 0
    start_off = 0
    for l, (ox, oy, c, inv, ncx, ncy) in enumerate(geo):
        sel = np.nonzero(level_of == l)[0]
        if len(sel) == 0:
            continue
        hi_ = cell_index_f32(lo[sel, 0], ox, inv)
        hj_ = cell_index_f32(lo[sel, 1], oy, inv)
        key = hj_ * ncx + hi_
        order = np.argsort(key, kind="stable")
        sel, key, hi_, hj_ = sel[order], key[order], hi_[order], hj_[order]
        counts = np.bincount(key, minlength=ncx * ncy)
        st = np.zeros(ncx * ncy + 1, dtype=np.int64)
        np.cumsum(counts, out=st[1:])
        starts.append(st + rec_base)
        # home-cell frame, computed in float32 exactly as the device does: hx = ox + i*c
        hx = (np.float32(ox) + hi_.astype(np.float32) * np.float32(c)).astype(np.float32).astype(np.float64)
        hy = (np.float32(oy) + hj_.astype(np.float32) * np.float32(c)).astype(np.float32).astype(np.float64)
        recs.append(_records(tri[sel], hx, hy, float(c), span))
        levels.append(ScanLevel(float(ox), float(oy), float(c), float(inv), ncx, ncy, start_off))
        start_off += ncx * ncy + 1
        rec_base += len(sel)
    cell_start = np.concatenate(starts)
    if cell_start.max() >= 2**31:
        raise RuntimeError("too many records for 32-bit offsets")
    return ScanGrid(levels, torch.from_numpy(cell_start.astype(np.int32)),
                    torch.from_numpy(np.concatenate(recs).astype(np.float32)), span, n_in, n_drop)


def _records(tri: np.ndarray, hx: np.ndarray, hy: np.ndarray, cell: float, span: int) -> np.ndarray:
    """12-float records in the home-cell frame; float64 math, float32 storage."""
    p = tri.copy()
    p[:, :, 0] -= hx[:, None]
    p[:, :, 1] -= hy[:, None]
    out = np.empty((len(p), RECORD_FLOATS), dtype=np.float64)
    reach = (span + 1) * cell  # |lx|, |ly| of any ray that tests this record
    eps32 = float(np.finfo(np.float32).eps)
    for k in range(3):
        a = p[:, k, :]
        b = p[:, (k + 1) % 3, :]
        # canonical direction (lexicographic on the float32 vertices) so that both triangles sharing the
        # edge derive their coefficients from identical operands before the sign flip
        swap = (a[:, 0] > b[:, 0]) | ((a[:, 0] == b[:, 0]) & (a[:, 1] > b[:, 1]))
        ax = np.where(swap, b[:, 0], a[:, 0])
        ay = np.where(swap, b[:, 1], a[:, 1])
        bx = np.where(swap, a[:, 0], b[:, 0])
        by = np.where(swap, a[:, 1], b[:, 1])
        A = -(by - ay)
        B = bx - ax
        C = -(A * ax + B * ay)
        sgn = np.where(swap, -1.0, 1.0)
        A, B, C = A * sgn, B * sgn, C * sgn
        # outward bias = first-order bound on the float32 error of fma(A, lx, fma(B, ly, C)) over the reach,
        # including the rounding of the stored coefficients:  2^-24 (|A lx| + 2 |B ly| + 2 |C|), times 1.25.
        # A ray on a shared edge is then accepted by both neighbours whatever their home frames: no cracks.
        # The price: a triangle is fatter by bias / |edge| (1e-8 m at a 0.2 m triangle, ~3e-5 m at a 40 m one).
        bias = 1.25 * 0.5 * eps32 * (np.abs(A) * reach + 2.0 * np.abs(B) * reach + 2.0 * np.abs(C))
        out[:, 3 * k + 0] = A
        out[:, 3 * k + 1] = B
        out[:, 3 * k + 2] = C + bias
    # plane through the three vertices: z = a*lx + b*ly + c
    e1 = p[:, 1, :] - p[:, 0, :]
    e2 = p[:, 2, :] - p[:, 0, :]
    nx = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
    ny = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
    nz = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    a = -nx / nz
    b = -ny / nz
    c = p[:, 0, 2] - a * p[:, 0, 0] - b * p[:, 0, 1]
    out[:, 9], out[:, 10], out[:, 11] = a, b, c
    return out
