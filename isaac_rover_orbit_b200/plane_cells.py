"""Init-time builder of the *plane-cell* table: the fast path of the height-scan kernel.

For vertical rays the mesh is a planar map: above every XY point sits at most a handful of triangles.  Lay a
rectilinear grid on the mesh (lines ``xs``/``ys``); for the cells of that grid whose surface is fully described by
ONE triangle, or by TWO triangles that share an edge and together cover the cell, the closest hit of any ray in
the cell is a closed form -- no candidate loop, no per-candidate edge tests:

    E  = A*lx + B*ly + C            (shared edge, E >= 0 on the first triangle's side)
    z  = a*lx + b*ly + c + k*min(E, 0)      (plane of triangle 1; ``k*E`` is the difference to plane 2,
                                             exact because both planes agree on the line E = 0)

with ``(lx, ly)`` relative to the cell's min corner.  An entry is 8 floats (two float4): ``a, b, c, k | A, B, C, tag``.
``tag`` is 0 for closed-form cells (empty cells are closed-form too: ``c = -inf``) and 1 for *general* cells, whose
rays take the home-grid walk of :mod:`scan_grid` instead.  Classification is conservative -- any doubt makes a cell
general -- so the table never changes a result, only how fast it is produced.

Grid lines: when the vertices sit on a rectilinear lattice (DEM-style terrains: the reference's Mars terrains and
the synthetic stand-in) the lattice itself is used, every quad becomes one two-plane cell; otherwise uniform lines
at the home grid's level-0 cell size are used and most cells stay general.

All predicates are evaluated in float64 on float32 inputs: coordinate differences and their pairwise products are
exact, so "touches the cell boundary" (exactly 0) is distinguished from "overlaps the open cell".
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

ENTRY_FLOATS = 8
MAX_PAIR_CELLS = 64_000_000  # refuse pathological (triangle x cell) enumerations


@dataclass
class PlaneCells:
    xs: torch.Tensor  # [nx+1] f32 grid lines (strictly increasing)
    ys: torch.Tensor  # [ny+1] f32
    entries: torch.Tensor  # [ny, nx, 8] f32
    inv_dx: float  # (nx) / (xs[-1] - xs[0]): arithmetic first guess of the column
    inv_dy: float
    n_general: int = 0
    n_empty: int = 0
    lattice: bool = False

    @property
    def nx(self) -> int:
        return int(self.xs.numel() - 1)

    @property
    def ny(self) -> int:
        return int(self.ys.numel() - 1)

    def nbytes(self) -> int:
        return self.entries.numel() * 4 + self.xs.numel() * 4 + self.ys.numel() * 4


def _lattice_lines(coord: np.ndarray, n_vertices: int):
    u = np.unique(coord)
    if 2 <= len(u) <= 4 * int(np.sqrt(n_vertices)) + 16:
        return u.astype(np.float32)
    return None


def _uniform_lines(lo: float, hi: float, cell: float) -> np.ndarray:
    n = max(int(np.ceil((hi - lo) / cell)), 1)
    lines = (np.float64(lo) + np.arange(n + 1, dtype=np.float64) * (np.float64(hi) - np.float64(lo)) / n)
    lines = lines.astype(np.float32)
    lines[0], lines[-1] = np.float32(lo), np.float32(hi)
    return np.unique(lines)


def _edge_fn(p, q):
    """Edge function of the directed edge p->q (interior of a CCW triangle on the left): returns A, B, C."""
    A = -(q[..., 1] - p[..., 1])
    B = q[..., 0] - p[..., 0]
    C = -(A * p[..., 0] + B * p[..., 1])
    return A, B, C


def _tri_overlaps_open_cell(tri, x0, x1, y0, y1):
    """Exact 2-D SAT: does the (closed, CCW) triangle overlap the OPEN box with positive area?  Vectorised."""
    tx, ty = tri[..., 0], tri[..., 1]
    ok = (tx.max(-1) > x0) & (tx.min(-1) < x1) & (ty.max(-1) > y0) & (ty.min(-1) < y1)
    for k in range(3):
        A, B, C = _edge_fn(tri[..., k, :], tri[..., (k + 1) % 3, :])
        # max of the edge function over the box corners; the box is on the outer side iff that max <= 0
        m = np.maximum(np.maximum(A * x0 + B * y0, A * x1 + B * y0), np.maximum(A * x0 + B * y1, A * x1 + B * y1)) + C
        ok &= m > 0
    return ok


def build_plane_cells(vertices, faces, fallback_cell: float = 0.2) -> PlaneCells:
    v = np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 3))
    f = np.ascontiguousarray(np.asarray(faces).astype(np.int64).reshape(-1, 3))
    if len(f) == 0 or len(v) == 0:
        z = torch.zeros
        ent = z(1, 1, ENTRY_FLOATS)
        ent[..., 2] = -float("inf")
        return PlaneCells(torch.tensor([0.0, 1.0]), torch.tensor([0.0, 1.0]), ent, 1.0, 1.0, 0, 1, False)
    tri = v[f].astype(np.float64)
    e1 = tri[:, 1, :2] - tri[:, 0, :2]
    e2 = tri[:, 2, :2] - tri[:, 0, :2]
    area2 = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    keep = (area2 != 0.0) & np.isfinite(tri).all(axis=(1, 2))
    flip = area2 < 0.0
    tri[flip] = tri[flip][:, [0, 2, 1], :]
    tri = tri[keep]
    used = np.unique(f[keep])
    xs = _lattice_lines(v[used, 0], len(used)) if len(used) else None
    ys = _lattice_lines(v[used, 1], len(used)) if len(used) else None
    lattice = xs is not None and ys is not None
    if not lattice:
        lo = tri[:, :, :2].min(axis=(0, 1)) if len(tri) else np.zeros(2)
        hi = tri[:, :, :2].max(axis=(0, 1)) if len(tri) else np.ones(2)
        xs = _uniform_lines(lo[0], max(hi[0], lo[0] + fallback_cell), fallback_cell)
        ys = _uniform_lines(lo[1], max(hi[1], lo[1] + fallback_cell), fallback_cell)
    nx, ny = len(xs) - 1, len(ys) - 1
    xs64, ys64 = xs.astype(np.float64), ys.astype(np.float64)

    # ---- (cell, triangle) pairs with positive-area overlap of the open cell
    bx0, bx1 = tri[:, :, 0].min(1), tri[:, :, 0].max(1)
    by0, by1 = tri[:, :, 1].min(1), tri[:, :, 1].max(1)
    i0 = np.clip(np.searchsorted(xs64, bx0, side="right") - 1, 0, nx - 1)
    i1 = np.clip(np.searchsorted(xs64, bx1, side="left") - 1, 0, nx - 1)
    j0 = np.clip(np.searchsorted(ys64, by0, side="right") - 1, 0, ny - 1)
    j1 = np.clip(np.searchsorted(ys64, by1, side="left") - 1, 0, ny - 1)
    wi, wj = i1 - i0 + 1, j1 - j0 + 1
    npairs = int((wi * wj).sum())
    if npairs > MAX_PAIR_CELLS:
        raise RuntimeError(f"plane-cell build: {npairs} (triangle, cell) pairs; use a coarser fallback_cell")
    t_idx = np.repeat(np.arange(len(tri)), wi * wj)
    start = np.cumsum(wi * wj) - wi * wj
    local = np.arange(npairs) - np.repeat(start, wi * wj)
    wi_r = np.repeat(wi, wi * wj)
    ci = np.repeat(i0, wi * wj) + local % wi_r
    cj = np.repeat(j0, wi * wj) + local // wi_r
    ov = _tri_overlaps_open_cell(tri[t_idx], xs64[ci], xs64[ci + 1], ys64[cj], ys64[cj + 1])
    t_idx, ci, cj = t_idx[ov], ci[ov], cj[ov]
    cell_id = cj * nx + ci
    order = np.argsort(cell_id, kind="stable")
    cell_id, t_idx = cell_id[order], t_idx[order]
    count = np.bincount(cell_id, minlength=nx * ny)
    first = np.cumsum(count) - count

    ent = np.zeros((ny * nx, ENTRY_FLOATS), dtype=np.float64)
    ent[:, 2] = -np.inf  # empty cell: closed form that never hits
    ent[:, 6] = 1.0  # E = 1 > 0 : plane 1 everywhere
    general = count > 2
    cx0 = np.tile(xs64[:-1], ny)
    cy0 = np.repeat(ys64[:-1], nx)
    cx1 = np.tile(xs64[1:], ny)
    cy1 = np.repeat(ys64[1:], nx)

    def plane(t, ox, oy):
        p = t.copy()
        p[:, :, 0] -= ox[:, None]
        p[:, :, 1] -= oy[:, None]
        u, w = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
        nxn = u[:, 1] * w[:, 2] - u[:, 2] * w[:, 1]
        nyn = u[:, 2] * w[:, 0] - u[:, 0] * w[:, 2]
        nzn = u[:, 0] * w[:, 1] - u[:, 1] * w[:, 0]
        a, b = -nxn / nzn, -nyn / nzn
        return a, b, p[:, 0, 2] - a * p[:, 0, 0] - b * p[:, 0, 1], p

    def corners_inside(A, B, C, x0, y0, x1, y1):
        return ((A * x0 + B * y0 + C >= 0) & (A * x1 + B * y0 + C >= 0) & (A * x0 + B * y1 + C >= 0) &
                (A * x1 + B * y1 + C >= 0))

    # ---- one triangle: closed form iff the whole cell lies inside it
    one = np.nonzero(count == 1)[0]
    if len(one):
        t = tri[t_idx[first[one]]]
        a, b, c, p = plane(t, cx0[one], cy0[one])
        w_, h_ = cx1[one] - cx0[one], cy1[one] - cy0[one]
        inside = np.ones(len(one), dtype=bool)
        for k in range(3):
            A, B, C = _edge_fn(p[:, k, :2], p[:, (k + 1) % 3, :2])
            inside &= corners_inside(A, B, C, 0.0, 0.0, w_, h_)
        ent[one[inside], 0], ent[one[inside], 1], ent[one[inside], 2] = a[inside], b[inside], c[inside]
        general[one[~inside]] = True

    # ---- two triangles sharing an edge whose other edges keep the whole cell on their inner side
    two = np.nonzero(count == 2)[0]
    if len(two):
        t1, t2 = tri[t_idx[first[two]]], tri[t_idx[first[two] + 1]]
        a1, b1, c1, p1 = plane(t1, cx0[two], cy0[two])
        a2, b2, c2, p2 = plane(t2, cx0[two], cy0[two])
        w_, h_ = cx1[two] - cx0[two], cy1[two] - cy0[two]
        good = np.zeros(len(two), dtype=bool)
        As, Bs, Cs = np.zeros(len(two)), np.zeros(len(two)), np.ones(len(two))
        for k in range(3):  # edge k of triangle 1 : p1[k] -> p1[k+1]
            pa, pb = p1[:, k], p1[:, (k + 1) % 3]
            for m in range(3):  # the same edge, reversed, in triangle 2 : p2[m] -> p2[m+1]
                qa, qb = p2[:, m], p2[:, (m + 1) % 3]
                shared = (pa == qb).all(-1) & (pb == qa).all(-1)  # x, y AND z: a continuous seam
                if not shared.any():
                    continue
                ok = shared & ~good
                A, B, C = _edge_fn(pa[:, :2], pb[:, :2])
                for kk in (1, 2):  # the two non-shared edges of triangle 1
                    Ae, Be, Ce = _edge_fn(p1[:, (k + kk) % 3, :2], p1[:, (k + kk + 1) % 3, :2])
                    ok &= corners_inside(Ae, Be, Ce, 0.0, 0.0, w_, h_)
                for mm in (1, 2):
                    Ae, Be, Ce = _edge_fn(p2[:, (m + mm) % 3, :2], p2[:, (m + mm + 1) % 3, :2])
                    ok &= corners_inside(Ae, Be, Ce, 0.0, 0.0, w_, h_)
                As[ok], Bs[ok], Cs[ok] = A[ok], B[ok], C[ok]
                good |= ok
        # plane 2 = plane 1 + k * E on the far side of the shared edge
        with np.errstate(divide="ignore", invalid="ignore"):
            kx = np.where(np.abs(As) >= np.abs(Bs), (a2 - a1) / As, (b2 - b1) / Bs)
        good &= np.isfinite(kx)
        # consistency of the one-parameter form (it must reproduce plane 2): guards against numerical surprises
        resid = np.abs(a1 + kx * As - a2) + np.abs(b1 + kx * Bs - b2) + np.abs(c1 + kx * Cs - c2)
        good &= resid <= 1e-9 * (1.0 + np.abs(a2) + np.abs(b2) + np.abs(c2))
        g = two[good]
        ent[g, 0], ent[g, 1], ent[g, 2], ent[g, 3] = a1[good], b1[good], c1[good], kx[good]
        ent[g, 4], ent[g, 5], ent[g, 6] = As[good], Bs[good], Cs[good]
        general[two[~good]] = True

    ent[general, 7] = 1.0
    ent[general, 2] = -np.inf
    span_x, span_y = float(xs64[-1] - xs64[0]), float(ys64[-1] - ys64[0])
    return PlaneCells(torch.from_numpy(xs.copy()), torch.from_numpy(ys.copy()),
                      torch.from_numpy(ent.astype(np.float32).reshape(ny, nx, ENTRY_FLOATS)),
                      nx / span_x, ny / span_y, int(general.sum()), int((count == 0).sum()), bool(lattice))


# ---------------------------------------------------------------------------------------------------------------------
# The same builder on a torch device (init-time preprocessing on the GPU).  A line-by-line port of
# build_plane_cells for LATTICE meshes: every float64 operation is one elementwise torch kernel (each result rounded
# once, as numpy does), so the table is bit-identical to the host builder's -- tests/test_host_cpu.py checks that on the
# CPU device, tests/test_gpu_terrain_build.py on the B200.  2,000,000 triangles: 5.2 s of numpy -> ~0.1 s.
# ---------------------------------------------------------------------------------------------------------------------
def _t_edge_fn(p, q):
    A = -(q[..., 1] - p[..., 1])
    B = q[..., 0] - p[..., 0]
    C = -(A * p[..., 0] + B * p[..., 1])
    return A, B, C


def _t_overlaps_open_cell(tri, x0, x1, y0, y1):
    tx, ty = tri[..., 0], tri[..., 1]
    ok = (tx.max(-1).values > x0) & (tx.min(-1).values < x1) & (ty.max(-1).values > y0) & (ty.min(-1).values < y1)
    mx = torch.maximum
    for k in range(3):
        A, B, C = _t_edge_fn(tri[..., k, :], tri[..., (k + 1) % 3, :])
        m = mx(mx(A * x0 + B * y0, A * x1 + B * y0), mx(A * x0 + B * y1, A * x1 + B * y1)) + C
        ok &= m > 0
    return ok


def _t_lattice_lines(coord: torch.Tensor, n_vertices: int):
    u = torch.unique(coord)
    if 2 <= u.numel() <= 4 * int(np.sqrt(n_vertices)) + 16:
        return u.to(torch.float32)
    return None


def build_plane_cells_torch(vertices, faces, device) -> PlaneCells | None:
    """``build_plane_cells`` on ``device`` for lattice meshes; ``None`` when the vertices do not sit on a rectilinear
    lattice (the caller then uses the host builder, which needs the home grid's cell size for its fallback lines)."""
    dev = torch.device(device)
    v = torch.as_tensor(np.ascontiguousarray(np.asarray(vertices, dtype=np.float32).reshape(-1, 3))).to(dev)
    f = torch.as_tensor(np.ascontiguousarray(np.asarray(faces).astype(np.int64).reshape(-1, 3))).to(dev)
    if f.numel() == 0 or v.numel() == 0:
        return None
    tri = v[f].to(torch.float64)
    e1 = tri[:, 1, :2] - tri[:, 0, :2]
    e2 = tri[:, 2, :2] - tri[:, 0, :2]
    area2 = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    keep = (area2 != 0.0) & torch.isfinite(tri).all(dim=2).all(dim=1)
    flip = area2 < 0.0
    tri[flip] = tri[flip][:, [0, 2, 1], :]
    tri = tri[keep]
    used = torch.unique(f[keep])
    if used.numel() == 0:
        return None
    xs = _t_lattice_lines(v[used, 0], used.numel())
    ys = _t_lattice_lines(v[used, 1], used.numel())
    if xs is None or ys is None:
        return None
    nx, ny = xs.numel() - 1, ys.numel() - 1
    xs64, ys64 = xs.to(torch.float64), ys.to(torch.float64)

    bx0, bx1 = tri[:, :, 0].min(1).values, tri[:, :, 0].max(1).values
    by0, by1 = tri[:, :, 1].min(1).values, tri[:, :, 1].max(1).values
    ss = torch.searchsorted
    i0 = torch.clamp(ss(xs64, bx0.contiguous(), right=True) - 1, 0, nx - 1)
    i1 = torch.clamp(ss(xs64, bx1.contiguous(), right=False) - 1, 0, nx - 1)
    j0 = torch.clamp(ss(ys64, by0.contiguous(), right=True) - 1, 0, ny - 1)
    j1 = torch.clamp(ss(ys64, by1.contiguous(), right=False) - 1, 0, ny - 1)
    wi, wj = i1 - i0 + 1, j1 - j0 + 1
    per = wi * wj
    npairs = int(per.sum())
    if npairs > MAX_PAIR_CELLS:
        raise RuntimeError(f"plane-cell build: {npairs} (triangle, cell) pairs; use a coarser fallback_cell")
    rep = torch.repeat_interleave
    t_idx = rep(torch.arange(tri.shape[0], device=dev), per)
    start = torch.cumsum(per, 0) - per
    local = torch.arange(npairs, device=dev) - rep(start, per)
    wi_r = rep(wi, per)
    ci = rep(i0, per) + local % wi_r
    cj = rep(j0, per) + torch.div(local, wi_r, rounding_mode="floor")
    ov = _t_overlaps_open_cell(tri[t_idx], xs64[ci], xs64[ci + 1], ys64[cj], ys64[cj + 1])
    t_idx, ci, cj = t_idx[ov], ci[ov], cj[ov]
    cell_id = cj * nx + ci
    order = torch.sort(cell_id, stable=True).indices
    cell_id, t_idx = cell_id[order], t_idx[order]
    count = torch.bincount(cell_id, minlength=nx * ny)
    first = torch.cumsum(count, 0) - count

    ent = torch.zeros(ny * nx, ENTRY_FLOATS, dtype=torch.float64, device=dev)
    ent[:, 2] = -float("inf")
    ent[:, 6] = 1.0
    general = count > 2
    cx0 = xs64[:-1].repeat(ny)
    cy0 = rep(ys64[:-1], nx)
    cx1 = xs64[1:].repeat(ny)
    cy1 = rep(ys64[1:], nx)

    def plane(t, ox, oy):
        p = t.clone()
        p[:, :, 0] -= ox[:, None]
        p[:, :, 1] -= oy[:, None]
        u, w = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
        nxn = u[:, 1] * w[:, 2] - u[:, 2] * w[:, 1]
        nyn = u[:, 2] * w[:, 0] - u[:, 0] * w[:, 2]
        nzn = u[:, 0] * w[:, 1] - u[:, 1] * w[:, 0]
        a, b = -nxn / nzn, -nyn / nzn
        return a, b, p[:, 0, 2] - a * p[:, 0, 0] - b * p[:, 0, 1], p

    def corners_inside(A, B, C, x0, y0, x1, y1):
        return ((A * x0 + B * y0 + C >= 0) & (A * x1 + B * y0 + C >= 0) & (A * x0 + B * y1 + C >= 0) &
                (A * x1 + B * y1 + C >= 0))

    one = torch.nonzero(count == 1).squeeze(-1)
    if one.numel():
        t = tri[t_idx[first[one]]]
        a, b, c, p = plane(t, cx0[one], cy0[one])
        w_, h_ = cx1[one] - cx0[one], cy1[one] - cy0[one]
        inside = torch.ones(one.numel(), dtype=torch.bool, device=dev)
        for k in range(3):
            A, B, C = _t_edge_fn(p[:, k, :2], p[:, (k + 1) % 3, :2])
            inside &= corners_inside(A, B, C, 0.0, 0.0, w_, h_)
        sel = one[inside]
        ent[sel, 0], ent[sel, 1], ent[sel, 2] = a[inside], b[inside], c[inside]
        general[one[~inside]] = True

    two = torch.nonzero(count == 2).squeeze(-1)
    if two.numel():
        t1, t2 = tri[t_idx[first[two]]], tri[t_idx[first[two] + 1]]
        a1, b1, c1, p1 = plane(t1, cx0[two], cy0[two])
        a2, b2, c2, p2 = plane(t2, cx0[two], cy0[two])
        w_, h_ = cx1[two] - cx0[two], cy1[two] - cy0[two]
        good = torch.zeros(two.numel(), dtype=torch.bool, device=dev)
        As = torch.zeros(two.numel(), dtype=torch.float64, device=dev)
        Bs = torch.zeros_like(As)
        Cs = torch.ones_like(As)
        for k in range(3):
            pa, pb = p1[:, k], p1[:, (k + 1) % 3]
            for m in range(3):
                qa, qb = p2[:, m], p2[:, (m + 1) % 3]
                shared = (pa == qb).all(-1) & (pb == qa).all(-1)
                if not bool(shared.any()):
                    continue
                ok = shared & ~good
                A, B, C = _t_edge_fn(pa[:, :2], pb[:, :2])
                for kk in (1, 2):
                    Ae, Be, Ce = _t_edge_fn(p1[:, (k + kk) % 3, :2], p1[:, (k + kk + 1) % 3, :2])
                    ok &= corners_inside(Ae, Be, Ce, 0.0, 0.0, w_, h_)
                for mm in (1, 2):
                    Ae, Be, Ce = _t_edge_fn(p2[:, (m + mm) % 3, :2], p2[:, (m + mm + 1) % 3, :2])
                    ok &= corners_inside(Ae, Be, Ce, 0.0, 0.0, w_, h_)
                As[ok], Bs[ok], Cs[ok] = A[ok], B[ok], C[ok]
                good |= ok
        kx = torch.where(As.abs() >= Bs.abs(), (a2 - a1) / As, (b2 - b1) / Bs)
        good &= torch.isfinite(kx)
        resid = (a1 + kx * As - a2).abs() + (b1 + kx * Bs - b2).abs() + (c1 + kx * Cs - c2).abs()
        good &= resid <= 1e-9 * (1.0 + a2.abs() + b2.abs() + c2.abs())
        g = two[good]
        ent[g, 0], ent[g, 1], ent[g, 2], ent[g, 3] = a1[good], b1[good], c1[good], kx[good]
        ent[g, 4], ent[g, 5], ent[g, 6] = As[good], Bs[good], Cs[good]
        general[two[~good]] = True

    ent[general, 7] = 1.0
    ent[general, 2] = -float("inf")
    span_x, span_y = float(xs64[-1] - xs64[0]), float(ys64[-1] - ys64[0])
    return PlaneCells(xs.clone(), ys.clone(), ent.to(torch.float32).reshape(ny, nx, ENTRY_FLOATS), nx / span_x, ny / span_y,
                      int(general.sum()), int((count == 0).sum()), True)
