"""Test helper: numpy emulation of the height-scan kernel's walk over the home grid.

Mirrors ``csrc/height_scan.cu::cast_down`` (same float32 cell function, same record test) so that the
host-side builder ``scan_grid.build_scan_grid`` can be validated against the oracle raycast on a CPU-only
box.  It is NOT a product path: nothing in ``isaac_rover_orbit_b200`` imports it.
"""
import numpy as np

from isaac_rover_orbit_b200.scan_grid import ScanGrid, cell_index_f32


def cast_down(grid: ScanGrid, X, Y, Z, max_d=100.0):
    X = np.asarray(X, np.float32)
    Y = np.asarray(Y, np.float32)
    Z = np.asarray(Z, np.float32)
    best = np.full(X.shape, -np.inf, dtype=np.float32)
    cs = grid.cell_start.numpy()
    rec = grid.records.numpy()
    f32 = np.float32
    for lv in grid.levels:
        i = cell_index_f32(X, f32(lv.ox), f32(lv.inv_cell))
        j = cell_index_f32(Y, f32(lv.oy), f32(lv.inv_cell))
        for dj in range(grid.span + 1):
            for di in range(grid.span + 1):
                ii, jj = i - di, j - dj
                ok = (ii >= 0) & (ii < lv.ncx) & (jj >= 0) & (jj < lv.ncy)
                iic, jjc = np.clip(ii, 0, lv.ncx - 1), np.clip(jj, 0, lv.ncy - 1)
                idx = lv.start_offset + jjc * lv.ncx + iic
                b = np.where(ok, cs[idx], 0)
                e = np.where(ok, cs[idx + 1], 0)
                hx = (f32(lv.ox) + iic.astype(f32) * f32(lv.cell)).astype(f32)
                hy = (f32(lv.oy) + jjc.astype(f32) * f32(lv.cell)).astype(f32)
                lx, ly = (X - hx).astype(f32), (Y - hy).astype(f32)
                for k in range(int((e - b).max()) if len(b) else 0):
                    act = (b + k) < e
                    r = rec[np.where(act, b + k, 0)]
                    # fp32 FMA emulated in float64 then rounded (differences are far below the test tolerances)
                    fma = lambda a, x, c: (a.astype(np.float64) * x + c).astype(f32)  # noqa: E731
                    e0 = fma(r[:, 0], lx, fma(r[:, 1], ly, r[:, 2]))
                    e1 = fma(r[:, 3], lx, fma(r[:, 4], ly, r[:, 5]))
                    e2 = fma(r[:, 6], lx, fma(r[:, 7], ly, r[:, 8]))
                    z = fma(r[:, 9], lx, fma(r[:, 10], ly, r[:, 11]))
                    t = Z - z
                    hit = act & (np.minimum(e0, np.minimum(e1, e2)) >= 0) & (t >= 0) & (t < max_d)
                    best = np.where(hit, np.maximum(best, z), best)
    return best


def cast_down_cells(cells, grid: ScanGrid, X, Y, Z, max_d=100.0):
    """numpy emulation of ``csrc/height_scan.cu::height_scan_cells_kernel`` (variant 2)."""
    X = np.asarray(X, np.float32)
    Y = np.asarray(Y, np.float32)
    Z = np.asarray(Z, np.float32)
    xs, ys = cells.xs.numpy(), cells.ys.numpy()
    ent = cells.entries.numpy().reshape(-1, 8)
    nx, ny = cells.nx, cells.ny
    i = np.clip(np.searchsorted(xs, X, side="right") - 1, 0, nx - 1)
    j = np.clip(np.searchsorted(ys, Y, side="right") - 1, 0, ny - 1)
    inside = (X >= xs[0]) & (X <= xs[-1]) & (Y >= ys[0]) & (Y <= ys[-1])
    e = ent[j * nx + i]
    lx, ly = (X - xs[i]).astype(np.float32), (Y - ys[j]).astype(np.float32)
    fma = lambda a, x, c: (a.astype(np.float64) * x + c).astype(np.float32)  # noqa: E731
    E = fma(e[:, 4], lx, fma(e[:, 5], ly, e[:, 6]))
    with np.errstate(invalid="ignore"):
        z = fma(e[:, 3], np.minimum(E, 0), fma(e[:, 0], lx, fma(e[:, 1], ly, e[:, 2])))
    t = Z - z
    closed = inside & (e[:, 7] == 0)
    out = np.where(closed & (t >= 0) & (t < max_d), z, -np.inf).astype(np.float32)
    gen = inside & (e[:, 7] != 0)
    if gen.any():
        out[gen] = cast_down(grid, X[gen], Y[gen], Z[gen], max_d)
    return out, gen
