"""numpy oracle for the LiDAR -> BEV projection (TEST INFRASTRUCTURE ONLY).

Restates ``SpatialLiDAREncoder.points_to_bev_coords`` and the index / scatter
part of ``forward_vectorized`` from the reference
(``src/models/lidar_encoder.py:42-55`` and ``:69-99``) with plain fp32 numpy
arithmetic, one IEEE rounding per reference op.

Pinned against the reference in ``tests/test_oracle_vs_reference.py`` and by
the known-answer vector of SURVEY.md section 4 (seed 123, 2x1500 points ->
1796 valid, sum(flat)=7,278,080, sha1[:16]=9d430af99ab6ca47).
The per-cell *mean* reduction is not in the reference: parity unpinned
(checked only against ``torch.scatter_reduce_(..., 'mean')``).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "range_constants", "bev_coords", "bev_cells", "bev_occupancy",
    "bev_scatter_max", "bev_scatter_mean", "bev_scatter_max_backward", "rasterize_bev", "range_cells",
]


def range_constants(point_cloud_range):
    """(x0, xspan, y0, yspan) as fp32, with the reference's dtype promotion.

    ``lidar_encoder.py:38-39`` builds ``x_range = torch.tensor([r[0], r[3]])``:
    int64 when both entries are Python ints, float32 otherwise; ``:47`` then
    forms ``x_range[1] - x_range[0]`` in that dtype and only afterwards
    promotes it to fp32 in the division.
    """
    out = []
    for lo, hi in ((point_cloud_range[0], point_cloud_range[3]),
                   (point_cloud_range[1], point_cloud_range[4])):
        if isinstance(lo, (int, np.integer)) and isinstance(hi, (int, np.integer)):
            x0 = np.float32(int(lo))
            span = np.float32(int(hi) - int(lo))
        else:
            x0 = np.float32(lo)
            span = np.float32(np.float32(hi) - np.float32(lo))
        out += [x0, span]
    return tuple(out)


def bev_coords(points: np.ndarray, point_cloud_range=(-50, -50, -5, 50, 50, 3)):
    """``points_to_bev_coords`` (``lidar_encoder.py:42-55``).

    points f32[..., >=2] -> (coords f32[..., 2], valid bool[...]).
    """
    pts = np.asarray(points, dtype=np.float32)
    x0, xs, y0, ys = range_constants(point_cloud_range)
    with np.errstate(invalid="ignore", over="ignore"):
        xn = ((pts[..., 0] - x0) / xs).astype(np.float32)   # :47  sub, then true division
        yn = ((pts[..., 1] - y0) / ys).astype(np.float32)   # :48
        valid = (xn >= 0) & (xn <= 1) & (yn >= 0) & (yn <= 1)  # :53 closed range, NaN -> False
    return np.stack([xn, yn], axis=-1), valid


def bev_cells(points: np.ndarray, grid_size=(64, 64),
              point_cloud_range=(-50, -50, -5, 50, 50, 3)) -> np.ndarray:
    """Per-point cell index ``row*W + col`` (int32), -1 for points outside.

    ``lidar_encoder.py:69-71``: ``(coords * [W-1, H-1]).long()`` (fp32 multiply,
    truncation toward zero) then clamp to the grid; ``:77-79`` row-major flat
    index.  The per-frame offset ``b*H*W`` is left to the caller.
    """
    H, W = grid_size
    coords, valid = bev_coords(points, point_cloud_range)
    with np.errstate(invalid="ignore", over="ignore"):
        gx = (coords[..., 0] * np.float32(W - 1)).astype(np.float32)
        gy = (coords[..., 1] * np.float32(H - 1)).astype(np.float32)
    gx = np.where(valid, gx, 0.0)
    gy = np.where(valid, gy, 0.0)
    col = np.clip(np.trunc(gx).astype(np.int64), 0, W - 1)
    row = np.clip(np.trunc(gy).astype(np.int64), 0, H - 1)
    cell = row * W + col
    return np.where(valid, cell, -1).astype(np.int32)


def bev_occupancy(cell: np.ndarray, grid_size=(64, 64)) -> np.ndarray:
    """Points per cell, int32[B, H*W] from cell int32[B, N]."""
    H, W = grid_size
    B = cell.shape[0]
    occ = np.zeros((B, H * W), dtype=np.int32)
    for b in range(B):
        c = cell[b]
        occ[b] = np.bincount(c[c >= 0], minlength=H * W).astype(np.int32)
    return occ


def bev_scatter_max(feats: np.ndarray, cell: np.ndarray, grid_size=(64, 64)):
    """``scatter_reduce_(amax, include_self=False)`` into a zero grid
    (``lidar_encoder.py:85-96``).

    feats f32[B, N, C] (point-major), cell int32[B, N] ->
    grid f32[B, H*W, C] (empty cells exactly 0) and
    ties int32[B, H*W, C] = number of sources equal to the max (0 if empty).
    """
    H, W = grid_size
    B, N, C = feats.shape
    grid = np.zeros((B, H * W, C), dtype=feats.dtype)
    ties = np.zeros((B, H * W, C), dtype=np.int32)
    for b in range(B):
        c = cell[b]
        sel = np.nonzero(c >= 0)[0]
        if sel.size == 0:
            continue
        order = sel[np.argsort(c[sel], kind="stable")]
        cs = c[order]
        starts = np.nonzero(np.r_[True, cs[1:] != cs[:-1]])[0]
        f = feats[b, order]
        m = np.maximum.reduceat(f, starts, axis=0)
        grid[b, cs[starts]] = m
        seg = np.repeat(np.arange(starts.size), np.diff(np.r_[starts, cs.size]))
        eq = (f == m[seg]).astype(np.int32)
        ties[b, cs[starts]] = np.add.reduceat(eq, starts, axis=0)
    return grid, ties


def bev_scatter_mean(feats: np.ndarray, cell: np.ndarray, grid_size=(64, 64)):
    """Per-cell mean (not in the reference; parity unpinned).  fp64 accumulate."""
    H, W = grid_size
    B, N, C = feats.shape
    grid = np.zeros((B, H * W, C), dtype=np.float64)
    occ = bev_occupancy(cell, grid_size)
    for b in range(B):
        c = cell[b]
        sel = c >= 0
        np.add.at(grid[b], c[sel], feats[b, sel].astype(np.float64))
    grid /= np.maximum(occ, 1)[..., None]
    return grid.astype(feats.dtype)


def bev_scatter_max_backward(grad_grid: np.ndarray, feats: np.ndarray, grid: np.ndarray,
                             ties: np.ndarray, cell: np.ndarray) -> np.ndarray:
    """Gradient of ``bev_scatter_max`` w.r.t. feats.

    ATen's ``ScatterReduceBackward`` for amax splits the cell gradient evenly
    among every source equal to the max ("amax backward
    ties"); points outside the grid get zero.

    Quirk kept for parity: ATen counts the zero-initialised ``self`` entry as
    one more tie whenever the cell max equals 0.0 (``self == result`` is
    evaluated even with ``include_self=False``), so the divisor there is
    ``ties + 1``.  In the network those gradients are then killed by the ReLU
    backward, but the op-level numbers match the reference only with it.
    """
    B, N, C = feats.shape
    out = np.zeros_like(feats)
    for b in range(B):
        c = cell[b]
        sel = np.nonzero(c >= 0)[0]
        cc = c[sel]
        hit = feats[b, sel] == grid[b, cc]
        div = (ties[b, cc] + (grid[b, cc] == 0)).astype(feats.dtype)
        out[b, sel] = np.where(hit, grad_grid[b, cc] / np.maximum(div, 1), 0).astype(feats.dtype)
    return out


def rasterize_bev(x: np.ndarray, y: np.ndarray, labels: np.ndarray, grid_size=(64, 64),
                  pc_range=(-50, 50, -50, 50)) -> np.ndarray:
    """``rasterize_bev`` (``src/data_loading/pandaset_dataset.py:23-45``), restated with the integer
    reduction the device kernel uses instead of the reference's per-point Python loop: a cell keeps the
    first non-zero label in point order = the label of its lowest-index non-zero-labelled point.

    x, y f32[N], labels int[N] -> int64[H, W].  Pinned against the reference's loop in
    ``tests/test_oracle_vs_reference.py`` and by ``tests/golden/raster_labels.npz``.
    """
    H, W = grid_size
    x_min, x_max, y_min, y_max = pc_range
    x, y = np.asarray(x, dtype=np.float32), np.asarray(y, dtype=np.float32)
    labels = np.asarray(labels).astype(np.int64)
    mask = np.zeros(H * W, dtype=np.int64)
    with np.errstate(invalid="ignore"):
        m = (x >= x_min) & (x <= x_max) & (y >= y_min) & (y <= y_max)              # :33 raw closed range
    idx = np.nonzero(m & (labels != 0))[0]                                          # :43 only non-zero labels claim
    if idx.size == 0:
        return mask.reshape(H, W)
    xs, ys = x[idx], y[idx]
    col = np.clip(((xs - x_min) / (x_max - x_min) * (W - 1)).astype(int), 0, W - 1)  # :39
    row = np.clip(((ys - y_min) / (y_max - y_min) * (H - 1)).astype(int), 0, H - 1)  # :40
    cell = row * W + col
    first = np.full(H * W, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(first, cell, idx)
    hit = first != np.iinfo(np.int64).max
    mask[hit] = labels[first[hit]]
    return mask.reshape(H, W)


def range_cells(points: np.ndarray, grid_size=(64, 512), fov_deg=(3.0, -25.0)):
    """Range-image cell per point in FLOAT64 (the spherical projection is not in the reference: parity unpinned; this is
    the written convention of include/kdfusion_b200.h evaluated exactly enough to judge the fp32 device version).

    points [..., >=3] -> (cell int32 [...] with -1 invalid, margin float64 [...]): ``margin`` is the distance of the
    point's fractional image coordinates (and of its pitch from the field-of-view limits, in rows) to the nearest cell
    boundary -- the fp32 kernel may legitimately differ only where it is tiny."""
    H, W = grid_size
    p = np.asarray(points, dtype=np.float32).astype(np.float64)
    x, y, z = p[..., 0], p[..., 1], p[..., 2]
    up, down = np.radians(np.float32(fov_deg[0]).astype(np.float64)), np.radians(np.float32(fov_deg[1]).astype(np.float64))
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        depth = np.sqrt(x * x + y * y + z * z)
        ok = np.isfinite(depth) & (depth > 0)
        pitch = np.arcsin(np.clip(z / np.where(ok, depth, 1.0), -1.0, 1.0))
        t = (pitch - down) / (up - down)
        ok &= (t >= 0) & (t <= 1)
        u = 0.5 * (-np.arctan2(y, x) / np.pi + 1.0) * W
        v = (1.0 - t) * H
    col = np.clip(np.floor(np.where(ok, u, 0.0)), 0, W - 1).astype(np.int64)
    row = np.clip(np.floor(np.where(ok, v, 0.0)), 0, H - 1).astype(np.int64)
    cell = np.where(ok, row * W + col, -1).astype(np.int32)
    frac = lambda a: np.minimum(a - np.floor(a), np.ceil(a) - a)
    with np.errstate(invalid="ignore"):
        margin = np.minimum(np.minimum(frac(u), frac(v)), np.minimum(np.abs(t), np.abs(1.0 - t)) * H)
    return cell, np.where(np.isfinite(margin), margin, 0.0)
