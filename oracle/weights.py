"""Deterministic weight filler (TEST INFRASTRUCTURE ONLY).

Generates a reference-format ``state_dict`` from ``model_oracle.state_dict_spec``
using numpy's PCG64 keyed by (seed, crc32(key)), so the values do not depend on
the torch version or on parameter creation order.  The golden fixtures store
only the seed; ``tests/golden/make_golden.py`` loads the same dict into the
reference's modules.
"""
from __future__ import annotations

import zlib
from typing import Dict

import numpy as np
import torch

from .model_oracle import state_dict_spec


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))


def make_state_dict(seed: int = 0, *, random_running_stats: bool = False,
                    grid_size=(64, 64), point_cloud_range=(-50, -50, -5, 50, 50, 3),
                    **spec_kwargs) -> Dict[str, torch.Tensor]:
    """Reference-format state_dict with He-style conv weights, BN gamma in
    [0.5,1.5], BN beta / conv bias in [-0.2,0.2]; running stats are the
    module defaults (0/1) unless ``random_running_stats`` (for eval-mode tests)."""
    spec = state_dict_spec(grid_size=grid_size, point_cloud_range=point_cloud_range, **spec_kwargs)
    H, W = grid_size
    r = point_cloud_range
    sd: Dict[str, torch.Tensor] = {}
    for key, (shape, kind) in spec.items():
        g = _rng(seed, key)
        if kind == "conv":
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
            a = g.standard_normal(shape).astype(np.float32) * np.float32(np.sqrt(2.0 / max(fan_in, 1)))
        elif kind in ("bias", "bn_b"):
            a = g.uniform(-0.2, 0.2, shape).astype(np.float32)
        elif kind == "bn_w":
            a = g.uniform(0.5, 1.5, shape).astype(np.float32)
        elif kind == "bn_rm":
            a = (g.uniform(-0.3, 0.3, shape) if random_running_stats else np.zeros(shape)).astype(np.float32)
        elif kind == "bn_rv":
            a = (g.uniform(0.5, 2.0, shape) if random_running_stats else np.ones(shape)).astype(np.float32)
        elif kind == "bn_n":
            sd[key] = torch.tensor(0, dtype=torch.long)
            continue
        elif kind == "buf":
            if key.endswith("x_range"):
                sd[key] = torch.tensor([r[0], r[3]])           # int64 for int ranges (lidar_encoder.py:38)
            elif key.endswith("y_range"):
                sd[key] = torch.tensor([r[1], r[4]])
            else:
                sd[key] = torch.tensor([W - 1, H - 1], dtype=torch.float32)
            continue
        else:  # pragma: no cover
            raise KeyError(kind)
        sd[key] = torch.from_numpy(np.ascontiguousarray(a))
    return sd


def synthetic_frames(seed: int, B: int, N: int, image_hw=(256, 256), grid_size=(64, 64),
                     edge_cases: bool = False, nonfinite: bool = True):
    """PandaSet-shaped synthetic batch, numpy-seeded:
    image U[0,1), points x,y~N(0,40^2), z~N(-1,2^2), intensity U[0,255],
    labels Bernoulli(0.13).  ``edge_cases`` injects exact +-50, NaN, +-inf,
    zero padding rows and a few ignore labels (``nonfinite=False`` leaves the
    NaN/inf points out: through train-mode BatchNorm a single NaN point poisons
    the whole batch in the reference too, so model-level tests use finite points)."""
    g = _rng(seed, "frames")
    img = g.random((B, 3, *image_hw), dtype=np.float32)
    pts = np.empty((B, N, 4), dtype=np.float32)
    pts[..., 0] = g.standard_normal((B, N), dtype=np.float32) * 40
    pts[..., 1] = g.standard_normal((B, N), dtype=np.float32) * 40
    pts[..., 2] = g.standard_normal((B, N), dtype=np.float32) * 2 - 1
    pts[..., 3] = g.random((B, N), dtype=np.float32) * 255
    lab = (g.random((B, *grid_size)) < 0.13).astype(np.int64)
    if edge_cases and N >= 64:
        pts[:, 0, :2] = (50.0, 50.0)
        pts[:, 1, :2] = (-50.0, -50.0)
        pts[:, 2, :2] = (50.0, -50.0)
        if nonfinite:
            pts[:, 3, 0] = np.nan
            pts[:, 4, 1] = np.inf
            pts[:, 5, 0] = -np.inf
        pts[:, 6, :2] = (-50.0001, 0.0)
        pts[:, 7, :2] = (np.nextafter(np.float32(50), np.float32(100)), 0.0)
        npad = max(1, int(0.03 * N))
        pts[:, -npad:, :] = 0.0                      # dataset zero padding (pandaset_dataset.py:124-126)
        pts[:, 8:12, :] = pts[:, 12:13, :]           # exact duplicates -> positive ties
        lab[:, 0, :3] = -1
    return torch.from_numpy(img), torch.from_numpy(pts), torch.from_numpy(lab)


def raster_inputs(seed, N, alphabet):
    """Labelled points for the rasteriser fixtures: sweep-like x,y plus the edge cases (exact range ends, the
    floats just outside them, NaN / inf, zero padding), labels from ``alphabet`` (0 = background)."""
    g = np.random.default_rng(seed)
    x, y = (g.normal(0, 40, N).astype(np.float32) for _ in range(2))
    edge = np.array([50.0, -50.0, np.nextafter(np.float32(50), np.float32(60)), np.nextafter(np.float32(-50), np.float32(-60)),
                     np.nan, np.inf, -np.inf, 0.0, 49.999996, -49.999996], dtype=np.float32)
    k = edge.size
    x[:k], y[:k] = edge, edge[::-1]
    x[k:2 * k], y[k:2 * k] = edge, 0.0
    x[-64:], y[-64:] = 0.0, 0.0                                        # zero padding rows (all in one cell)
    labels = g.choice(np.asarray(alphabet), size=N).astype(np.int64)
    return x, y, labels


RASTER_CASES = [("bin64", 11, 30000, (0, 1), (64, 64), (-50, 50, -50, 50)),
                ("multi64", 12, 20000, (0, 0, 3, 7, 12), (64, 64), (-50, 50, -50, 50)),
                ("bin48x80f", 13, 15000, (0, 0, 0, 1), (48, 80), (-40.5, 40.5, -30.25, 61.0))]
