"""Seeded synthetic PandaSet-shaped frames.

Follows the output contract of the reference's ``PandaSetDataset.__getitem__``
(src/data_loading/pandaset_dataset.py:108-141): a dict with
``image`` f32[3,256,256] in [0,1), ``points`` f32[N,4] = (x, y, z, raw intensity
0..255), ``segmentation`` int64[64,64] in {0,1} (a few -1 on request) and
``sample_token``.  x,y ~ N(0, 40^2) so ~62 % of a sweep falls inside +-50 m,
z ~ N(-1, 2^2), labels Bernoulli(0.13) (drivable share, train_pandaset.py:135).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
from torch.utils.data import DataLoader, Dataset


def make_frames(batch: int, num_points: int = 170_000, image_size: Tuple[int, int] = (256, 256),
                grid_size: Tuple[int, int] = (64, 64), seed: int = 0, device="cpu",
                ignore_fraction: float = 0.0) -> Dict[str, torch.Tensor]:
    """One batch, generated directly on ``device`` from ``seed`` (1000*rank + step in the benches)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    img = torch.rand(batch, 3, *image_size, generator=g, device=device)
    pts = torch.randn(batch, num_points, 4, generator=g, device=device)
    pts[..., 0:2] *= 40.0
    pts[..., 2] = pts[..., 2] * 2.0 - 1.0
    pts[..., 3] = torch.rand(batch, num_points, generator=g, device=device) * 255.0
    u = torch.rand(batch, *grid_size, generator=g, device=device)
    seg = (u < 0.13).long()
    if ignore_fraction > 0:
        seg[u > 1.0 - ignore_fraction] = -1
    return {"image": img, "points": pts, "segmentation": seg,
            "sample_token": [f"synthetic_{seed}_{i}" for i in range(batch)]}


class SyntheticPandaSetFrames(Dataset):
    """Map-style dataset of deterministic frames; item i depends only on (seed, i)."""

    def __init__(self, num_samples: int = 64, num_points: int = 5000, image_size=(256, 256), grid_size=(64, 64),
                 seed: int = 0):
        self.num_samples, self.num_points = num_samples, num_points
        self.image_size, self.grid_size, self.seed = image_size, grid_size, seed

    def __len__(self):
        return self.num_samples

    def __getitem__(self, idx):
        f = make_frames(1, self.num_points, self.image_size, self.grid_size, seed=self.seed * 1_000_003 + idx)
        return {"image": f["image"][0], "points": f["points"][0], "segmentation": f["segmentation"][0],
                "sample_token": f"synthetic_{self.seed}_{idx}"}


def create_synthetic_dataloaders(num_train: int = 64, num_val: int = 16, batch_size: int = 4, num_workers: int = 0,
                                 num_points: int = 5000, seed: int = 0, distributed: bool = False):
    """Two DataLoaders with the batch-dict contract of ``create_pandaset_dataloaders``
    (pandaset_dataset.py:144-157); pinned memory so the H2D copies are asynchronous."""
    tr = SyntheticPandaSetFrames(num_train, num_points, seed=seed)
    va = SyntheticPandaSetFrames(num_val, num_points, seed=seed + 1)
    ts = vs = None
    if distributed:
        from torch.utils.data.distributed import DistributedSampler
        ts, vs = DistributedSampler(tr, shuffle=True, seed=seed), DistributedSampler(va, shuffle=False)
    pin = torch.cuda.is_available()
    return (DataLoader(tr, batch_size=batch_size, shuffle=ts is None, sampler=ts, num_workers=num_workers, pin_memory=pin),
            DataLoader(va, batch_size=batch_size, shuffle=False, sampler=vs, num_workers=num_workers, pin_memory=pin))
