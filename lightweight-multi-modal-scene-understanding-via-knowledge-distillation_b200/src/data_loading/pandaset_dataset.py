"""PandaSet loader with the reference's interface (src/data_loading/pandaset_dataset.py):
``remap_semantic``, ``rasterize_bev``, ``PandaSetDataset`` and
``create_pandaset_dataloaders`` yielding {"image","points","segmentation","sample_token"}.

The loader itself is outside the accelerated path (it needs the real dataset and
pandas/PIL on the host); what matters here is that it feeds the same batch
contract as ``synthetic_frames``.  The BEV label rasterisation, a Python loop over
~100k points per frame in the reference (:42-44), has a device version built on the
same index arithmetic: ``rasterize_bev_cuda`` (``kdf_bev_rasterize``).
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

DRIVABLE_IDS = (6, 7, 8, 9, 10, 12)      # ground, road, lane / stop / other markings, driveway (:13)


def remap_semantic(raw_ids: np.ndarray) -> np.ndarray:
    """PandaSet class ids -> {0 background, 1 drivable} (:15-20)."""
    return np.isin(raw_ids, DRIVABLE_IDS).astype(np.int64)


def rasterize_bev(x: np.ndarray, y: np.ndarray, labels: np.ndarray, grid_size: Tuple[int, int] = (64, 64),
                  pc_range: Tuple[float, float, float, float] = (-50, 50, -50, 50)) -> np.ndarray:
    """Host rasteriser with the reference's semantics (:23-45): a cell takes the first
    non-zero label that lands in it -- for {0,1} labels that is the per-cell maximum,
    which is what this vectorised form computes."""
    H, W = grid_size
    x_min, x_max, y_min, y_max = pc_range
    mask = np.zeros((H, W), dtype=np.int64)
    inside = (x >= x_min) & (x <= x_max) & (y >= y_min) & (y <= y_max)
    if not inside.any():
        return mask
    xs, ys, ls = x[inside], y[inside], labels[inside]
    col = np.clip(((xs - x_min) / (x_max - x_min) * (W - 1)).astype(int), 0, W - 1)
    row = np.clip(((ys - y_min) / (y_max - y_min) * (H - 1)).astype(int), 0, H - 1)
    if ls.max(initial=0) <= 1 and ls.min(initial=0) >= 0:
        np.maximum.at(mask, (row, col), ls.astype(np.int64))
    else:                                   # general labels: keep "first non-zero wins"
        for r, c, lab in zip(row, col, ls):
            if mask[r, c] == 0:
                mask[r, c] = lab
    return mask


def rasterize_bev_cuda(points: torch.Tensor, labels: torch.Tensor, grid_size: Tuple[int, int] = (64, 64),
                       pc_range: Tuple[float, float, float, float] = (-50, 50, -50, 50)) -> torch.Tensor:
    """Device rasteriser with exactly ``rasterize_bev``'s semantics (reference :23-45) for a whole batch:
    points f32[B,N,>=2] (x, y first), labels int[B,N] (any alphabet) -> int64[B,H,W].  One call into
    ``kdf_bev_rasterize``: cell ids by the reference's fp32 formula, "first non-zero label in point order wins"
    as a per-cell integer min-reduction over point indices.  ``pc_range`` = (x_min, x_max, y_min, y_max) like the
    host function."""
    from ..native import call, ptr, require_cuda, stream_ptr
    dev = require_cuda(points, labels)
    if points.dim() != 3 or points.shape[-1] < 2 or points.dtype != torch.float32:
        raise ValueError(f"points must be float32 [B, N, >=2], got {points.dtype} {tuple(points.shape)}")
    B, N, D = points.shape
    if tuple(labels.shape) != (B, N):
        raise ValueError(f"labels must be [B, N] = {(B, N)}, got {tuple(labels.shape)}")
    points, labels = points.contiguous(), labels.long().contiguous()
    H, W = grid_size
    x_min, x_max, y_min, y_max = pc_range
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))
    # the reference subtracts Python scalars: an integral range gives an exact integer span before it meets fp32
    xspan = f32(x_max - x_min) if isinstance(x_min, int) and isinstance(x_max, int) else f32(f32(x_max) - f32(x_min))
    yspan = f32(y_max - y_min) if isinstance(y_min, int) and isinstance(y_max, int) else f32(f32(y_max) - f32(y_min))
    first = torch.empty(B, H * W, dtype=torch.int32, device=dev)
    out = torch.empty(B, H, W, dtype=torch.int64, device=dev)
    call("kdf_bev_rasterize", ptr(points), D, ptr(labels), B, N, f32(x_min), f32(x_max), f32(y_min), f32(y_max),
         xspan, yspan, H, W, ptr(first), ptr(out), stream_ptr(dev))
    return out


class PandaSetDataset(Dataset):
    """2-class PandaSet frames: front camera jpg, lidar + semseg pickles (:48-141)."""

    def __init__(self, root: str, scene_ids: List[str], image_size=(256, 256), grid_size=(64, 64),
                 max_points: int = 5000, verbose: bool = True):
        self.root, self.scene_ids = root, scene_ids
        self.image_size, self.grid_size, self.max_points = image_size, grid_size, max_points
        self.pc_range = (-50, 50, -50, 50)
        self.samples = []
        for sid in scene_ids:
            dirs = {k: os.path.join(root, sid, *sub) for k, sub in
                    (("image", ("camera", "front_camera")), ("lidar", ("lidar",)), ("semseg", ("annotations", "semseg")))}
            if not all(os.path.isdir(d) for d in dirs.values()):
                continue
            frames = sorted(f[:-4] for f in os.listdir(dirs["image"]) if f.endswith(".jpg"))
            usable = 0
            for fid in frames:
                paths = {"image": os.path.join(dirs["image"], fid + ".jpg"),
                         "lidar": os.path.join(dirs["lidar"], fid + ".pkl"),
                         "semseg": os.path.join(dirs["semseg"], fid + ".pkl")}
                if all(os.path.exists(p) for p in paths.values()):
                    self.samples.append({"scene": sid, "frame": fid, **paths})
                    usable += 1
            if verbose:
                print(f"Scene {sid}: {usable}/{len(frames)} frames usable")
        if verbose:
            print(f"Indexed {len(self.samples)} valid samples from {len(scene_ids)} scenes")

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, idx: int) -> Dict[str, torch.Tensor]:
        import pandas as pd
        from PIL import Image
        s = self.samples[idx]
        img = Image.open(s["image"]).convert("RGB").resize(self.image_size, Image.BILINEAR)
        img_t = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1).contiguous()
        df = pd.read_pickle(s["lidar"])
        xyz_i = [df[k].to_numpy(dtype=np.float32) for k in ("x", "y", "z", "i")]
        pts = np.stack(xyz_i, axis=1)
        n = pts.shape[0]
        if n > self.max_points:
            pts = pts[np.random.choice(n, self.max_points, replace=False)]
        elif n < self.max_points:
            pts = np.vstack([pts, np.zeros((self.max_points - n, 4), dtype=np.float32)])
        ids = remap_semantic(pd.read_pickle(s["semseg"])["class"].to_numpy(dtype=np.int64))
        bev = rasterize_bev(xyz_i[0], xyz_i[1], ids, grid_size=self.grid_size, pc_range=self.pc_range)
        return {"image": img_t, "points": torch.from_numpy(pts).contiguous(),
                "segmentation": torch.from_numpy(bev.astype(np.int64)),
                "sample_token": f"{s['scene']}_{s['frame']}"}


def create_pandaset_dataloaders(root: str, train_scenes: List[str], val_scenes: List[str], batch_size: int = 4,
                                num_workers: int = 0, verbose: bool = True, distributed: bool = False):
    """Same call as the reference (:144-157); ``distributed`` shards both datasets over the ranks with a
    DistributedSampler (the Trainer advances its epoch) instead of every rank iterating everything."""
    train_ds = PandaSetDataset(root, train_scenes, verbose=verbose)
    val_ds = PandaSetDataset(root, val_scenes, verbose=verbose)
    ts = vs = None
    if distributed:
        from torch.utils.data.distributed import DistributedSampler
        ts, vs = DistributedSampler(train_ds, shuffle=True), DistributedSampler(val_ds, shuffle=False)
    pin = torch.cuda.is_available()
    return (DataLoader(train_ds, batch_size=batch_size, shuffle=ts is None, sampler=ts, num_workers=num_workers, pin_memory=pin),
            DataLoader(val_ds, batch_size=batch_size, shuffle=False, sampler=vs, num_workers=num_workers, pin_memory=pin))
