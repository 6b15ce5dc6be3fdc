"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

Inputs and weights are regenerated from seeds by ``oracle.weights`` (numpy
PCG64, torch-version independent), so the fixtures hold only the reference's
outputs: bit-exact cell indices / occupancy, logits, the CE loss, sampled
intermediates and a few parameter gradients.  The fixtures travel to the GPU
box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))          # tests/
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))  # repo root
from conftest import build_reference_model, load_reference_module  # noqa: E402
from oracle.weights import RASTER_CASES, make_state_dict, raster_inputs, synthetic_frames  # noqa: E402

SAMPLE_STRIDE = 97
GRAD_KEYS = ["head.cls.weight", "head.cls.bias", "lidar_encoder.encoder.point_mlp.0.weight",
             "lidar_encoder.encoder.point_mlp.6.bias", "camera_encoder.stem.0.weight",
             "camera_fpn.post.net.3.weight"]
FUSION_GRAD_KEYS = {"weighted": ["fusion.attention.0.weight", "fusion.attention.2.bias", "fusion.cam_proj.conv.0.weight"],
                    "concat": ["fusion.fuse.0.weight", "fusion.fuse.3.weight", "fusion.camera_proj.conv.1.weight"],
                    "minimal": ["fusion.lidar_proj.conv.0.weight"]}


def sample(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().reshape(-1)[::SAMPLE_STRIDE].numpy().copy()


def golden_bev():
    le = load_reference_module("models/lidar_encoder")
    out = {}
    for name, grid, rng, seed, B, N in [("g64", (64, 64), [-50, -50, -5, 50, 50, 3], 7, 3, 20000),
                                        ("g128", (128, 128), [-50, -50, -5, 50, 50, 3], 8, 2, 30000),
                                        ("g48x80f", (48, 80), [-40.5, -30.25, -5, 40.5, 61.0, 3], 9, 2, 10000)]:
        enc = le.SpatialLiDAREncoder(grid_size=grid, point_cloud_range=rng)
        _, pts, _ = synthetic_frames(seed, B, N, image_hw=(8, 8), grid_size=grid, edge_cases=True)
        H, W = grid
        coords, valid = enc.points_to_bev_coords(pts)                       # lidar_encoder.py:42-55
        g = (coords * enc.grid_tensor).long()                               # :69
        cell = g[..., 1].clamp(0, H - 1) * W + g[..., 0].clamp(0, W - 1)    # :70-71,77-79
        cell = torch.where(valid, cell, torch.full_like(cell, -1)).to(torch.int32).numpy()
        occ = np.stack([np.bincount(c[c >= 0], minlength=H * W) for c in cell]).astype(np.int32)
        out[name + "_cell"] = cell.astype(np.int16 if H * W < 32768 else np.int32)
        out[name + "_occ"] = occ.astype(np.int16)
        out[name + "_meta"] = np.array([seed, B, N, H, W], dtype=np.int64)
        out[name + "_range"] = np.array(rng, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "bev_cells.npz"), **out)


def golden_raster():
    ds = load_reference_module("data_loading/pandaset_dataset")
    out = {}
    for name, seed, N, alphabet, grid, rng in RASTER_CASES:
        x, y, labels = raster_inputs(seed, N, alphabet)
        out[name] = ds.rasterize_bev(x, y, labels, grid_size=grid, pc_range=rng).astype(np.int8)   # pandaset_dataset.py:23-45
    np.savez_compressed(os.path.join(HERE, "raster_labels.npz"), **out)


def golden_model(fusion_type, train):
    torch.manual_seed(0)
    model = build_reference_model(fusion_type, 2)
    sd = make_state_dict(5, fusion_type=fusion_type, num_classes=2, random_running_stats=not train)
    model.load_state_dict(sd)
    model.train(train)
    img, pts, lab = synthetic_frames(21, 2, 4000, image_hw=(256, 256), edge_cases=True, nonfinite=False)
    crit = torch.nn.CrossEntropyLoss(ignore_index=-1, weight=torch.tensor([0.4, 3.5]))   # trainer.py:55
    logits, mid = model(img, pts, return_intermediates=True)
    out = {"logits": logits.detach().numpy(), "meta": np.array([5, 21, 2, 4000, 256, 256], dtype=np.int64)}
    for k in ("camera_feat", "lidar_feat", "pre_fusion", "post_fusion"):
        out["sample_" + k] = sample(mid[k])
        out["sum_" + k] = np.array([mid[k].double().sum().item(), mid[k].double().abs().sum().item()])
    if train:
        loss = crit(logits, lab)
        loss.backward()
        out["loss"] = np.array(loss.item(), dtype=np.float64)
        named = dict(model.named_parameters())
        for k in GRAD_KEYS + FUSION_GRAD_KEYS[fusion_type]:
            out["grad_" + k] = named[k].grad.numpy()
        out["bn_running_mean_lidar7"] = model.state_dict()["lidar_encoder.encoder.point_mlp.7.running_mean"].numpy()
        out["bn_running_var_fpn_post4"] = model.state_dict()["camera_fpn.post.net.4.running_var"].numpy()
    np.savez_compressed(os.path.join(HERE, f"model_{fusion_type}_{'train' if train else 'eval'}.npz"), **out)


if __name__ == "__main__":
    golden_bev()
    golden_raster()
    for ft in ("weighted", "concat", "minimal"):
        for train in (True, False):
            golden_model(ft, train)
    print("golden fixtures written to", HERE)
