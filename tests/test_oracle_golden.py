"""CPU oracle vs the committed golden fixtures (outputs of the reference itself,
made by tests/golden/make_golden.py).  Runs anywhere, no GPU, no /root/reference."""
import os

import numpy as np
import pytest
import torch

from oracle import bev_oracle, kd_oracle, model_oracle
from oracle.weights import make_state_dict, synthetic_frames

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STRIDE = 97


def _sample(t):
    return t.detach().contiguous().reshape(-1)[::STRIDE].numpy()


@pytest.mark.parametrize("name", ["g64", "g128", "g48x80f"])
def test_bev_cells_golden(name):
    z = np.load(os.path.join(GOLDEN, "bev_cells.npz"))
    seed, B, N, H, W = (int(v) for v in z[name + "_meta"])
    rng = [float(v) for v in z[name + "_range"]]
    if all(float(v).is_integer() for v in rng):
        rng = [int(v) for v in rng]           # int ranges -> int64 buffers in the reference
    _, pts, _ = synthetic_frames(seed, B, N, image_hw=(8, 8), grid_size=(H, W), edge_cases=True)
    cell = bev_oracle.bev_cells(pts.numpy(), (H, W), rng)
    np.testing.assert_array_equal(cell, z[name + "_cell"].astype(np.int32))
    np.testing.assert_array_equal(bev_oracle.bev_occupancy(cell, (H, W)), z[name + "_occ"].astype(np.int32))


@pytest.mark.parametrize("fusion_type", ["weighted", "concat", "minimal"])
@pytest.mark.parametrize("train", [True, False])
def test_model_golden(fusion_type, train):
    z = np.load(os.path.join(GOLDEN, f"model_{fusion_type}_{'train' if train else 'eval'}.npz"))
    wseed, fseed, B, N, ih, iw = (int(v) for v in z["meta"])
    sd = make_state_dict(wseed, fusion_type=fusion_type, num_classes=2, random_running_stats=not train)
    sd = model_oracle.clone_state(sd, requires_grad=train)
    img, pts, lab = synthetic_frames(fseed, B, N, image_hw=(ih, iw), edge_cases=True, nonfinite=False)
    logits, mid = model_oracle.model_forward(img, pts, sd, fusion_type=fusion_type, train=train)
    # same ATen ops in the same order on the same machine class: tight tolerance
    np.testing.assert_allclose(logits.detach().numpy(), z["logits"], rtol=2e-5, atol=2e-6)
    for k in ("camera_feat", "lidar_feat", "pre_fusion", "post_fusion"):
        np.testing.assert_allclose(_sample(mid[k]), z["sample_" + k], rtol=2e-5, atol=2e-6, err_msg=k)
        assert mid[k].double().abs().sum().item() == pytest.approx(z["sum_" + k][1], rel=1e-6)
    if train:
        loss = kd_oracle.ce_loss(logits, lab, torch.tensor([0.4, 3.5]))
        assert loss.item() == pytest.approx(float(z["loss"]), rel=1e-6)
        loss.backward()
        for k in z.files:
            if k.startswith("grad_"):
                g = sd[k[5:]].grad.numpy()
                ref = z[k]
                scale = np.abs(ref).max() + 1e-12
                np.testing.assert_allclose(g / scale, ref / scale, rtol=0, atol=2e-5, err_msg=k)
        np.testing.assert_allclose(sd["lidar_encoder.encoder.point_mlp.7.running_mean"].detach().numpy(),
                                   z["bn_running_mean_lidar7"], rtol=1e-5, atol=1e-7)


def test_kd_oracle_self_consistency():
    """KL term: zero when teacher == student, positive otherwise; per-pixel mean
    normalisation; loss composition."""
    g = torch.Generator().manual_seed(1)
    zs = torch.randn(2, 2, 64, 64, generator=g)
    zt = torch.randn(2, 2, 64, 64, generator=g)
    lab = (torch.rand(2, 64, 64, generator=g) < 0.13).long()
    w = torch.tensor([0.4, 3.5])
    same = kd_oracle.kd_loss(zs, zs, lab, w)
    assert abs(same["kl"].item()) < 1e-7
    out = kd_oracle.kd_loss(zs, zt, lab, w, [zs * 2], [zt], T=4.0, alpha=0.5, beta=1.0)
    assert out["kl"].item() > 0
    # manual per-pixel KL
    T = 4.0
    p = torch.softmax(zt / T, 1)
    q = torch.log_softmax(zs / T, 1)
    kl = (p * (p.log() - q)).sum(1).mean() * T * T
    assert out["kl"].item() == pytest.approx(kl.item(), rel=1e-5)
    assert out["mse"].item() == pytest.approx(((zs * 2 - zt) ** 2).mean().item(), rel=1e-6)
    assert out["loss"].item() == pytest.approx(0.5 * out["ce"].item() + 0.5 * out["kl"].item() + out["mse"].item(), rel=1e-6)


def test_scatter_mean_matches_torch_scatter_reduce():
    _, pts, _ = synthetic_frames(3, 2, 3000, image_hw=(8, 8), edge_cases=True)
    cell = bev_oracle.bev_cells(pts.numpy(), (64, 64))
    g = np.random.default_rng(0)
    feats = g.standard_normal((2, 3000, 16)).astype(np.float32)
    got = bev_oracle.bev_scatter_mean(feats, cell, (64, 64))
    out = torch.zeros(2 * 4096, 16)
    valid = torch.from_numpy(cell >= 0)
    flat = torch.from_numpy(cell.astype(np.int64)) + torch.arange(2).view(2, 1) * 4096
    out.scatter_reduce_(0, flat[valid].unsqueeze(1).expand(-1, 16), torch.from_numpy(feats)[valid],
                        reduce="mean", include_self=False)
    np.testing.assert_allclose(got.reshape(-1, 16), out.numpy(), rtol=1e-5, atol=1e-6)


def test_rasterize_oracle_golden():
    """oracle.bev_oracle.rasterize_bev against outputs of the reference's rasterize_bev
    (src/data_loading/pandaset_dataset.py:23-45; tests/golden/raster_labels.npz), and the product's host rasteriser."""
    from oracle.weights import RASTER_CASES, raster_inputs
    from src.data_loading.pandaset_dataset import rasterize_bev
    z = np.load(os.path.join(GOLDEN, "raster_labels.npz"))
    for name, seed, N, alphabet, grid, rng in RASTER_CASES:
        x, y, labels = raster_inputs(seed, N, alphabet)
        want = z[name].astype(np.int64)
        np.testing.assert_array_equal(bev_oracle.rasterize_bev(x, y, labels, grid, rng), want, err_msg=name)
        np.testing.assert_array_equal(rasterize_bev(x, y, labels, grid, rng), want, err_msg=name)
    empty = bev_oracle.rasterize_bev(np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.int64))
    assert empty.shape == (64, 64) and empty.sum() == 0
