"""Pins the CPU oracle against the reference's own modules.

Only runs where /root/reference is mounted (the build container); the golden
fixtures produced from the same runs (tests/golden) cover the GPU box."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import build_reference_model, load_reference_module
from oracle import bev_oracle, kd_oracle, model_oracle
from oracle.weights import make_state_dict, synthetic_frames

pytestmark = pytest.mark.reference


def _ref_cells(enc, pts):
    """Flat cell ids with the reference's own lines (lidar_encoder.py:63,69-71)."""
    H, W = enc.grid_size
    coords, valid = enc.points_to_bev_coords(pts)
    g = (coords * enc.grid_tensor).long()
    col = g[..., 0].clamp(0, W - 1)
    row = g[..., 1].clamp(0, H - 1)
    cell = row * W + col
    return torch.where(valid, cell, torch.full_like(cell, -1)), coords, valid


def test_known_answer_vector_survey_s4():
    le = load_reference_module("models/lidar_encoder")
    torch.manual_seed(123)
    pts = le.create_test_point_cloud(2, 1500)
    cell = bev_oracle.bev_cells(pts.numpy(), (64, 64))
    valid = cell >= 0
    flat = (np.arange(2)[:, None] * 4096 + cell)[valid].astype(np.int64)
    assert valid.sum() == 1796
    assert flat.sum() == 7278080
    occ = bev_oracle.bev_occupancy(cell, (64, 64))
    assert (occ > 0).sum() == 1594 and occ.max() == 4
    assert hashlib.sha1(flat.tobytes()).hexdigest()[:16] == "9d430af99ab6ca47"


@pytest.mark.parametrize("grid", [(64, 64), (128, 128), (48, 80)])
@pytest.mark.parametrize("rng", [[-50, -50, -5, 50, 50, 3], [-40.5, -30.25, -5, 40.5, 61.0, 3]])
def test_cells_bit_exact_vs_reference(grid, rng):
    le = load_reference_module("models/lidar_encoder")
    enc = le.SpatialLiDAREncoder(grid_size=grid, point_cloud_range=rng)
    _, pts, _ = synthetic_frames(7, 3, 20000, image_hw=(8, 8), edge_cases=True)
    # add points exactly on cell boundaries
    pts[0, 100:164, 0] = torch.linspace(-50, 50, 64)
    ref_cell, ref_coords, ref_valid = _ref_cells(enc, pts)
    coords, valid = bev_oracle.bev_coords(pts.numpy(), rng)
    np.testing.assert_array_equal(valid, ref_valid.numpy())
    np.testing.assert_array_equal(coords[valid], ref_coords.numpy()[valid])
    cell = bev_oracle.bev_cells(pts.numpy(), grid, rng)
    np.testing.assert_array_equal(cell, ref_cell.numpy().astype(np.int32))


def test_scatter_max_and_backward_vs_reference():
    le = load_reference_module("models/lidar_encoder")
    enc = le.SpatialLiDAREncoder(grid_size=(64, 64)).eval()
    _, pts, _ = synthetic_frames(3, 2, 6000, image_hw=(8, 8), edge_cases=True)
    B, N, C = 2, 6000, 128
    g = np.random.default_rng(0)
    feats = torch.from_numpy(np.maximum(g.standard_normal((B, N, C)).astype(np.float32), 0))
    feats[:, 8:12] = feats[:, 12:13]                       # positive ties
    feats.requires_grad_(True)
    # the reference's scatter lines (lidar_encoder.py:74-99) on given features
    H, W = 64, 64
    coords, valid = enc.points_to_bev_coords(pts)
    gc = (coords * enc.grid_tensor).long()
    gc[..., 0] = gc[..., 0].clamp(0, W - 1); gc[..., 1] = gc[..., 1].clamp(0, H - 1)
    bi = torch.arange(B).view(B, 1).expand(B, N)
    flat = bi[valid] * (H * W) + gc[valid][:, 1] * W + gc[valid][:, 0]
    out = torch.zeros(B * H * W, C)
    out.scatter_reduce_(0, flat.unsqueeze(1).expand(-1, C), feats[valid], reduce="amax", include_self=False)
    gg = torch.from_numpy(g.standard_normal((B * H * W, C)).astype(np.float32))
    out.backward(gg)

    cell = bev_oracle.bev_cells(pts.numpy(), (64, 64))
    grid, ties = bev_oracle.bev_scatter_max(feats.detach().numpy(), cell, (64, 64))
    np.testing.assert_array_equal(grid.reshape(B * H * W, C), out.detach().numpy())
    gf = bev_oracle.bev_scatter_max_backward(gg.numpy().reshape(B, H * W, C), feats.detach().numpy(), grid, ties, cell)
    np.testing.assert_allclose(gf, feats.grad.numpy(), rtol=1e-6, atol=0)
    assert ties.max() >= 4


def test_c_oracle_matches_numpy_oracle():
    import ctypes, os, subprocess
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(here, "oracle")])
    lib = ctypes.CDLL(os.path.join(here, "oracle", "_build", "libbev_oracle.so"))
    _, pts, _ = synthetic_frames(11, 2, 5000, image_hw=(8, 8), edge_cases=True)
    p = np.ascontiguousarray(pts.numpy())
    cell = np.empty((2, 5000), dtype=np.int32)
    x0, xs, y0, ys = bev_oracle.range_constants([-50, -50, -5, 50, 50, 3])
    lib.bevo_cells(p.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(p.shape[0] * p.shape[1]), ctypes.c_int(4),
                   ctypes.c_float(x0), ctypes.c_float(xs), ctypes.c_float(y0), ctypes.c_float(ys),
                   ctypes.c_int(64), ctypes.c_int(64), cell.ctypes.data_as(ctypes.c_void_p))
    np.testing.assert_array_equal(cell, bev_oracle.bev_cells(p, (64, 64)))


@pytest.mark.parametrize("fusion_type,num_classes,mode", [("weighted", 2, "same"), ("concat", 2, "same"),
                                                          ("minimal", 2, "same"), ("concat", 3, "x4")])
@pytest.mark.parametrize("train", [True, False])
def test_model_oracle_vs_reference(fusion_type, num_classes, mode, train):
    ref = build_reference_model(fusion_type, num_classes, output_mode=mode)
    sd = make_state_dict(5, fusion_type=fusion_type, num_classes=num_classes, output_mode=mode,
                         random_running_stats=not train)
    assert list(ref.state_dict().keys()) == list(sd.keys())
    for k, v in ref.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape) and v.dtype == sd[k].dtype, k
    ref.load_state_dict(sd)
    ref.train(train)
    img, pts, _ = synthetic_frames(21, 2, 3000, image_hw=(128, 128), edge_cases=True, nonfinite=False)
    with torch.no_grad():
        ref_logits, ref_mid = ref(img, pts, return_intermediates=True)
        sd2 = model_oracle.clone_state(sd)
        logits, mid = model_oracle.model_forward(img, pts, sd2, fusion_type=fusion_type, output_mode=mode, train=train)
    for k in ("camera_feat", "lidar_feat", "pre_fusion", "post_fusion", "logits"):
        assert torch.equal(mid[k], ref_mid[k]), k
    assert mid["lidar_feat"].stride() == ref_mid["lidar_feat"].stride()       # NHWC-strided view
    if train:   # running statistics updated the same way
        for k, v in ref.state_dict().items():
            assert torch.equal(v, sd2[k]), k


def test_param_counts_match_published():
    # fusion_ablation_results.json:2-16
    want = {"concat": 573442, "minimal": 494978, "weighted": 528132}
    for ft, n in want.items():
        spec = model_oracle.state_dict_spec(fusion_type=ft)
        got = sum(int(np.prod(s)) for s, kind in spec.values() if kind in ("conv", "bias", "bn_w", "bn_b"))
        assert got == n


def test_ce_and_metrics_vs_reference_trainer():
    tr = load_reference_module("training/trainer")
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(3, 2, 64, 64, generator=g)
    labels = (torch.rand(3, 64, 64, generator=g) < 0.13).long()
    labels[0, :2] = -1
    w = torch.tensor([0.4, 3.5])
    crit = torch.nn.CrossEntropyLoss(ignore_index=-1, weight=w)         # trainer.py:55
    assert torch.equal(kd_oracle.ce_loss(logits, labels, w), crit(logits, labels))
    m = tr.SegmentationMetrics(num_classes=2)
    m.update(logits, labels)
    conf = kd_oracle.confusion_matrix(logits, labels, 2)
    np.testing.assert_array_equal(conf.numpy(), m.confusion)
    assert kd_oracle.miou(conf)["miou"] == pytest.approx(m.compute()["miou"], abs=1e-12)


def test_rasterize_oracle_vs_reference_loop():
    """The integer-min restatement equals the reference's per-point loop (pandaset_dataset.py:23-45) on sweep-like
    inputs with edge cases, for binary and multi-valued labels and a fractional, non-square geometry."""
    from oracle import bev_oracle
    from oracle.weights import RASTER_CASES, raster_inputs
    ds = load_reference_module("data_loading/pandaset_dataset")
    for name, seed, N, alphabet, grid, rng in RASTER_CASES:
        x, y, labels = raster_inputs(seed + 100, N // 3, alphabet)
        want = ds.rasterize_bev(x, y, labels, grid_size=grid, pc_range=rng)
        np.testing.assert_array_equal(bev_oracle.rasterize_bev(x, y, labels, grid, rng), want, err_msg=name)
