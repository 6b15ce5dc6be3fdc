"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol the header
declares, the drop-in modules have the reference's interface / state_dict / error behaviour, the
product refuses to run without CUDA (no fallback), and the N>1 plumbing works over gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import PKG, ROOT
from oracle import model_oracle


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "kdfusion_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kdf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib_path = os.path.join(PKG, "libkdfusion_b200.so")
    if not os.path.exists(lib_path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(lib_path)
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/kdfusion_b200.h but not exported"
    lib.kdf_abi_version.restype = ctypes.c_int
    assert lib.kdf_abi_version() == 1
    # the Python binding table covers the same set
    from src import native
    assert sorted(native.EXPORTED_SYMBOLS) == syms


def test_argument_validation_without_a_gpu():
    """argument errors are reported before anything touches the device."""
    from src import native
    rc = native.lib.kdf_bev_index(None, 1, 10, 1, 0.0, 1.0, 0.0, 1.0, 64, 64, None, None, None, None)
    assert rc == 1 and b"point_stride" in native.lib.kdf_last_error()
    rc = native.lib.kdf_kd_loss_fwd_bwd(None, None, None, None, 1, 99, 10, 0, 4.0, 0.5, 1.0, -1,
                                        None, None, None, 0, None, None, None, 0, 0, 1.0, None, None, None, None)
    assert rc == 1 and b"K=99" in native.lib.kdf_last_error()
    assert native.lib.kdf_bev_workspace_bytes(2, 1000, 64, 64) >= 2 * 4 * 2000 + 4 * 2 * 4097


def _build(ft="weighted", **kw):
    from src.models.camera_encoder import TwinLiteEncoder
    from src.models.fusion_module import CompleteSegmentationModel
    from src.models.lidar_encoder import LiDAREncoder
    return CompleteSegmentationModel(TwinLiteEncoder(return_multiscale=True),
                                     LiDAREncoder("spatial", grid_size=(64, 64)), num_classes=kw.pop("num_classes", 2),
                                     fusion_type=ft, fusion_out_channels=256 if ft == "concat" else 128,
                                     camera_fpn_stages=["stage3", "stage4", "stage5"], camera_fpn_channels=128, **kw)


@pytest.mark.parametrize("ft,total,fusion", [("concat", "573,442", "161,920"), ("minimal", "494,978", "93,056"),
                                             ("weighted", "528,132", "126,210")])
def test_parameter_counts_match_published(ft, total, fusion):
    s = _build(ft).get_architecture_summary()                      # fusion_ablation_results.json:2-16
    assert s["total_params"] == total and s["fusion_params"] == fusion
    assert s["camera_params"] == "363,520" and s["lidar_params"] == "25,792"
    assert s["fusion_type"] == ft and s["output_mode"] == "same" and s["use_multiscale"] is True


@pytest.mark.parametrize("ft", ["concat", "minimal", "weighted"])
@pytest.mark.parametrize("mode,classes", [("same", 2), ("x4", 3)])
def test_state_dict_is_reference_format(ft, mode, classes):
    m = _build(ft, output_mode=mode, num_classes=classes)
    spec = model_oracle.state_dict_spec(fusion_type=ft, output_mode=mode, num_classes=classes)
    sd = m.state_dict()
    assert list(sd.keys()) == list(spec.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == spec[k][0], k
    enc = m.lidar_encoder.encoder
    assert enc.x_range.dtype == torch.int64 and enc.x_range.tolist() == [-50, 50]
    assert enc.grid_tensor.tolist() == [63.0, 63.0]
    if mode == "same":                                       # SURVEY.md section 5 key counts
        assert len(sd) == {"concat": 194, "minimal": 182, "weighted": 186}[ft]


def test_api_surface_and_errors():
    from src.models.camera_encoder import InvertedResidual, TwinLiteEncoder
    from src.models.fusion_module import CompleteSegmentationModel
    from src.models.lidar_encoder import LiDAREncoder, MMDet3D_AVAILABLE, SpatialLiDAREncoder, create_test_point_cloud
    cam = TwinLiteEncoder()
    assert cam.return_multiscale is False and cam.out_channels == 128 and cam.count_parameters() == 363520
    assert cam.get_feature_info() == {"stage2": 64, "stage3": 64, "stage4": 128, "stage5": 128}
    assert InvertedResidual(32, 32).use_residual and not InvertedResidual(32, 64, stride=2).use_residual
    enc = SpatialLiDAREncoder()
    assert enc.grid_size == (128, 128) and enc.feature_dim == 128 and enc.use_vectorized
    assert LiDAREncoder("spatial", grid_size=(64, 64)).get_output_shape() == (128, 64, 64)
    assert LiDAREncoder("pointpillars").encoder_type == "spatial" and MMDet3D_AVAILABLE is False
    with pytest.raises(ValueError, match="Unknown encoder type"):
        LiDAREncoder("voxel")
    with pytest.raises(ValueError, match="Unknown fusion_type"):
        CompleteSegmentationModel(cam, LiDAREncoder(), fusion_type="sum")
    with pytest.raises(ValueError, match="Unknown output_mode"):
        CompleteSegmentationModel(cam, LiDAREncoder(), output_mode="x2")
    # fixture keeps the reference's RNG consumption and value ranges (lidar_encoder.py:227-234)
    torch.manual_seed(123)
    pts = create_test_point_cloud(2, 1500)
    torch.manual_seed(123)
    raw = torch.randn(2, 1500, 4)
    assert torch.equal(pts[..., 0], raw[..., 0] * 40) and torch.equal(pts[..., 2], raw[..., 2] * 4 - 1)
    assert torch.equal(pts[..., 3], torch.sigmoid(raw[..., 3]))
    coords, valid = enc.points_to_bev_coords(pts)
    assert coords.shape == (2, 1500, 2) and valid.sum().item() == 1796          # SURVEY.md section 4 KAT


def test_no_cpu_fallback():
    """the product path must fail loudly off-GPU instead of computing something else."""
    from src import ops
    from src.models.lidar_encoder import create_test_point_cloud
    from src.training.trainer import SegmentationMetrics, Trainer
    m = _build("weighted")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.rand(1, 3, 256, 256), create_test_point_cloud(1, 100))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.bev_index(torch.zeros(1, 4, 4), (-50.0, 100.0, -50.0, 100.0), (64, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.kd_loss_fwd_bwd(torch.zeros(1, 2, 4, 4), None, torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        SegmentationMetrics().update(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        Trainer(m, [], [], "cpu", save_dir="/tmp/kdf_never")
    # nothing in the product imports the oracle
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_range_constants_promotion():
    from oracle import bev_oracle
    from src import ops
    for rng in ([-50, -50, -5, 50, 50, 3], [-40.5, -30.25, -5, 40.5, 61.0, 3], [0, -7, 0, 33, 9.5, 1]):
        want = tuple(float(v) for v in bev_oracle.range_constants(rng))
        assert ops.bev_range_constants(rng) == want


def test_host_rasterizer_matches_reference_semantics():
    from src.data_loading.pandaset_dataset import rasterize_bev, remap_semantic
    g = np.random.default_rng(0)
    x, y = g.normal(0, 40, 5000).astype(np.float32), g.normal(0, 40, 5000).astype(np.float32)
    raw = g.integers(0, 43, 5000)
    lab = remap_semantic(raw)
    assert set(np.unique(lab)) <= {0, 1} and lab.sum() == np.isin(raw, [6, 7, 8, 9, 10, 12]).sum()
    got = rasterize_bev(x, y, lab)
    # literal restatement of the reference loop (pandaset_dataset.py:32-45)
    want = np.zeros((64, 64), np.int64)
    m = (x >= -50) & (x <= 50) & (y >= -50) & (y <= 50)
    col = np.clip(((x[m] + 50) / 100 * 63).astype(int), 0, 63)
    row = np.clip(((y[m] + 50) / 100 * 63).astype(int), 0, 63)
    for r, c, l in zip(row, col, lab[m]):
        if want[r, c] == 0:
            want[r, c] = l
    np.testing.assert_array_equal(got, want)


def test_batch_counters_deferred_into_one_add():
    """nn.BatchNorm advances num_batches_tracked once per training forward; inside the Trainer's step the 29 increments
    are collected and applied as one multi-tensor add (src/native.py).  Same counts either way; a BatchNorm with
    momentum=None (cumulative average) needs the counter at once and is never deferred."""
    import torch.nn as nn
    from src import native
    bns = [nn.BatchNorm2d(4) for _ in range(3)]
    cum = nn.BatchNorm2d(4, momentum=None)
    assert native.bump_batch_counter(bns[0]) == pytest.approx(0.1) and bns[0].num_batches_tracked.item() == 1
    with native.deferred_batch_counters():
        for b in bns:
            assert native.bump_batch_counter(b) == pytest.approx(0.1)
        assert [b.num_batches_tracked.item() for b in bns] == [1, 0, 0]          # nothing applied yet
        assert native.bump_batch_counter(cum) == pytest.approx(1.0) and cum.num_batches_tracked.item() == 1
        with native.deferred_batch_counters():                                   # nests
            native.bump_batch_counter(bns[1])
        assert bns[1].num_batches_tracked.item() == 1
    assert [b.num_batches_tracked.item() for b in bns] == [2, 2, 1]
    assert native.bump_batch_counter(cum) == pytest.approx(0.5)


def test_microbench_torch_scatter_baseline_is_the_reference_op():
    """tools/projection_microbench.py times `torch_scatter_projection` as "the reference's PyTorch scatter": it must be
    the reference's statements (lidar_encoder.py:42-55, 69-99) -- checked here on the CPU against the oracle (which is
    pinned bit-exact against the reference) on frames with edge cases: exact +-50, NaN / inf, zero-padding rows, ties."""
    import importlib.util
    from oracle import bev_oracle
    from oracle.weights import synthetic_frames
    spec = importlib.util.spec_from_file_location("_pmb", os.path.join(ROOT, "tools", "projection_microbench.py"))
    pmb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pmb)
    B, N, C = 2, 6000, 16
    _, pts, _ = synthetic_frames(17, B, N, image_hw=(8, 8), edge_cases=True)
    g = np.random.default_rng(3)
    feats = np.maximum(g.standard_normal((B, N, C)).astype(np.float32), 0)
    feats[:, 8:12] = feats[:, 12:13]                                            # ties
    got = pmb.torch_scatter_projection(pts, torch.from_numpy(feats), 64, 64)     # [B,H,W,C]
    cell = bev_oracle.bev_cells(pts.numpy(), (64, 64))
    ref, _ = bev_oracle.bev_scatter_max(feats, cell, (64, 64))
    np.testing.assert_array_equal(got.reshape(B, 4096, C).numpy(), ref)


def test_shard_range_and_seeds():
    from src.training.parallel import frame_seed, shard_range
    for n, w in ((64, 8), (10, 4), (3, 8), (256, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert frame_seed(3, 7) == 3007


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, {pkg!r})
import torch, torch.distributed as dist
from src.training.parallel import allreduce_gradients_, reduce_max, reduce_sum, shard_range, world
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, ws = world()
assert ws == 2
# flat-bucket gradient all-reduce; the mean is applied as grad_scale = 1/world in the optimizer
g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
allreduce_gradients_(g)
assert torch.equal(g, torch.arange(1000, dtype=torch.float32) * 3)
assert torch.allclose(g * (1.0 / ws), torch.arange(1000, dtype=torch.float32) * 1.5)
# step time = max over ranks; frames = sum over ranks (weak scaling)
assert reduce_max(10.0 + rank) == 11.0
a, b = shard_range(64, rank, ws)
assert reduce_sum(b - a) == 64.0
# int64 confusion matrices add up
c = torch.tensor([[rank + 1, 2], [3, 4]], dtype=torch.int64)
dist.all_reduce(c)
assert c.tolist() == [[3, 4], [6, 8]]
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(pkg=PKG, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in o, o


def test_eval_caches_follow_the_parameter_generation():
    """Kernels move parameters and running statistics through raw pointers (no ``_version`` bump), so the eval-mode
    BatchNorm affine cache also keys on src.native's generation counter: every optimizer / training step and every
    running-statistics update invalidates it -- except for modules marked frozen (the teacher)."""
    import torch.nn as nn
    from src import native, ops
    bn = nn.BatchNorm2d(4).eval()
    a = ops._eval_affine(bn)
    assert ops._eval_affine(bn) is a                                   # cached
    with torch.no_grad():
        bn.running_var.data_ptr()                                      # a raw-pointer write looks like this to torch:
        bn.running_var.view(-1).numpy()[:] = 4.0                       # memory changes, _version does not
    assert ops._eval_affine(bn) is a                                   # ... and nothing else noticed either
    native.bump_generation()                                           # what training_step / FlatAdamW.step / stats updates do
    b = ops._eval_affine(bn)
    assert b is not a and torch.allclose(b[3], torch.full((4,), (4.0 + bn.eps) ** -0.5))
    native.mark_frozen(bn)
    c = ops._eval_affine(bn)
    native.bump_generation()
    assert ops._eval_affine(bn) is c                                   # frozen modules keep their cache
    g0 = native._generation
    native.bump_batch_counter(nn.BatchNorm2d(4))
    assert native._generation == g0 + 1


def test_activation_commutes_with_bf16_rounding():
    """The kernels take ReLU / ReLU6 AFTER the conversion to bf16, on the packed pair (`max/min.bf16x2`), where the layer-by-
    layer form clamps in fp32 and then rounds.  Both agree exactly because 0 and 6 are bf16 numbers and round-to-nearest-even is
    monotonic: clamp(round(x)) == round(clamp(x)) for every fp32 x (checked on a dense sample around the thresholds, the
    rounding boundaries next to them, and random values)."""
    import torch
    g = torch.Generator().manual_seed(0)
    near = torch.cat([torch.linspace(-1e-2, 1e-2, 20001), 6 + torch.linspace(-5e-2, 5e-2, 20001),
                      torch.randn(200000, generator=g) * 4, torch.tensor([0.0, -0.0, 6.0, 5.984375, 6.03125, float("inf"), -float("inf")])])
    for lo, hi in ((0.0, float("inf")), (0.0, 6.0)):
        a = near.clamp(lo, hi).to(torch.bfloat16)
        b = near.to(torch.bfloat16).clamp(lo, hi)
        assert torch.equal(a.float(), b.float())                     # (value equality: +0 and -0 are the same activation)
