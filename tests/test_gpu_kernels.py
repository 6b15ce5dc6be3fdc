"""GPU parity tests: each CUDA kernel, called through the C ABI (ctypes binding in
src/native.py), against the CPU oracle on the same seeded inputs.

Bars: indices / occupancy / max-grid bit-exact; fp32 losses, fused features and
gradients within 1e-5 relative; bf16 within the tolerance written in each test."""
import os

import numpy as np
import pytest
import torch

from oracle import bev_oracle, kd_oracle, model_oracle
from oracle.weights import make_state_dict, synthetic_frames

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ops():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from src import ops as _ops
    return _ops


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


# ----------------------------------------------------------------------------- projection: indices
@pytest.mark.parametrize("name", ["g64", "g128", "g48x80f"])
def test_bev_index_golden_bit_exact(ops, name):
    """cells + occupancy equal the REFERENCE's own outputs (tests/golden/bev_cells.npz)."""
    z = np.load(os.path.join(GOLDEN, "bev_cells.npz"))
    seed, B, N, H, W = (int(v) for v in z[name + "_meta"])
    rng = [float(v) for v in z[name + "_range"]]
    if all(v.is_integer() for v in rng):
        rng = [int(v) for v in rng]
    _, pts, _ = synthetic_frames(seed, B, N, image_hw=(8, 8), grid_size=(H, W), edge_cases=True)
    cell, occ = ops.bev_index(pts.cuda(), ops.bev_range_constants(rng), (H, W))
    np.testing.assert_array_equal(cell.cpu().numpy(), z[name + "_cell"].astype(np.int32))
    np.testing.assert_array_equal(occ.cpu().numpy(), z[name + "_occ"].astype(np.int32))


def test_bev_index_known_answer_survey_s4(ops):
    """SURVEY.md section 4 known-answer vector (torch CPU RNG, seed 123)."""
    import hashlib
    from src.models.lidar_encoder import create_test_point_cloud
    torch.manual_seed(123)
    pts = create_test_point_cloud(2, 1500)
    cell, occ = ops.bev_index(pts.cuda(), ops.bev_range_constants([-50, -50, -5, 50, 50, 3]), (64, 64))
    cell = cell.cpu().numpy()
    valid = cell >= 0
    flat = (np.arange(2)[:, None] * 4096 + cell)[valid].astype(np.int64)
    assert valid.sum() == 1796 and flat.sum() == 7278080
    assert (occ.cpu().numpy() > 0).sum() == 1594 and occ.max().item() == 4
    assert hashlib.sha1(flat.tobytes()).hexdigest()[:16] == "9d430af99ab6ca47"


@pytest.mark.parametrize("B,N,stride", [(1, 1, 4), (3, 31, 4), (2, 1000, 4), (2, 777, 6), (1, 0, 4), (4, 33, 2)])
def test_bev_index_ragged_sizes_and_strides(ops, B, N, stride):
    g = np.random.default_rng(B * 1000 + N)
    pts = (g.standard_normal((B, N, stride)) * 40).astype(np.float32)
    cell, occ, rank = ops.bev_index(torch.from_numpy(pts).cuda(), ops.bev_range_constants([-50, -50, -5, 50, 50, 3]),
                                    (64, 64), want_rank=True)
    ref = bev_oracle.bev_cells(pts, (64, 64))
    np.testing.assert_array_equal(cell.cpu().numpy(), ref)
    np.testing.assert_array_equal(occ.cpu().numpy(), bev_oracle.bev_occupancy(ref, (64, 64)))
    # ranks: a permutation of 0..count-1 inside every cell
    r, c = rank.cpu().numpy(), cell.cpu().numpy()
    for b in range(B):
        for cid in np.unique(c[b][c[b] >= 0]):
            got = np.sort(r[b][c[b] == cid])
            np.testing.assert_array_equal(got, np.arange(got.size))


def test_bev_index_edge_semantics(ops):
    """closed range, x=+50 -> last cell, NaN/inf/-50.0001 invalid, zero padding -> cell (31,31)
    (SURVEY.md section 7 'Validity and edge semantics')."""
    p = torch.zeros(1, 8, 4)
    p[0, 0, :2] = torch.tensor([50.0, 50.0])
    p[0, 1, :2] = torch.tensor([-50.0, -50.0])
    p[0, 2, 0] = float("nan")
    p[0, 3, 1] = float("inf")
    p[0, 4, :2] = torch.tensor([-50.0001, 0.0])
    p[0, 5, :2] = torch.tensor([49.99, 0.0])
    cell, occ = ops.bev_index(p.cuda(), ops.bev_range_constants([-50, -50, -5, 50, 50, 3]), (64, 64))
    c = cell.cpu().numpy()[0]
    assert c[0] == 63 * 64 + 63 and c[1] == 0 and c[2] == -1 and c[3] == -1 and c[4] == -1
    assert c[5] == 31 * 64 + 62 and c[6] == 31 * 64 + 31 and c[7] == 31 * 64 + 31
    np.testing.assert_array_equal(c, bev_oracle.bev_cells(p.numpy(), (64, 64))[0])
    assert occ.sum().item() == 5


def test_bev_index_full_size_properties(ops):
    """BASELINE size (32 frames x 170k points): size-independent properties."""
    from src.data_loading.synthetic_frames import make_frames
    f = make_frames(32, 170_000, seed=3, device="cuda")
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    cell, occ = ops.bev_index(f["points"], geom, (64, 64))
    assert cell.min().item() >= -1 and cell.max().item() < 4096
    assert occ.sum(dim=1).tolist() == (cell >= 0).sum(dim=1).tolist()            # checksum of checksums
    cell2, occ2 = ops.bev_index(f["points"], geom, (64, 64))
    assert torch.equal(cell, cell2) and torch.equal(occ, occ2)                  # deterministic
    # a sample of frames against the oracle
    ref = bev_oracle.bev_cells(f["points"][:2].cpu().numpy(), (64, 64))
    np.testing.assert_array_equal(cell[:2].cpu().numpy(), ref)
    x, y = f["points"][..., 0], f["points"][..., 1]
    inside = (x >= -50) & (x <= 50) & (y >= -50) & (y <= 50)
    assert torch.equal(inside, cell >= 0)


# ----------------------------------------------------------------------------- projection: reduce + backward
def _proj_inputs(B, N, C, seed, ties=True):
    _, pts, _ = synthetic_frames(seed, B, N, image_hw=(8, 8), edge_cases=N >= 64)
    g = np.random.default_rng(seed)
    feats = np.maximum(g.standard_normal((B, N, C)).astype(np.float32), 0)        # post-ReLU like the MLP output
    if ties and N >= 64:
        feats[:, 8:12] = feats[:, 12:13]
    return pts, torch.from_numpy(feats)


@pytest.mark.parametrize("B,N,C", [(2, 6000, 128), (1, 100, 128), (3, 2500, 64), (2, 4000, 256), (1, 1, 4), (2, 0, 128)])
def test_bev_project_max_fp32_exact(ops, B, N, C):
    pts, feats = _proj_inputs(B, N, C, seed=B * 100 + C)
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    f = feats.cuda().requires_grad_(True)
    grid, count, cell = ops.bev_project(pts.cuda(), f, geom, (64, 64), "max")
    ref_cell = bev_oracle.bev_cells(pts.numpy(), (64, 64))
    ref_grid, ref_ties = bev_oracle.bev_scatter_max(feats.numpy(), ref_cell, (64, 64))
    assert grid.shape == (B, C, 64, 64) and grid.stride() == (64 * 64 * C, 1, 64 * C, C)     # NHWC view (lidar_encoder.py:99)
    np.testing.assert_array_equal(cell.cpu().numpy(), ref_cell)
    np.testing.assert_array_equal(count.cpu().numpy(), bev_oracle.bev_occupancy(ref_cell, (64, 64)))
    np.testing.assert_array_equal(grid.detach().permute(0, 2, 3, 1).reshape(B, 4096, C).cpu().numpy(), ref_grid)
    # backward: even split among ties (+ ATen's zero-max quirk), zero for points outside
    gg = torch.from_numpy(np.random.default_rng(1).standard_normal((B, 4096, C)).astype(np.float32))
    grid.backward(gg.view(B, 64, 64, C).permute(0, 3, 1, 2).cuda())
    ref_gf = bev_oracle.bev_scatter_max_backward(gg.numpy(), feats.numpy(), ref_grid, ref_ties, ref_cell)
    np.testing.assert_allclose(f.grad.cpu().numpy(), ref_gf, rtol=1e-6, atol=0)


def test_bev_project_outside_points_give_zero_grid(ops):
    """the reference's own (commented) assertion: all points out of range => output max == 0
    (test_lidar_encoder.py:227-233)."""
    pts = torch.full((2, 500, 4), 500.0)
    feats = torch.rand(2, 500, 128) + 1
    grid, count, cell = ops.bev_project(pts.cuda(), feats.cuda(), ops.bev_range_constants([-50, -50, -5, 50, 50, 3]), (64, 64))
    assert grid.abs().max().item() == 0.0 and count.sum().item() == 0 and (cell == -1).all()


def test_bev_project_all_in_one_cell_and_negative_features(ops):
    """heavy cell (every point in one cell) and features that are not post-ReLU:
    include_self=False semantics -> max of the sources even when negative."""
    B, N, C = 2, 3000, 128
    pts = torch.zeros(B, N, 4)
    pts[..., 0] = 10.2
    pts[..., 1] = -7.7
    g = np.random.default_rng(5)
    feats = torch.from_numpy(g.standard_normal((B, N, C)).astype(np.float32) - 3.0)
    grid, count, cell = ops.bev_project(pts.cuda(), feats.cuda(), ops.bev_range_constants([-50, -50, -5, 50, 50, 3]), (64, 64))
    ref_cell = bev_oracle.bev_cells(pts.numpy(), (64, 64))
    ref_grid, _ = bev_oracle.bev_scatter_max(feats.numpy(), ref_cell, (64, 64))
    np.testing.assert_array_equal(grid.permute(0, 2, 3, 1).reshape(B, 4096, C).cpu().numpy(), ref_grid)
    assert count.max().item() == N and (ref_grid.min() < 0)


def test_bev_project_bf16_features_exact_max(ops):
    """bf16 features: the max of bf16 values is exactly representable, so the grid must equal
    the oracle run on the bf16-rounded features; points / indices stay fp32."""
    B, N, C = 2, 5000, 128
    pts, feats = _proj_inputs(B, N, C, seed=9)
    fb = feats.to(torch.bfloat16)
    grid, count, cell = ops.bev_project(pts.cuda(), fb.cuda(), ops.bev_range_constants([-50, -50, -5, 50, 50, 3]), (64, 64))
    ref_cell = bev_oracle.bev_cells(pts.numpy(), (64, 64))
    ref_grid, _ = bev_oracle.bev_scatter_max(fb.float().numpy(), ref_cell, (64, 64))
    assert grid.dtype == torch.bfloat16
    np.testing.assert_array_equal(cell.cpu().numpy(), ref_cell)
    np.testing.assert_array_equal(grid.float().permute(0, 2, 3, 1).reshape(B, 4096, C).cpu().numpy(), ref_grid)


@pytest.mark.parametrize("dtype,C", [(torch.float32, 128), (torch.float32, 32), (torch.bfloat16, 128), (torch.bfloat16, 64)])
def test_bev_project_tie_free_forward_equals_tie_counting_path(ops, dtype, C):
    """The pure-maximum forward + tie-counting backward (ties = NULL through the C ABI) against the forward that
    stores tie counts (the earlier kernels, still what other channel counts use): identical grids and gradients,
    bit for bit, including ties, zero maxima (ATen's extra tie) and points outside the grid."""
    from src import native
    B, N, H, W = 3, 7000, 64, 64
    pts, feats = _proj_inputs(B, N, C, seed=31)
    pts, feats = pts.cuda(), feats.to(dtype).cuda()
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    gg = torch.randn(B, H * W, C, device="cuda").to(dtype)
    p, st = native.ptr, native.stream_ptr(pts.device)
    wsb = native.lib.kdf_bev_workspace_bytes(B, N, H, W)
    outs = []
    for with_ties in (True, False):
        grid = torch.empty(B, H, W, C, dtype=dtype, device="cuda")
        cnt = torch.empty(B, H * W, dtype=torch.int32, device="cuda")
        cel = torch.empty(B, N, dtype=torch.int32, device="cuda")
        ties = torch.empty(B, H * W, C, dtype=torch.int32, device="cuda") if with_ties else None
        order = torch.empty(B, N, dtype=torch.int32, device="cuda")
        offs = torch.empty(B, H * W + 1, dtype=torch.int32, device="cuda")
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        gf = torch.full((B, N, C), 7.0, dtype=dtype, device="cuda")
        native.call("kdf_bev_project_fwd", p(pts), 4, p(feats), native.dtype_code(feats), B, N, C, *geom, H, W, 0,
                    p(grid), p(cnt), p(cel), p(ties), p(order), p(offs), p(ws), wsb, st)
        native.call("kdf_bev_project_bwd", p(gg), p(feats), p(grid), p(ties), None, p(cel), p(order), p(offs),
                    native.dtype_code(feats), B, N, C, H, W, 0, p(gf), st)
        outs.append((grid, gf))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert (outs[1][1] != 0).any()


def test_bev_project_mean(ops):
    """per-cell mean (north star; not in the reference -> parity unpinned, torch scatter_reduce as oracle)."""
    B, N, C = 2, 5000, 64
    pts, feats = _proj_inputs(B, N, C, seed=4, ties=False)
    f = feats.cuda().requires_grad_(True)
    grid, count, cell = ops.bev_project(pts.cuda(), f, ops.bev_range_constants([-50, -50, -5, 50, 50, 3]), (64, 64), "mean")
    ref_cell = bev_oracle.bev_cells(pts.numpy(), (64, 64))
    ref = bev_oracle.bev_scatter_mean(feats.numpy(), ref_cell, (64, 64))
    np.testing.assert_allclose(grid.detach().permute(0, 2, 3, 1).reshape(B, 4096, C).cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    gg = torch.randn(B, C, 64, 64, device="cuda")
    grid.backward(gg)
    occ = bev_oracle.bev_occupancy(ref_cell, (64, 64))
    ggr = gg.permute(0, 2, 3, 1).reshape(B, 4096, C).cpu().numpy()
    want = np.zeros((B, N, C), np.float32)
    for b in range(B):
        v = ref_cell[b] >= 0
        want[b, v] = ggr[b, ref_cell[b, v]] / occ[b, ref_cell[b, v]][:, None]
    np.testing.assert_allclose(f.grad.cpu().numpy(), want, rtol=1e-6, atol=1e-7)


def test_bev_project_full_size_properties(ops):
    """32 x 170k x 128 bf16 (the bench shape): order independence (point permutation leaves the
    grid bit-identical), idempotence, occupancy checksum, and max >= any member."""
    from src.data_loading.synthetic_frames import make_frames
    B, N, C = 32, 170_000, 128
    pts = make_frames(B, N, seed=11, device="cuda")["points"]
    feats = torch.rand(B, N, C, device="cuda", dtype=torch.bfloat16)
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    grid, count, cell = ops.bev_project(pts, feats, geom, (64, 64))
    perm = torch.randperm(N, device="cuda")
    grid2, count2, _ = ops.bev_project(pts[:, perm].contiguous(), feats[:, perm].contiguous(), geom, (64, 64))
    assert torch.equal(grid, grid2) and torch.equal(count, count2)
    assert count.sum().item() == (cell >= 0).sum().item()
    # every valid point is <= its cell max; empty cells are exactly zero
    g = grid.permute(0, 2, 3, 1).reshape(B, 4096, C)
    b = 5
    v = cell[b] >= 0
    assert (feats[b][v] <= g[b][cell[b][v].long()]).all()
    assert (g[count == 0] == 0).all()


# ----------------------------------------------------------------------------- KD loss
def _kd_inputs(B, K, seed, feat_shapes=((128, 64, 64), (128, 64, 64))):
    g = torch.Generator().manual_seed(seed)
    zs = torch.randn(B, K, 64, 64, generator=g) * 2
    zt = torch.randn(B, K, 64, 64, generator=g) * 2
    lab = (torch.rand(B, 64, 64, generator=g) < 0.13).long()
    lab[0, :3] = -1
    sf = [torch.randn(B, *s, generator=g) for s in feat_shapes]
    tf = [torch.randn(B, *s, generator=g) for s in feat_shapes]
    return zs, zt, lab, sf, tf


@pytest.mark.parametrize("K,weights", [(2, [0.4, 3.5]), (3, [0.39, 2.61, 33.09]), (2, None)])
def test_kd_loss_fp32_vs_oracle(ops, K, weights):
    zs, zt, lab, sf, tf = _kd_inputs(3, K, seed=K)
    w = None if weights is None else torch.tensor(weights)
    zs_r = zs.clone().requires_grad_(True)
    sf_r = [s.clone().requires_grad_(True) for s in sf]
    ref = kd_oracle.kd_loss(zs_r, zt, lab, w, sf_r, tf, T=4.0, alpha=0.5, beta=1.0)
    ref["loss"].backward()
    terms, dz, dfe = ops.kd_loss_fwd_bwd(zs.cuda(), zt.cuda(), lab.cuda(), None if w is None else w.cuda(),
                                         [s.cuda() for s in sf], [t.cuda() for t in tf], T=4.0, alpha=0.5, beta=1.0)
    t = terms.cpu()
    for i, k in enumerate(("loss", "ce", "kl", "mse")):
        assert t[i].item() == pytest.approx(ref[k].item(), rel=1e-5), k
    assert rel_err(dz.cpu(), zs_r.grad) < 1e-5
    for d, s in zip(dfe, sf_r):
        assert rel_err(d.cpu(), s.grad) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_kd_loss_with_hoisted_label_count_is_bit_identical(ops, dtype):
    """kdf_kd_label_count (taken early, on another stream) + kdf_kd_loss_fwd_bwd_counted == kdf_kd_loss_fwd_bwd, bit for
    bit: terms, logit gradients, tap gradients."""
    zs, zt, lab, sf, tf = _kd_inputs(3, 2, seed=11)
    w = torch.tensor([0.4, 3.5]).cuda()
    args = (zs.cuda().to(dtype), zt.cuda().to(dtype), lab.cuda(), w, [s.cuda().to(dtype) for s in sf], [t.cuda().to(dtype) for t in tf])
    ref = ops.kd_loss_fwd_bwd(*args)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        ws = ops.kd_label_count(lab.cuda(), 2)
    torch.cuda.current_stream().wait_stream(side)
    got = ops.kd_loss_fwd_bwd(*args, counted_ws=ws)
    assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
    for a, b in zip(got[2], ref[2]):
        assert torch.equal(a, b)
    assert got[0][7].item() == (lab != -1).sum().item()


def test_kd_loss_ce_only_matches_reference_criterion(ops):
    """teacher=None, alpha=beta=0 -> exactly nn.CrossEntropyLoss(ignore_index=-1, weight) (trainer.py:55)."""
    zs, _, lab, _, _ = _kd_inputs(2, 2, seed=7)
    w = torch.tensor([0.4, 3.5])
    zs_r = zs.clone().requires_grad_(True)
    ref = torch.nn.CrossEntropyLoss(ignore_index=-1, weight=w)(zs_r, lab)
    ref.backward()
    terms, dz, _ = ops.kd_loss_fwd_bwd(zs.cuda(), None, lab.cuda(), w.cuda(), [], [], alpha=0.0, beta=0.0)
    assert terms[0].item() == pytest.approx(ref.item(), rel=1e-5) and terms[1].item() == pytest.approx(ref.item(), rel=1e-5)
    assert rel_err(dz.cpu(), zs_r.grad) < 1e-5
    assert terms[7].item() == (lab != -1).sum().item()


def test_kd_loss_autograd_function_and_nhwc_taps(ops):
    """KDLossFn through autograd, with channels-last / NHWC-strided taps like the model produces."""
    zs, zt, lab, sf, tf = _kd_inputs(2, 2, seed=3)
    w = torch.tensor([0.4, 3.5])
    zs_r = zs.clone().requires_grad_(True)
    sf_r = [s.clone().requires_grad_(True) for s in sf]
    ref = kd_oracle.kd_loss(zs_r, zt, lab, w, sf_r, tf)
    (ref["loss"] * 0.5).backward()
    zc = zs.cuda().requires_grad_(True)
    s_c = [sf[0].cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True),     # NHWC view
           sf[1].cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)]
    t_c = [tf[0].cuda(), tf[1].cuda()]
    loss, terms = ops.KDLossFn.apply(zc, zt.cuda(), lab.cuda(), w.cuda(), 4.0, 0.5, 1.0, -1, *s_c, *t_c)
    (loss * 0.5).backward()
    assert loss.item() == pytest.approx(ref["loss"].item(), rel=1e-5)
    assert rel_err(zc.grad.cpu(), zs_r.grad) < 1e-5
    for s, r in zip(s_c, sf_r):
        assert rel_err(s.grad.cpu(), r.grad) < 1e-5


def test_kd_loss_bf16_tolerance(ops):
    """bf16 logits/taps, fp32 accumulation.  Stated tolerance: 1e-3 relative on the
    scalar terms against the oracle evaluated on the same bf16-rounded inputs, 2e-2 on gradients."""
    zs, zt, lab, sf, tf = _kd_inputs(4, 2, seed=5)
    w = torch.tensor([0.4, 3.5])
    bf = torch.bfloat16
    zs_b, zt_b = zs.to(bf), zt.to(bf)
    sf_b, tf_b = [s.to(bf) for s in sf], [t.to(bf) for t in tf]
    zs_r = zs_b.float().requires_grad_(True)
    sf_r = [s.float().requires_grad_(True) for s in sf_b]
    ref = kd_oracle.kd_loss(zs_r, zt_b.float(), lab, w, sf_r, [t.float() for t in tf_b])
    ref["loss"].backward()
    terms, dz, dfe = ops.kd_loss_fwd_bwd(zs_b.cuda(), zt_b.cuda(), lab.cuda(), w.cuda(),
                                         [s.cuda() for s in sf_b], [t.cuda() for t in tf_b])
    for i, k in enumerate(("loss", "ce", "kl", "mse")):
        assert terms[i].item() == pytest.approx(ref[k].item(), rel=1e-3), k
    assert dz.dtype == bf and rel_err(dz.float().cpu(), zs_r.grad) < 2e-2
    for d, s in zip(dfe, sf_r):
        assert d.dtype == bf and rel_err(d.float().cpu(), s.grad) < 2e-2


def test_kd_loss_all_ignored_is_nan_like_torch(ops):
    zs, zt, lab, _, _ = _kd_inputs(1, 2, seed=1)
    lab[:] = -1
    terms, _, _ = ops.kd_loss_fwd_bwd(zs.cuda(), None, lab.cuda(), None, [], [], alpha=0.0, beta=0.0)
    assert torch.isnan(terms[1]).item() and torch.isnan(kd_oracle.ce_loss(zs, lab)).item()


# ----------------------------------------------------------------------------- fusion
def _fusion_oracle(cam_feat, lid_feat, sd, fusion_type, train):
    pre, fused, extras = model_oracle.fusion(cam_feat, lid_feat, sd, fusion_type, "fusion", train)
    return pre, fused, extras


@pytest.mark.parametrize("fusion_type", ["weighted", "minimal", "concat"])
@pytest.mark.parametrize("train", [True, False])
def test_fusion_fp32_forward_backward_vs_oracle(ops, fusion_type, train):
    """Fused fusion block (incl. BatchNorm batch statistics in training) against the oracle's
    restatement of fusion_module.py:242-256: outputs and every gradient within 1e-5 relative."""
    from src.models import fusion_module as fm
    B, C = 2, 128
    sd_full = make_state_dict(11, fusion_type=fusion_type, random_running_stats=True)
    sd = {k: v for k, v in sd_full.items() if k.startswith("fusion.")}
    g = torch.Generator().manual_seed(2)
    cam = torch.randn(B, C, 64, 64, generator=g).relu()
    lid = torch.randn(B, C, 64, 64, generator=g).relu() * (torch.rand(B, 1, 64, 64, generator=g) > 0.3)
    # oracle (CPU autograd)
    sdo = model_oracle.clone_state(sd, requires_grad=True)
    cam_o, lid_o = cam.clone().requires_grad_(True), lid.clone().requires_grad_(True)
    pre_o, fused_o, _ = _fusion_oracle(cam_o, lid_o, sdo, fusion_type, train)
    gout = torch.randn(fused_o.shape, generator=g)
    fused_o.backward(gout)
    # product (CUDA)
    cls = {"weighted": fm.WeightedFusion, "minimal": fm.MinimalFusion, "concat": fm.ConcatenationFusion}[fusion_type]
    mod = cls(128, 128, 256) if fusion_type == "concat" else cls(128, 128, 128)
    mod.load_state_dict({k[len("fusion."):]: v for k, v in sd.items()})
    mod.cuda().train(train)
    cam_c = cam.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    lid_c = lid.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    pre_c, fused_c = mod.forward_with_pre(cam_c, lid_c)
    fused_c.backward(gout.cuda())
    assert rel_err(pre_c.detach().cpu(), pre_o.detach()) < 1e-5
    assert rel_err(fused_c.detach().cpu(), fused_o.detach()) < 1e-5
    assert rel_err(cam_c.grad.cpu(), cam_o.grad) < 2e-5
    assert rel_err(lid_c.grad.cpu(), lid_o.grad) < 2e-5
    for name, p in mod.named_parameters():
        assert rel_err(p.grad.cpu(), sdo["fusion." + name].grad) < 2e-5, name
    if train:   # running statistics advanced like nn.BatchNorm2d
        for k, v in mod.state_dict().items():
            if "running_" in k or "num_batches" in k:
                assert rel_err(v.cpu().float(), sdo["fusion." + k].float()) < 1e-5, k


def test_fusion_weighted_bf16_tolerance(ops):
    """bf16 rows through the same kernel: stated tolerance 2e-2 relative (bf16 has 8 mantissa bits)
    against the fp32 oracle on the bf16-rounded inputs."""
    from src.models import fusion_module as fm
    sd = {k: v for k, v in make_state_dict(11, fusion_type="weighted").items() if k.startswith("fusion.")}
    g = torch.Generator().manual_seed(4)
    cam = torch.randn(2, 128, 64, 64, generator=g).relu().to(torch.bfloat16)
    lid = torch.randn(2, 128, 64, 64, generator=g).relu().to(torch.bfloat16)
    with torch.no_grad():
        _, fused_o, _ = _fusion_oracle(cam.float(), lid.float(), model_oracle.clone_state(sd), "weighted", True)
    mod = fm.WeightedFusion(128, 128, 128)
    mod.load_state_dict({k[len("fusion."):]: v for k, v in sd.items()})
    mod.cuda().train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = mod(cam.cuda().contiguous(memory_format=torch.channels_last),
                  lid.cuda().contiguous(memory_format=torch.channels_last))
    assert out.dtype == torch.bfloat16 and rel_err(out.float().cpu(), fused_o) < 2e-2


# ----------------------------------------------------------------------------- step helpers
def test_confusion_matrix_vs_oracle(ops):
    g = torch.Generator().manual_seed(0)
    for K in (2, 3):
        logits = torch.randn(3, K, 64, 64, generator=g)
        labels = torch.randint(0, K, (3, 64, 64), generator=g)
        labels[0, :2] = -1
        conf = torch.zeros(K, K, dtype=torch.int64, device="cuda")
        ops.confusion_matrix_(conf, logits.cuda(), labels.cuda())
        ops.confusion_matrix_(conf, logits.cuda(), labels.cuda())             # accumulates
        np.testing.assert_array_equal(conf.cpu().numpy(), 2 * kd_oracle.confusion_matrix(logits, labels, K).numpy())


def test_adamw_flat_matches_torch_adamw(ops):
    g = torch.Generator().manual_seed(0)
    n = 10_007
    p0 = torch.randn(n, generator=g)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, weight_decay=1e-3, foreach=False)
    p = p0.clone().cuda()
    pad = (n + 3) // 4 * 4
    buf = [torch.zeros(pad, device="cuda") for _ in range(4)]
    buf[0][:n] = p
    hyper = torch.zeros(2, device="cuda")
    for step in range(1, 6):
        grad = torch.randn(n, generator=g)
        ref_p.grad = grad.clone()
        opt.step()
        buf[1][:n] = grad.cuda() * 4.0
        hyper.copy_(torch.tensor([1e-3, float(step)]))
        ops.adamw_flat_(buf[0], buf[1], buf[2], buf[3], hyper, 0.9, 0.999, 1e-8, 1e-3, grad_scale=0.25)
        assert rel_err(buf[0][:n].cpu(), ref_p.detach()) < 1e-6


# ----------------------------------------------------------------------------- BatchNorm (+act) over rows
@pytest.mark.parametrize("C,act,train,residual", [(64, "relu", True, False), (192, "relu6", True, False),
                                                   (128, None, True, True), (32, "relu6", False, False),
                                                   (768, "relu6", True, False), (16, "relu", True, False)])
def test_rowbn_fp32_vs_torch_batchnorm(ops, C, act, train, residual):
    """bn_act against nn.BatchNorm2d + activation (+ shortcut add) as the reference composes them
    (camera_encoder.py:19-51, fusion_module.py:11-32): output, d input, d gamma, d beta and the running
    statistics within 1e-5 relative (fp32)."""
    import torch.nn as nn
    g = torch.Generator().manual_seed(C)
    x = torch.randn(3, C, 20, 28, generator=g) * 2 + 0.5
    res = torch.randn(3, C, 20, 28, generator=g) if residual else None
    gout = torch.randn(3, C, 20, 28, generator=g)
    bn_ref = nn.BatchNorm2d(C)
    with torch.no_grad():
        bn_ref.weight.copy_(torch.rand(C, generator=g) + 0.5); bn_ref.bias.copy_(torch.randn(C, generator=g) * 0.2)
        bn_ref.running_mean.copy_(torch.randn(C, generator=g) * 0.1); bn_ref.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    import copy
    bn_cuda = copy.deepcopy(bn_ref).cuda()
    bn_ref.train(train); bn_cuda.train(train)
    xr = x.clone().requires_grad_(True)
    y = bn_ref(xr)
    y = torch.relu(y) if act == "relu" else (torch.nn.functional.relu6(y) if act == "relu6" else y)
    if residual:
        y = y + res
    y.backward(gout)
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    rc = None if res is None else res.cuda().contiguous(memory_format=torch.channels_last)
    yc = ops.bn_act(xc, bn_cuda, act, rc)
    yc.backward(gout.cuda())
    assert rel_err(yc.detach().cpu(), y.detach()) < 1e-5
    assert rel_err(xc.grad.cpu(), xr.grad) < 2e-5
    assert rel_err(bn_cuda.weight.grad.cpu(), bn_ref.weight.grad) < 2e-5
    assert rel_err(bn_cuda.bias.grad.cpu(), bn_ref.bias.grad) < 2e-5
    assert rel_err(bn_cuda.running_mean.cpu(), bn_ref.running_mean) < 1e-5
    assert rel_err(bn_cuda.running_var.cpu(), bn_ref.running_var) < 1e-5
    assert bn_cuda.num_batches_tracked.item() == bn_ref.num_batches_tracked.item()


def test_rowbn_rows_bf16_and_large_m(ops):
    """2-D rows (the point MLP case), bf16 storage with fp32 statistics: stated tolerance 1e-2 relative
    against fp32 BatchNorm1d+ReLU on the bf16-rounded input; M large enough for several CTAs per column."""
    import torch.nn as nn
    M, C = 300_000, 128
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(M, C, generator=g) * 3 + 1).to(torch.bfloat16)
    bn_ref = nn.BatchNorm1d(C)
    import copy
    bn_cuda = copy.deepcopy(bn_ref).cuda()
    y = torch.relu(bn_ref(x.float()))
    yc = ops.bn_act(x.cuda(), bn_cuda, "relu")
    assert yc.dtype == torch.bfloat16 and rel_err(yc.float().cpu(), y) < 1e-2
    assert rel_err(bn_cuda.running_var.cpu(), bn_ref.running_var) < 1e-4


@pytest.mark.parametrize("M", [128 * 3, 128 * 148 + 57, 8192])
def test_fusion_weighted_tensor_core_kernels_vs_fp32_autograd(ops, M):
    """The tcgen05 weighted-fusion kernels (bf16 rows, C=128) through the C ABI against fp32 torch autograd of
    fusion_module.py:126-136 on the same bf16 inputs.  bf16 operands / fp32 accumulation: stated tolerance
    1e-2 relative L2 on outputs and gradients (2e-2 max-norm on the output)."""
    from src import native
    C = 128
    g = torch.Generator().manual_seed(M)
    dev = "cuda"
    cam = (torch.randn(M, C, generator=g)).to(torch.bfloat16).to(dev)
    lid = (torch.randn(M, C, generator=g) * (torch.rand(M, 1, generator=g) > 0.3)).to(torch.bfloat16).to(dev)
    f = lambda *s, k=1.0: (torch.randn(*s, generator=g) * k).to(dev)
    csc, csh, lsc, lsh = f(C).abs() + 0.5, f(C, k=0.3), -(f(C).abs() + 0.5), f(C, k=0.3)
    w1, b1, w2, b2 = f(C, 2 * C, k=0.08), f(C, k=0.2), f(2, C, k=0.3), f(2, k=0.1)
    gout = f(M, C).to(torch.bfloat16)
    p, st = native.ptr, native.stream_ptr(torch.device(dev))
    out = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
    attn = torch.empty(M, 2, device=dev)
    native.call("kdf_fusion_weighted_fwd", p(cam), p(lid), native.KDF_BF16, M, C, p(csc), p(csh), p(lsc), p(lsh),
                p(w1), p(b1), p(w2), p(b2), p(out), p(attn), st)
    gcam, glid = torch.empty_like(cam), torch.empty_like(lid)
    gaff, gw1, gb1 = torch.empty(4, C, device=dev), torch.empty(C, 2 * C, device=dev), torch.empty(C, device=dev)
    gw2, gb2 = torch.empty(2, C, device=dev), torch.empty(2, device=dev)
    native.call("kdf_fusion_weighted_bwd", p(gout), p(cam), p(lid), native.KDF_BF16, M, C, p(csc), p(csh), p(lsc), p(lsh),
                p(w1), p(b1), p(w2), p(b2), p(attn), p(gcam), p(glid), p(gaff), p(gw1), p(gb1), p(gw2), p(gb2), st)
    # autograd reference in fp64 on the operands the tensor cores see: activations and W1 rounded to bf16
    # (straight-through), so that the hidden layer's ReLU masks are the same ones (a mask that flips because
    # of operand rounding moves the gradient by a whole unit's worth, which is not what this test measures)
    leaves = [t.clone().double().requires_grad_(True) for t in (cam, lid, csc, csh, lsc, lsh, w1, b1, w2, b2)]
    xc, xl, rcsc, rcsh, rlsc, rlsh, rw1, rb1, rw2, rb2 = leaves
    ste = lambda t: t + (t.detach().float().to(torch.bfloat16).double() - t.detach())
    yc = ste(torch.relu((xc.float() * rcsc.float() + rcsh.float()).double()))
    yl = ste(torch.relu((xl.float() * rlsc.float() + rlsh.float()).double()))
    hid = torch.relu(torch.cat([yc, yl], 1) @ ste(rw1).t() + rb1)
    w = torch.softmax(hid @ rw2.t() + rb2, dim=1)
    ref = yc * w[:, :1] + yl * w[:, 1:]
    ref.backward(gout.double())

    def l2(a, b):
        a, b = a.double(), b.double()
        return ((a - b).norm() / (b.norm() + 1e-30)).item()
    assert rel_err(out.float(), ref.detach()) < 2e-2 and l2(out.float(), ref.detach()) < 5e-3
    assert l2(attn, w.detach()) < 5e-3
    assert l2(gcam.float(), xc.grad) < 1e-2, l2(gcam.float(), xc.grad)
    assert l2(glid.float(), xl.grad) < 1e-2, l2(glid.float(), xl.grad)
    ref_aff = torch.stack([rcsc.grad, rcsh.grad, rlsc.grad, rlsh.grad])
    assert l2(gaff, ref_aff) < 1e-2, l2(gaff, ref_aff)
    for name, got, want in (("w1", gw1, rw1.grad), ("b1", gb1, rb1.grad), ("w2", gw2, rw2.grad), ("b2", gb2, rb2.grad)):
        assert l2(got, want) < 1e-2, (name, l2(got, want))


@pytest.mark.parametrize("dtype,two", [(torch.float32, True), (torch.float32, False), (torch.bfloat16, True)])
def test_fpn_merge_vs_interpolate(ops, dtype, two):
    """out = base + bilinear(lo_a) (+ bilinear(lo_b)) against F.interpolate(align_corners=False) + adds
    (fusion_module.py:61-63), forward and the three input gradients: 1e-5 relative in fp32, 1e-2 in bf16."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    B, C, H, W = 2, 128, 64, 48
    mk = lambda *s: torch.randn(*s, generator=g).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    base, lo_a, lo_b = mk(B, C, H, W), mk(B, C, H // 2, W // 2), mk(B, C, H // 2, W // 2)
    lows = [lo_a, lo_b] if two else [lo_a]
    leaves = [t.clone().requires_grad_(True) for t in [base] + lows]
    out = ops.fpn_merge(leaves[0], leaves[1:])
    assert out is not None and out.shape == (B, C, H, W) and out.dtype == dtype
    gout = torch.randn(B, C, H, W, generator=g).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    out.backward(gout)
    ref_leaves = [t.clone().float().requires_grad_(True) for t in [base] + lows]
    ref = ref_leaves[0]
    for t in ref_leaves[1:]:
        ref = ref + F.interpolate(t, size=(H, W), mode="bilinear", align_corners=False)
    ref.backward(gout.float())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(out.float().cpu(), ref.detach().cpu()) < tol
    for a, b in zip(leaves, ref_leaves):
        assert rel_err(a.grad.float().cpu(), b.grad.cpu()) < tol
    # layouts the fused path does not serve are declined (the module then composes torch ops)
    assert ops.fpn_merge(base, [mk(B, C, H // 4, W // 4)]) is None


@pytest.mark.parametrize("C,H,W,stride,dtype", [(32, 20, 28, 1, torch.float32), (192, 17, 23, 2, torch.float32),
                                                (384, 16, 16, 1, torch.float32), (64, 9, 31, 2, torch.float32),
                                                (768, 8, 8, 1, torch.bfloat16), (192, 32, 32, 2, torch.bfloat16),
                                                (128, 64, 64, 1, torch.bfloat16)])
def test_dwconv3x3_vs_torch_conv2d(ops, C, H, W, stride, dtype):
    """Depthwise 3x3 (padding 1, stride 1|2, no bias) against nn.Conv2d(groups=C) as the reference builds it
    (camera_encoder.py:27-33, fusion_module.py:24-27): output, d input, d weight.  fp32: 1e-5 relative;
    bf16 storage with fp32 accumulation: 1e-2 against the fp32 convolution of the same bf16 values."""
    import torch.nn as nn
    g = torch.Generator().manual_seed(C + H)
    conv = nn.Conv2d(C, C, 3, stride=stride, padding=1, groups=C, bias=False)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(C, 1, 3, 3, generator=g) * 0.3)
    x = torch.randn(3, C, H, W, generator=g).to(dtype)
    ref_x = x.float().clone().requires_grad_(True)
    ref = conv(ref_x)
    gout = torch.randn(ref.shape, generator=g).to(dtype)
    ref.backward(gout.float())
    ref_gw = conv.weight.grad.clone()
    conv.weight.grad = None
    conv.cuda()
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = ops.dwconv3x3(conv, xc)
    assert out is not None and out.dtype == dtype and tuple(out.shape) == tuple(ref.shape)
    out.backward(gout.cuda())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(out.float().cpu(), ref.detach()) < tol
    assert rel_err(xc.grad.float().cpu(), ref_x.grad) < tol
    assert rel_err(conv.weight.grad.cpu(), ref_gw) < (2e-5 if dtype == torch.float32 else 1e-2)
    # not a depthwise 3x3: declined
    assert ops.dwconv3x3(nn.Conv2d(C, C, 1).cuda(), xc) is None


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("stride,act", [(1, "relu6"), (2, "relu6"), (1, "relu"), (2, None)])
def test_dwconv3x3_eval_batchnorm_epilogue_bit_identical(ops, dtype, stride, act):
    """Inference: depthwise conv + running-statistics BatchNorm + activation in one kernel (run_fused under no_grad)
    against the two-kernel sequence (stencil kernel, then the row BatchNorm kernel): bit-identical."""
    import torch.nn as nn
    C, H, W = 96, 20, 24
    g = torch.Generator().manual_seed(7 + stride)
    layers = [nn.Conv2d(C, C, 3, stride=stride, padding=1, groups=C, bias=False), nn.BatchNorm2d(C)]
    if act:
        layers.append(nn.ReLU6() if act == "relu6" else nn.ReLU())
    seq = nn.Sequential(*layers)
    with torch.no_grad():
        seq[0].weight.copy_(torch.randn(C, 1, 3, 3, generator=g) * 0.5)
        seq[1].weight.copy_(torch.rand(C, generator=g) + 0.5)
        seq[1].bias.copy_(torch.randn(C, generator=g))
        seq[1].running_mean.copy_(torch.randn(C, generator=g) * 0.2)
        seq[1].running_var.copy_(torch.rand(C, generator=g) + 0.5)
    seq.cuda().eval()
    x = (torch.randn(2, C, H, W, generator=g) * 3).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        fused = ops.run_fused(seq, x)
    with torch.enable_grad():                                   # the epilogue is an inference-only path
        two = ops.run_fused(seq, x).detach()
    assert fused.dtype == dtype and fused.shape == two.shape
    assert torch.equal(fused, two)
    ref = seq(x.float())
    assert rel_err(fused.float().cpu(), ref.cpu()) < (1e-5 if dtype == torch.float32 else 1e-2)
    # a shortcut added after this BatchNorm keeps the separate kernel (and still agrees)
    with torch.no_grad():
        res = ops.run_fused(nn.Sequential(*list(seq)[:2]), x, residual=x) if stride == 1 else None
    if res is not None:
        assert rel_err(res.float().cpu(), (nn.Sequential(*list(seq)[:2])(x.float()) + x.float()).cpu()) < (1e-5 if dtype == torch.float32 else 1e-2)


def test_frozen_conv_bf16_weight_cache(ops):
    """Inference under bf16 autocast: run_fused keeps the bf16 copy of a frozen convolution's weight (the teacher) instead
    of letting autocast re-cast it every step; same output as the module, and the copy follows weight updates."""
    import torch.nn as nn
    conv = nn.Conv2d(32, 48, 1, bias=True).cuda()
    seq = nn.Sequential(conv)
    x = torch.randn(2, 32, 16, 16, device="cuda").contiguous(memory_format=torch.channels_last)
    with torch.autocast("cuda", dtype=torch.bfloat16), torch.no_grad():
        a = ops.run_fused(seq, x)
        b = conv(x)
        assert a.dtype == torch.bfloat16 and torch.equal(a, b)
        w16 = conv._kdf_w16[1]
        assert ops.run_fused(seq, x).data_ptr() != a.data_ptr() and conv._kdf_w16[1] is w16      # cached
        conv.weight.mul_(2.0)                                                                   # in-place update: new version
        c = ops.run_fused(seq, x)
        assert conv._kdf_w16[1] is not w16
    with torch.autocast("cuda", dtype=torch.bfloat16), torch.no_grad():      # a fresh context: autocast's own cast cache is per context
        assert torch.equal(c, conv(x)) and not torch.equal(c, a)
    with torch.autocast("cuda", dtype=torch.bfloat16):                                          # training: the module itself
        y = ops.run_fused(seq, x)
        assert y.requires_grad


def test_camera_side_kernels_accept_empty_batches(ops):
    import torch.nn as nn
    conv = nn.Conv2d(64, 64, 3, padding=1, groups=64, bias=False).cuda()
    x = torch.empty(0, 64, 8, 8, device="cuda").contiguous(memory_format=torch.channels_last)
    y = ops.dwconv3x3(conv, x)
    assert y is not None and tuple(y.shape) == (0, 64, 8, 8)
    base = torch.empty(0, 64, 8, 8, device="cuda").contiguous(memory_format=torch.channels_last)
    lo = torch.empty(0, 64, 4, 4, device="cuda").contiguous(memory_format=torch.channels_last)
    out = ops.fpn_merge(base, [lo])
    assert out is not None and tuple(out.shape) == (0, 64, 8, 8)


# ----------------------------------------------------------------------------- range-view projection (parity unpinned: not in the reference)
def _sweep(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    pts = torch.randn(B, N, 4, generator=g)
    pts[..., :2] *= 30.0
    pts[..., 2] = pts[..., 2] * 1.5 - 1.0
    pts[0, :6] = torch.tensor([[0.0, 0.0, 0.0, 1.0], [float("nan"), 1.0, 0.0, 0.0], [float("inf"), 0.0, 0.0, 0.0],
                               [10.0, 0.0, 30.0, 0.0], [10.0, 0.0, -30.0, 0.0], [-5.0, 0.0, 0.0, 0.0]])   # origin, NaN, inf, above, below, behind
    return pts


@pytest.mark.parametrize("grid,fov", [((64, 512), (3.0, -25.0)), ((32, 1024), (15.0, -15.0))])
def test_range_index_vs_float64_oracle(ops, grid, fov):
    """kdf_range_index (fp32 atan2f / asinf) against the float64 evaluation of the written convention: identical cells
    except for points whose image coordinates are within 1e-3 of a cell boundary (or of the field-of-view limit); the
    occupancy is the histogram of the kernel's own cells; invalid points (origin, NaN / inf, outside the vertical field
    of view) are -1."""
    from oracle import bev_oracle
    pts = _sweep(3, 40000, 5)
    cell, count = ops.range_index(pts.cuda(), grid, fov)
    ref, margin = bev_oracle.range_cells(pts.numpy(), grid, fov)
    got = cell.cpu().numpy()
    differ = got != ref
    assert differ.mean() < 2e-3 and (margin[differ] < 1e-3).all(), (differ.sum(), margin[differ].max() if differ.any() else 0)
    assert (got[0, :5] == -1).all() and got[0, 5] >= 0
    assert ((got >= 0).mean() > 0.3) and got.max() < grid[0] * grid[1]
    H, W = grid
    occ = np.stack([np.bincount(c[c >= 0], minlength=H * W) for c in got]).astype(np.int32)
    np.testing.assert_array_equal(count.cpu().numpy(), occ)


@pytest.mark.parametrize("dtype,reduce", [(torch.float32, "max"), (torch.bfloat16, "max"), (torch.float32, "mean")])
def test_range_project_matches_torch_scatter_on_its_cells(ops, dtype, reduce):
    """The range image is the per-cell max / mean over the kernel's own cell ids (exact for max), and the gradient is
    ATen's for the same scatter (even tie split)."""
    B, N, C, H, W = 2, 20000, 128, 32, 256
    pts = _sweep(B, N, 9).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    feats = torch.rand(B, N, C, generator=g, device="cuda").to(dtype).requires_grad_(True)
    img, count, cell = ops.range_project(pts, feats, (H, W), (3.0, -25.0), reduce)
    assert img.shape == (B, C, H, W) and img.dtype == dtype
    flat = (cell.long() + torch.arange(B, device="cuda").view(B, 1) * (H * W)).reshape(-1)
    ok = (cell >= 0).reshape(-1)
    fr = feats.detach().float().reshape(-1, C).requires_grad_(True)
    ref = torch.zeros(B * H * W, C, device="cuda").scatter_reduce(0, flat[ok][:, None].expand(-1, C), fr[ok],
                                                                     "amax" if reduce == "max" else "mean", include_self=False)
    got = img.permute(0, 2, 3, 1).reshape(B * H * W, C).float()
    if reduce == "max":
        assert torch.equal(got, ref.detach())
    else:
        assert rel_err(got.cpu(), ref.detach().cpu()) < 1e-5
    gout = torch.randn(B, C, H, W, device="cuda").to(dtype)
    img.backward(gout)
    ref.backward(gout.permute(0, 2, 3, 1).reshape(B * H * W, C).float())
    assert rel_err(feats.grad.float().cpu(), fr.grad.view(B, N, C).cpu()) < (1e-5 if dtype == torch.float32 else 1e-2)
    assert (feats.grad[~(cell >= 0)] == 0).all()


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (3, 37, 51), (1, 256, 256), (2, 9, 130)])
def test_stem_conv_vs_torch_conv2d(ops, B, H, W):
    """The camera stem (camera_encoder.py:63-67) from the fp32 NCHW image: rows + BatchNorm column sums in training,
    folded BatchNorm + ReLU6 in inference, weight gradient -- against nn.Conv2d on the bf16-rounded image and taps
    (what the autocast convolution sees), fp32 accumulation."""
    import torch.nn as nn
    from src.native import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(H * 7 + W)
    conv = nn.Conv2d(3, 32, 3, stride=2, padding=1, bias=False)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(32, 3, 3, 3, generator=g) * 0.3)
    x = torch.rand(B, 3, H, W, generator=g)
    xr, wr = x.bfloat16().float(), conv.weight.detach().bfloat16().float()
    ref_w = wr.clone().requires_grad_(True)
    ref = torch.nn.functional.conv2d(xr, ref_w, None, 2, 1)
    OH, OW = ref.shape[2:]
    gout = torch.randn(ref.shape, generator=g).bfloat16()
    ref.backward(gout.float())
    dev = "cuda"
    xc, wc = x.to(dev), conv.weight.detach().to(dev)
    out = torch.empty(B, OH, OW, 32, dtype=torch.bfloat16, device=dev)
    stats = torch.empty(2, 32, dtype=torch.float64, device=dev)
    call("kdf_stem_conv_fwd", ptr(xc), ptr(wc), B, H, W, None, None, 0, ptr(out), ptr(stats), stream_ptr(xc.device))
    got = out.permute(0, 3, 1, 2).float().cpu()
    assert rel_err(got, ref.detach()) < 4e-3                       # bf16 storage of an fp32 accumulation
    assert (got - ref.detach()).abs().max() <= 2.0 ** -7 * ref.detach().abs().max()
    st = torch.stack([got.double().sum((0, 2, 3)), (got.double() ** 2).sum((0, 2, 3))])
    np.testing.assert_allclose(stats.cpu().numpy(), st.numpy(), rtol=2e-6, atol=1e-6)
    gw = torch.empty(32, 27, dtype=torch.float32, device=dev)
    gr = gout.to(dev).permute(0, 2, 3, 1).contiguous()
    call("kdf_stem_conv_bwd_weight", ptr(xc), ptr(gr), B, H, W, ptr(gw), stream_ptr(xc.device))
    assert rel_err(gw.view(32, 3, 3, 3).cpu(), ref_w.grad) < 2e-5
    # inference: folded BatchNorm + ReLU6 in the same kernel == the two-step sequence on the stored rows
    scale, shift = (torch.rand(32, generator=g) + 0.5).to(dev), torch.randn(32, generator=g).to(dev)
    fused = torch.empty_like(out)
    call("kdf_stem_conv_fwd", ptr(xc), ptr(wc), B, H, W, ptr(scale), ptr(shift), 2, ptr(fused), None, stream_ptr(xc.device))
    two = torch.clamp(torch.addcmul(shift, out.float(), scale), 0, 6).to(torch.bfloat16)
    assert torch.equal(fused, two)


def test_stem_module_path_matches_library_path(ops):
    """TwinLiteEncoder.stem through ops.stem_conv (training and inference) against the same Sequential on the library
    convolution (run_fused over the channels-last image): outputs and parameter gradients within bf16 noise."""
    import copy
    import torch.nn as nn
    g = torch.Generator().manual_seed(3)
    stem = nn.Sequential(nn.Conv2d(3, 32, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(32), nn.ReLU6()).cuda()
    with torch.no_grad():
        stem[1].weight.copy_(torch.rand(32, generator=g) + 0.5)
        stem[1].bias.copy_(torch.randn(32, generator=g) * 0.2)
    lib = copy.deepcopy(stem)
    x = torch.rand(4, 3, 96, 128, generator=g).cuda()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ops.stem_conv(stem, x)
        y_lib = ops.run_fused(lib, x.contiguous(memory_format=torch.channels_last))
    assert y is not None and y.dtype == torch.bfloat16 and y.shape == y_lib.shape
    assert rel_err(y.float().cpu(), y_lib.float().cpu()) < 1e-2
    gy = torch.randn(y.shape, generator=g).cuda().to(torch.bfloat16)
    y.backward(gy)
    y_lib.backward(gy)
    for a, b in zip(stem.parameters(), lib.parameters()):
        assert rel_err(a.grad.cpu(), b.grad.cpu()) < 2e-2
    for a, b in zip(stem.buffers(), lib.buffers()):
        assert torch.allclose(a.float().cpu(), b.float().cpu(), rtol=1e-3, atol=1e-4)
    stem.eval(); lib.eval()
    with torch.autocast("cuda", dtype=torch.bfloat16), torch.no_grad():
        e = ops.stem_conv(stem, x)
        e_lib = ops.run_fused(lib, x.contiguous(memory_format=torch.channels_last))
    assert e is not None and rel_err(e.float().cpu(), e_lib.float().cpu()) < 1e-2
    # not the stem's shape / dtype: declined
    assert ops.stem_conv(stem, x.double()) is None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert ops.stem_conv(nn.Sequential(nn.Conv2d(3, 16, 3, stride=2, padding=1, bias=False), nn.BatchNorm2d(16)).cuda(), x) is None


@pytest.mark.parametrize("K,B,H,W", [(2, 3, 64, 64), (1, 2, 5, 7), (4, 2, 16, 24), (3, 1, 33, 9)])
def test_cls_conv_vs_torch_conv2d(ops, K, B, H, W):
    """The head's classifier nn.Conv2d(32, K, 1) with bias (fusion_module.py:162-173) on the classifier kernel: planar
    bf16 logits, data / weight / bias gradients, against fp32 torch on the bf16-rounded operands."""
    import torch.nn as nn
    g = torch.Generator().manual_seed(K * 100 + H)
    conv = nn.Conv2d(32, K, 1).cuda()
    with torch.no_grad():
        conv.weight.copy_(torch.randn(K, 32, 1, 1, generator=g) * 0.3)
        conv.bias.copy_(torch.randn(K, generator=g))
    x = torch.randn(B, 32, H, W, generator=g).to(torch.bfloat16)
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ops.cls_conv(conv, xc)
    assert y is not None and y.dtype == torch.bfloat16 and y.is_contiguous() and tuple(y.shape) == (B, K, H, W)
    ref_x = x.float().requires_grad_(True)
    ref_w = conv.weight.detach().cpu().bfloat16().float().requires_grad_(True)
    ref_b = conv.bias.detach().cpu().clone().requires_grad_(True)
    ref = torch.nn.functional.conv2d(ref_x, ref_w, ref_b)
    assert rel_err(y.float().cpu(), ref.detach()) < 4e-3
    gout = torch.randn(ref.shape, generator=g).to(torch.bfloat16)
    y.backward(gout.cuda())
    ref.backward(gout.float())
    assert rel_err(xc.grad.float().cpu(), ref_x.grad) < 4e-3
    assert rel_err(conv.weight.grad.cpu(), ref_w.grad) < 2e-5
    assert rel_err(conv.bias.grad.cpu(), ref_b.grad) < 2e-5
    # other shapes / dtypes are declined (the module then runs)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert ops.cls_conv(nn.Conv2d(64, 2, 1).cuda(), xc) is None
    assert ops.cls_conv(conv, xc.float()) is None
