"""GPU tests of the fused LiDAR branch (src/point_mlp.py): the projection kernels that apply BatchNorm+ReLU on
the fly, the point moments, and the whole branch (forward, parameter gradients, running statistics) against
(a) an fp32 torch restatement of the reference's branch (lidar_encoder.py:25-35, 57-99) and (b) the
layer-by-layer kernels of this repo under the same bf16 autocast."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _frames(seed, B, N, dev):
    from oracle.weights import synthetic_frames
    _, pts, _ = synthetic_frames(seed, B, N, edge_cases=True, nonfinite=False)
    return pts.to(dev)


def test_point_moments():
    from src import point_mlp
    pts = _frames(3, 3, 5000, "cuda")
    m = point_mlp.point_moments(pts).cpu()
    p = pts.reshape(-1, 4).double().cpu()
    ref = [p[:, k].sum() for k in range(4)] + [(p[:, i] * p[:, j]).sum() for i in range(4) for j in range(i, 4)]
    mag = [p[:, k].abs().sum() for k in range(4)] + [(p[:, i] * p[:, j]).abs().sum() for i in range(4) for j in range(i, 4)]
    # fp32 partial sums per thread / warp, fp64 across CTAs: 2e-6 of the sum of magnitudes
    assert ((m - torch.stack(ref)).abs() <= 2e-6 * torch.stack(mag)).all()


@pytest.mark.parametrize("B,N", [(2, 4000), (3, 20000)])
def test_bev_affine_reduce_and_backward(B, N):
    """Per-cell max of bf16(relu(z*scale+shift)) via the per-cell extreme of z (max for scale >= 0, min for
    scale < 0), and the gradient w.r.t. the BatchNorm output shared among the rows at the extreme."""
    from src import ops, point_mlp
    dev = "cuda"
    pts = _frames(11, B, N, dev)
    g = torch.Generator(device="cpu").manual_seed(5)
    C, H, W = 128, 64, 64
    z = (torch.randn(B * N, C, generator=g) * 1.5).to(torch.bfloat16).to(dev)
    scale = (torch.randn(C, generator=g)).to(dev)            # both signs
    shift = (torch.randn(C, generator=g) * 0.5).to(dev)
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    cell, count, order, offsets = point_mlp.bev_build_order(pts, geom, (H, W))
    grid, grid_z = point_mlp.bev_reduce_affine(z, scale, shift, order, offsets, B, N, (H, W), True)
    # reference: exact affine in fp64 -> fp32 -> relu -> bf16 for every row, then scatter amax (the reference's order)
    a3 = torch.relu((z.double() * scale.double() + shift.double()).float()).to(torch.bfloat16).float()
    flat = (cell.long() + torch.arange(B, device=dev)[:, None] * (H * W)).reshape(-1)
    valid = cell.reshape(-1) >= 0
    idx = flat[valid][:, None].expand(-1, C)
    ref = torch.zeros(B * H * W, C, device=dev)
    ref.scatter_reduce_(0, idx, a3[valid], "amax", include_self=True)
    assert torch.equal(grid.reshape(B * H * W, C).float(), ref)
    # the extreme of z: max where scale >= 0, min where scale < 0
    zs = torch.where(scale >= 0, z.float(), -z.float())
    ext = torch.full((B * H * W, C), -float("inf"), device=dev)
    ext.scatter_reduce_(0, idx, zs[valid], "amax", include_self=True)
    ext = torch.where(scale >= 0, ext, -ext)
    occupied = (count.reshape(-1) > 0)[:, None].expand(-1, C)
    assert torch.equal(grid_z.reshape(B * H * W, C).float()[occupied], ext[occupied])
    assert (grid_z.reshape(B * H * W, C)[~occupied] == 0).all() and (grid.reshape(B * H * W, C)[~occupied] == 0).all()

    gg = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).to(dev)
    dy, sums = point_mlp.bev_bwd_affine(gg, z, grid, grid_z, order, offsets, cell, B, N, (H, W))
    at_ext = (z.float() == ext[flat.clamp_min(0)]) & valid[:, None]
    k = torch.zeros(B * H * W, C, device=dev)
    k.index_add_(0, flat[valid], at_ext[valid].float())
    share = (gg.reshape(-1, C).float() / k.clamp_min(1)).to(torch.bfloat16).float()
    ref_dy = torch.where(at_ext & (ref[flat.clamp_min(0)] > 0), share[flat.clamp_min(0)], torch.zeros((), device=dev))
    assert torch.equal(dy.float(), ref_dy)
    np.testing.assert_allclose(sums[0].cpu().numpy(), ref_dy.double().sum(0).cpu().numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(sums[1].cpu().numpy(), (ref_dy.double() * z.double()).sum(0).cpu().numpy(), rtol=1e-5, atol=1e-4)
    # the rows at the extreme are arg-max rows of the activation: the gradient lands only where a3 equals the cell max
    assert ((ref_dy != 0) <= (a3 == ref[flat.clamp_min(0)])).all()


def _reference_branch(enc, pts):
    """fp32 torch restatement of the reference's branch on the GPU, using the module's own layers."""
    B, N, _ = pts.shape
    H, W = enc.grid_size
    feats = enc.point_mlp(pts.transpose(1, 2)).transpose(1, 2)           # [B,N,C] (lidar_encoder.py:66)
    cell, _ = enc.bev_cells(pts)
    flat = (cell.long() + torch.arange(B, device=pts.device)[:, None] * (H * W)).reshape(-1)
    valid = flat >= 0
    valid &= (cell.reshape(-1) >= 0)
    C = feats.shape[-1]
    out = torch.zeros(B * H * W, C, device=pts.device)
    out = out.scatter_reduce(0, flat[valid][:, None].expand(-1, C), feats.reshape(-1, C)[valid], "amax", include_self=False)
    return out.view(B, H, W, C).permute(0, 3, 1, 2)


@pytest.mark.parametrize("B,N", [(2, 6000), (4, 30011)])
def test_fused_branch_matches_reference_and_layerwise(B, N):
    from src.models.lidar_encoder import SpatialLiDAREncoder
    dev = "cuda"
    torch.manual_seed(B * 100 + 7)
    enc = SpatialLiDAREncoder(grid_size=(64, 64)).to(dev).train()
    with torch.no_grad():                                        # non-trivial BatchNorm affine parameters
        for m in enc.point_mlp:
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    ref_enc, lay_enc = copy.deepcopy(enc), copy.deepcopy(enc)
    lay_enc.fuse_point_mlp = False
    pts = _frames(B, B, N, dev)
    gout = torch.randn(B, 128, 64, 64, device=dev) * (torch.rand(B, 128, 64, 64, device=dev) < 0.5)

    ref = _reference_branch(ref_enc, pts)
    ref.backward(gout)
    outs = {}
    for name, e in (("fused", enc), ("layerwise", lay_enc)):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = e(pts)
        assert y.dtype == torch.bfloat16 and y.shape == (B, 128, 64, 64)
        y.backward(gout.to(torch.bfloat16))
        outs[name] = y
    # forward: bf16 features through three layers
    assert _l2(outs["fused"].float(), ref) < 1.5e-2, _l2(outs["fused"].float(), ref)
    assert _l2(outs["fused"].float(), outs["layerwise"].float()) < 1.5e-2
    # empty cells are exactly zero, occupancy / cell ids identical to the unfused path
    empty = (enc.last_occupancy == 0).view(B, 1, 64, 64).expand(-1, 128, -1, -1)
    assert (outs["fused"][empty] == 0).all() and (outs["layerwise"][empty] == 0).all()
    assert (ref[empty] == 0).all()
    assert torch.equal(enc.last_cells, lay_enc.last_cells) and torch.equal(enc.last_occupancy, lay_enc.last_occupancy)
    # parameter gradients and running statistics
    for (n, p), (_, pr), (_, pl) in zip(enc.named_parameters(), ref_enc.named_parameters(), lay_enc.named_parameters()):
        assert p.grad is not None, n
        if n.endswith("0.bias") or n.endswith("3.bias") or n.endswith("6.bias"):
            assert p.grad.abs().max().item() == 0.0              # conv biases cancel under batch statistics
            continue
        e_ref, e_lay = _l2(p.grad, pr.grad), _l2(pl.grad, pr.grad)
        assert e_ref < max(4e-2, 2.0 * e_lay), (n, e_ref, e_lay)
    for (n, b), (_, br) in zip(enc.named_buffers(), ref_enc.named_buffers()):
        if "running" in n:
            assert _l2(b, br) < 3e-3, (n, _l2(b, br))
        elif "num_batches" in n:
            assert torch.equal(b, br)


def test_fused_branch_eval_mode_matches_layerwise():
    """Teacher path: eval-mode BatchNorm folded into the prologues, no statistics, no gradient state."""
    from src.models.lidar_encoder import SpatialLiDAREncoder
    dev = "cuda"
    torch.manual_seed(5)
    enc = SpatialLiDAREncoder(grid_size=(64, 64)).to(dev)
    with torch.no_grad():
        for m in enc.point_mlp:
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.5)
                m.running_var.uniform_(0.5, 2.0)
                m.weight.uniform_(0.5, 1.5)
    enc.eval()
    pts = _frames(9, 2, 9000, dev)
    with torch.no_grad():
        ref = _reference_branch(enc, pts)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = enc(pts)
    assert y.dtype == torch.bfloat16
    assert _l2(y.float(), ref) < 1.5e-2, _l2(y.float(), ref)


def test_fused_kernels_accept_empty_inputs():
    """M = 0 / B = 0 / N = 0 are no-ops that still define their reduction outputs (zeros)."""
    from src import ops, point_mlp
    dev = "cuda"
    f32 = dict(device=dev, dtype=torch.float32)
    W3 = torch.zeros(128, 128, device=dev, dtype=torch.bfloat16)
    z, st = point_mlp.mlp_layer_fwd_raw(1, torch.empty(0, 128, device=dev, dtype=torch.bfloat16), torch.ones(128, **f32),
                                        torch.zeros(128, **f32), W3)
    assert z.shape == (0, 128) and (st == 0).all()
    e = torch.empty(0, 128, device=dev, dtype=torch.bfloat16)
    dyp, sums, dW = ops.mlp_layer_bwd(1, e, e, torch.ones(128, **f32), torch.zeros(128, **f32), torch.zeros(128, **f32), e,
                                      torch.ones(128, **f32), torch.zeros(128, **f32), W3)
    assert dyp.shape == (0, 128) and (sums == 0).all() and (dW == 0).all()
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    pts = torch.empty(2, 0, 4, **f32)
    cell, count, order, offsets = point_mlp.bev_build_order(pts, geom, (64, 64))
    assert (count == 0).all() and (offsets == 0).all()
    grid, grid_z = point_mlp.bev_reduce_affine(e, torch.ones(128, **f32), torch.zeros(128, **f32), order, offsets, 2, 0, (64, 64), True)
    assert (grid == 0).all() and (grid_z == 0).all()
    assert (point_mlp.point_moments(pts) == 0).all()


def test_trainer_side_stream_teacher_matches_single_stream():
    """The teacher's forward on a side stream (Trainer(overlap_teacher=True)) is a scheduling change only."""
    from oracle.weights import make_state_dict, synthetic_frames
    from src.models.camera_encoder import TwinLiteEncoder
    from src.models.fusion_module import CompleteSegmentationModel
    from src.models.lidar_encoder import LiDAREncoder
    from src.training.trainer import Trainer

    def make(ft, oc):
        return CompleteSegmentationModel(TwinLiteEncoder(return_multiscale=True), LiDAREncoder("spatial", grid_size=(64, 64)),
                                         num_classes=2, fusion_type=ft, fusion_out_channels=oc,
                                         camera_fpn_stages=["stage3", "stage4", "stage5"], camera_fpn_channels=128, output_mode="same")
    img, pts, lab = synthetic_frames(4, 2, 3000, edge_cases=True, nonfinite=False)
    outs = []
    for overlap in (False, True):
        s, t = make("weighted", 128), make("concat", 256)
        s.load_state_dict(make_state_dict(5, fusion_type="weighted"))
        t.load_state_dict(make_state_dict(6, fusion_type="concat", random_running_stats=True))
        tr = Trainer(s.cuda().train(), [], [], "cuda", class_weights=[0.4, 3.5], teacher=t.cuda().eval(), verbose=False,
                     amp_dtype=torch.bfloat16, overlap_teacher=overlap, save_dir="gpurun_out/test_ckpt")
        terms, logits = tr.training_step(img.cuda(), pts.cuda(), lab.cuda())
        outs.append((terms[:4].cpu(), logits.float().cpu()))
    # fp64 atomics in the statistics make runs differ in the last bits at most
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-4, atol=1e-5)
    assert _l2(outs[0][1], outs[1][1]) < 1e-2


@pytest.mark.parametrize("B,N", [(2, 4000), (3, 20001), (1, 9000)])
def test_bev_build_sorted_is_the_counting_sort_applied_to_the_points(B, N):
    """the points themselves in cell order.  sorted_points = points[order] with every cell's rows
    contiguous ([offsets[c], offsets[c+1])), the points outside the grid behind them, cell_sorted = the global cell id
    of every sorted row; count / cell / offsets identical to kdf_bev_build_order."""
    from src import ops, point_mlp
    dev = "cuda"
    pts = _frames(21, B, N, dev)
    H = W = 64
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    cell0, count0, _order0, offsets0 = point_mlp.bev_build_order(pts, geom, (H, W))
    cell, count, offsets, spts, cell_sorted, order = point_mlp.bev_build_sorted(pts, geom, (H, W), want_order=True)
    assert torch.equal(cell, cell0) and torch.equal(count, count0) and torch.equal(offsets, offsets0)
    for b in range(B):
        o = order[b].long()
        assert torch.equal(torch.sort(o).values, torch.arange(N, device=dev))              # a permutation
        assert torch.equal(spts[b].view(torch.int32), pts[b][o].view(torch.int32))          # bit-exact copy of the points
        c = cell[b][o]
        nv = int(offsets[b, H * W])
        assert (c[:nv] >= 0).all() and (c[nv:] < 0).all()
        assert (c[:nv][1:] >= c[:nv][:-1]).all()                                            # grouped by cell, ascending
        want = torch.where(c >= 0, c + b * H * W, torch.full_like(c, -1))
        assert torch.equal(cell_sorted[b], want)
        seg = torch.repeat_interleave(torch.arange(H * W, device=dev), count[b].long())
        assert torch.equal(c[:nv].long(), seg)


@pytest.mark.parametrize("B,N", [(2, 4000), (3, 20000)])
def test_cell_sorted_projection_equals_the_indexed_one(B, N):
    """Cell-sorted rows: the contiguous reduce gives the same grids as the indexed one on the unsorted rows, and
    (share, bits) encode exactly the gradient rows kdf_bev_bwd_affine writes."""
    from src import ops, point_mlp
    dev = "cuda"
    pts = _frames(12, B, N, dev)
    g = torch.Generator(device="cpu").manual_seed(6)
    C, H, W = 128, 64, 64
    z = (torch.randn(B * N, C, generator=g) * 1.5).to(torch.bfloat16).to(dev)
    z[: (B * N) // 2] = z[(B * N) // 2: 2 * ((B * N) // 2)]                                 # duplicate rows: ties inside cells
    scale = (torch.randn(C, generator=g)).to(dev)
    shift = (torch.randn(C, generator=g) * 0.5).to(dev)
    geom = ops.bev_range_constants([-50, -50, -5, 50, 50, 3])
    cell0, count0, order0, offsets0 = point_mlp.bev_build_order(pts, geom, (H, W))
    grid0, gz0 = point_mlp.bev_reduce_affine(z, scale, shift, order0, offsets0, B, N, (H, W), True)
    cell, count, offsets, spts, cell_sorted, order = point_mlp.bev_build_sorted(pts, geom, (H, W), want_order=True)
    rows = (order.long() + torch.arange(B, device=dev)[:, None] * N).reshape(-1)            # sorted row -> original row
    zs = z[rows].contiguous()
    grid, gz = point_mlp.bev_reduce_affine(zs, scale, shift, None, offsets, B, N, (H, W), True)
    assert torch.equal(grid, grid0) and torch.equal(gz, gz0)
    gg = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).to(dev)
    dy0, sums0 = point_mlp.bev_bwd_affine(gg, z, grid0, gz0, order0, offsets0, cell0, B, N, (H, W))
    share, bits, sums = point_mlp.bev_bwd_share(gg, zs, grid, gz, offsets, B, N, (H, W))
    cs = cell_sorted.reshape(-1).long()
    inside = cs >= 0
    bit = ((bits[inside].long()[:, :, None] >> torch.arange(8, device=dev)) & 1).reshape(-1, C).bool()
    dy = torch.zeros(B * N, C, dtype=torch.bfloat16, device=dev)
    dy[inside] = torch.where(bit, share[cs[inside]], torch.zeros((), dtype=torch.bfloat16, device=dev))
    assert torch.equal(dy, dy0[rows])
    np.testing.assert_allclose(sums.cpu().numpy(), sums0.cpu().numpy(), rtol=1e-5, atol=1e-5)


def test_mlp_layer_bwd_share_equals_the_materialised_gradient_rows():
    """Layer-3 backward with dy formed from (cell_sorted, share, bits) in the prologue == the same kernel fed the
    materialised rows: dy_prev bit-identical, sums / dW to summation-order noise.  Rows outside hold garbage bits."""
    from src import point_mlp, ops
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(9)
    M, ncell = 128 * 37 + 50, 301
    n_in = M - 777
    cs = torch.sort(torch.randint(0, ncell, (n_in,), generator=g)).values.to(torch.int32)
    cs = torch.cat([cs, torch.full((M - n_in,), -1, dtype=torch.int32)]).to(dev)
    share = (torch.randn(ncell, 128, generator=g)).to(torch.bfloat16).to(dev)
    bits = torch.randint(0, 256, (M, 16), generator=g).to(torch.uint8).to(dev)
    z = (torch.randn(M, 128, generator=g) * 1.5).to(torch.bfloat16).to(dev)
    zprev = (torch.randn(M, 128, generator=g) * 2).to(torch.bfloat16).to(dev)
    gs, ga, gb = (torch.rand(128, generator=g) + 0.5).to(dev), (torch.randn(128, generator=g) * 0.01).to(dev), (torch.randn(128, generator=g) * 0.01).to(dev)
    scale, shift = (torch.rand(128, generator=g) + 0.5).to(dev), (torch.randn(128, generator=g) * 0.3).to(dev)
    Wt = (torch.randn(128, 128, generator=g) / 11.3).to(torch.bfloat16).to(dev)
    inside = cs >= 0
    bit = ((bits.long()[:, :, None] >> torch.arange(8, device=dev)) & 1).reshape(M, 128).bool()
    dy = torch.where(bit & inside[:, None], share[cs.clamp_min(0).long()], torch.zeros((), dtype=torch.bfloat16, device=dev))
    ref = ops.mlp_layer_bwd(1, dy, z, gs, ga, gb, zprev, scale, shift, Wt)
    got = point_mlp.mlp_layer_bwd_share(cs, share, bits, z, gs, ga, gb, zprev, scale, shift, Wt)
    assert torch.equal(got[0], ref[0])
    np.testing.assert_allclose(got[1].cpu().numpy(), ref[1].cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert _l2(got[2], ref[2]) < 1e-5
