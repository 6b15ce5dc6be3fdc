"""GPU tests of the fused tcgen05 point-MLP kernels against plain torch arithmetic on the same bf16-rounded
operands (the MMA multiplies bf16 x bf16 exactly and accumulates in fp32)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


@pytest.mark.parametrize("M", [128, 1000, 128 * 148 * 3 + 77])
def test_mlp_layer_fwd_mode1_affine_relu_gemm(M):
    from src import ops
    g = torch.Generator().manual_seed(M)
    zprev = (torch.randn(M, 128, generator=g) * 2).to(torch.bfloat16)
    scale = torch.rand(128, generator=g) + 0.5
    shift = torch.randn(128, generator=g) * 0.3
    W = (torch.randn(128, 128, generator=g) / 11.3).to(torch.bfloat16)
    z, stats = ops.mlp_layer_fwd(1, zprev.cuda(), scale.cuda(), shift.cuda(), W.cuda())
    a32 = torch.relu(zprev.float() * scale + shift)
    a = a32.to(torch.bfloat16).float()
    ref = a @ W.float().t()
    assert z.dtype == torch.bfloat16 and z.shape == (M, 128)
    err = (z.float().cpu() - ref).abs()
    # output rounding (2^-9 |z|) + the kernel forms z*scale+shift with one FMA rounding instead of two, so an
    # activation within rounding of a bf16 boundary may land one bf16 ulp away (<= 2^-8 |a| |W| per term)
    tol = ref.abs() * 2 ** -8 + (a32.abs() @ W.float().abs().t()) * 2 ** -8 + 1e-3
    assert (err <= tol).all(), (err / tol).max()
    assert (err.mean() / ref.abs().mean()).item() < 2e-3
    zc = z.float().cpu().double()
    np.testing.assert_allclose(stats[0].cpu().numpy(), zc.sum(0).numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(stats[1].cpu().numpy(), (zc * zc).sum(0).numpy(), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("M", [64, 5000, 128 * 148 * 2 + 5])
def test_mlp_layer_fwd_mode0_first_layer_recomputed(M):
    from src import ops
    g = torch.Generator().manual_seed(M + 1)
    pts = torch.randn(M, 4, generator=g) * torch.tensor([40.0, 40.0, 2.0, 70.0])
    q = torch.randn(64, 4, generator=g) * 0.02
    r = torch.randn(64, generator=g) * 0.3
    W = (torch.randn(128, 64, generator=g) / 8).to(torch.bfloat16)
    z, stats = ops.mlp_layer_fwd(0, pts.cuda(), q.cuda(), r.cuda(), W.cuda())
    a1 = torch.relu(pts @ q.t() + r)
    ref = a1.to(torch.bfloat16).float() @ W.float().t()
    err = (z.float().cpu() - ref).abs()
    # a1 is formed with fp32 FMAs in a different order than the matmul above: a value within rounding of a
    # bf16 boundary may round the other way, which moves z by <= 2^-8 * |a1| * |W|
    tol = ref.abs() * 2 ** -8 + (a1.abs() @ W.float().abs().t()) * 2 ** -8 + 1e-3
    assert (err <= tol).all(), (err / tol).max()
    zc = z.float().cpu().double()
    np.testing.assert_allclose(stats[0].cpu().numpy(), zc.sum(0).numpy(), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("M", [1, 129, 5000, 128 * 148 * 3 + 17, 32 * 170_000])
def test_mlp_eval3_bit_identical_to_layer_kernels(M):
    """The one-kernel eval-mode MLP (running statistics; z2 stays on the SM) against kdf_mlp_layer_fwd mode 0
    followed by mode 1: the same operations in the same order, so the pre-BatchNorm-3 rows must be bit-identical."""
    import subprocess, sys, os, textwrap
    # runs in its own interpreter: the largest case allocates 8 GB that the test process would otherwise keep cached
    code = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
        sys.path.insert(0, {os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200")!r})
        from src import point_mlp
        M = {M}
        g = torch.Generator(device="cuda").manual_seed(M % 1000)
        pts = torch.randn(M, 4, generator=g, device="cuda") * torch.tensor([40.0, 40.0, 2.0, 70.0], device="cuda")
        q = torch.randn(64, 4, generator=g, device="cuda") * 0.02
        r = torch.randn(64, generator=g, device="cuda") * 0.3
        W2 = (torch.randn(128, 64, generator=g, device="cuda") / 8).to(torch.bfloat16)
        W3 = (torch.randn(128, 128, generator=g, device="cuda") / 11.3).to(torch.bfloat16)
        sc = torch.rand(128, generator=g, device="cuda") + 0.5
        sc[::7] *= -1
        sh = torch.randn(128, generator=g, device="cuda") * 0.3
        z2, _ = point_mlp.mlp_layer_fwd_raw(0, pts, q, r, W2)
        z3, _ = point_mlp.mlp_layer_fwd_raw(1, z2, sc, sh, W3)
        got = point_mlp.mlp_eval3_fwd(pts, q, r, W2, sc, sh, W3)
        torch.cuda.synchronize()
        assert got.shape == z3.shape and torch.equal(got.view(torch.int16), z3.view(torch.int16)), (got.float() - z3.float()).abs().max()
        assert z3.float().abs().max() > 0
        print("ok")
    """)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def _bwd_inputs(M, seed):
    g = torch.Generator().manual_seed(seed)
    dy = (torch.randn(M, 128, generator=g) * (torch.rand(M, 128, generator=g) < 0.1)).to(torch.bfloat16)
    z = (torch.randn(M, 128, generator=g) * 1.5).to(torch.bfloat16)
    gs = torch.rand(128, generator=g) + 0.5
    ga = torch.randn(128, generator=g) * 0.01
    gb = torch.randn(128, generator=g) * 0.01
    return g, dy, z, gs, ga, gb


@pytest.mark.parametrize("M", [128, 1000, 128 * 148 * 2 + 77])
def test_mlp_layer_bwd_mode1(M):
    """dgrad + wgrad + BatchNorm-backward prologue + ReLU-mask/column-sum epilogue of the 128->128 layer."""
    from src import ops
    g, dy, z, gs, ga, gb = _bwd_inputs(M, M)
    zprev = (torch.randn(M, 128, generator=g) * 2).to(torch.bfloat16)
    scale = torch.rand(128, generator=g) + 0.5
    shift = torch.randn(128, generator=g) * 0.3
    W = (torch.randn(128, 128, generator=g) / 11.3).to(torch.bfloat16)
    c = lambda t: t.cuda()
    dyp, sums, dW = ops.mlp_layer_bwd(1, c(dy), c(z), c(gs), c(ga), c(gb), c(zprev), c(scale), c(shift), c(W))
    dz = (gs * dy.float() + ga + gb * z.float()).to(torch.bfloat16).double()
    a = torch.relu(zprev.float() * scale + shift).to(torch.bfloat16).double()
    dA = dz @ W.double()
    ref = dA * (a > 0)
    got = dyp.float().cpu().double()
    # a handful of elements may differ by the one-FMA-vs-two rounding of dz / a at a bf16 boundary
    tol = ref.abs() * 2 ** -7 + (dz.abs() @ W.double().abs()) * 2 ** -8 + 1e-3
    assert ((got - ref).abs() <= tol).all(), ((got - ref).abs() / tol).max()
    assert ((got - ref).abs().mean() / ref.abs().mean()).item() < 4e-3
    np.testing.assert_allclose(sums[0].cpu().numpy(), got.sum(0).numpy(), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(sums[1].cpu().numpy(), (got * zprev.double()).sum(0).numpy(), rtol=1e-4, atol=1e-2)
    ref_dW = dz.t() @ a
    assert rel_err(dW.cpu(), ref_dW) < 2e-3, rel_err(dW.cpu(), ref_dW)


@pytest.mark.parametrize("M", [64, 5000, 128 * 148 * 2 + 5])
def test_mlp_layer_bwd_mode0(M):
    """64->128 layer backwards with the first layer recomputed from the raw points; only 64x5 sums leave the kernel."""
    from src import ops
    g, dy, z, gs, ga, gb = _bwd_inputs(M, M + 7)
    pts = torch.randn(M, 4, generator=g) * torch.tensor([40.0, 40.0, 2.0, 70.0])
    q = torch.randn(64, 4, generator=g) * 0.02
    r = torch.randn(64, generator=g) * 0.3
    W = (torch.randn(128, 64, generator=g) / 8).to(torch.bfloat16)
    c = lambda t: t.cuda()
    none, sums, dW = ops.mlp_layer_bwd(0, c(dy), c(z), c(gs), c(ga), c(gb), c(pts), c(q), c(r), c(W))
    assert none is None
    dz = (gs * dy.float() + ga + gb * z.float()).to(torch.bfloat16).double()
    a1 = torch.relu(pts @ q.t() + r).to(torch.bfloat16).double()
    dy1 = (dz @ W.double()) * (a1 > 0)
    ref = torch.cat([dy1.sum(0, keepdim=True), dy1.t() @ pts.double()[:, :4]]).reshape(5, 64) if False else \
        torch.stack([dy1.sum(0)] + [(dy1 * pts.double()[:, d:d + 1]).sum(0) for d in range(4)])
    scale = torch.stack([dy1.abs().sum(0)] + [(dy1.abs() * pts.double()[:, d:d + 1].abs()).sum(0) for d in range(4)])
    err = (sums.cpu() - ref).abs() / (scale + 1e-6)
    # the TMA-staged kernel feeds the raw dy / z tiles to the tensor cores with the BatchNorm backward folded into the
    # weights (nothing is re-rounded, unlike `dz` above) and stages dy1 in bf16: a few 1e-3 of the sum of magnitudes
    assert err.max().item() < 8e-3, err.max()
    ref_dW = dz.t() @ a1
    assert rel_err(dW.cpu(), ref_dW) < 2e-3, rel_err(dW.cpu(), ref_dW)


def test_mlp_layer_bwd_row_mask_ignores_unwritten_rows():
    """Rows flagged by row_cell < 0 (points outside the grid) count as dy == 0 whatever their memory holds
    (kdf_bev_bwd_affine with cell = NULL leaves them unwritten)."""
    from src import ops
    M = 128 * 5 + 9
    g, dy, z, gs, ga, gb = _bwd_inputs(M, 77)
    zprev = (torch.randn(M, 128, generator=g) * 2).to(torch.bfloat16)
    scale, shift = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.3
    W = (torch.randn(128, 128, generator=g) / 11.3).to(torch.bfloat16)
    cell = torch.randint(-1, 5, (M,), generator=g).to(torch.int32)
    dy_clean = torch.where((cell >= 0)[:, None], dy, torch.zeros((), dtype=dy.dtype))
    dy_dirty = torch.where((cell >= 0)[:, None], dy, torch.full((), float("nan"), dtype=dy.dtype))
    c = lambda t: t.cuda()
    ref = ops.mlp_layer_bwd(1, c(dy_clean), c(z), c(gs), c(ga), c(gb), c(zprev), c(scale), c(shift), c(W))
    got = ops.mlp_layer_bwd(1, c(dy_dirty), c(z), c(gs), c(ga), c(gb), c(zprev), c(scale), c(shift), c(W), row_cell=c(cell))
    assert torch.equal(got[0], ref[0])
    np.testing.assert_allclose(got[1].cpu().numpy(), ref[1].cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert rel_err(got[2].cpu(), ref[2].cpu()) < 1e-5
