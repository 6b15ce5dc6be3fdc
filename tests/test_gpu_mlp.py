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
