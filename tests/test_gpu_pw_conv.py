"""kdf_pw_conv_fwd (the fused tcgen05 1x1-convolution layer) against plain fp32 PyTorch of the same op, at every
(K, N) the camera branch / FPN / fusion / head use and at row counts that do not fill the last tile."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(32, 32), (32, 192), (192, 64), (64, 384), (384, 64), (384, 128), (128, 768), (768, 128),
          (64, 128), (128, 128), (128, 64), (64, 32), (256, 256), (256, 64)]


def _act(v, code):
    return v if code == 0 else (torch.relu(v) if code == 1 else torch.clamp(v, 0.0, 6.0))


def _l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _inputs(M, K, N, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = (torch.randn(M, K, generator=g, device="cuda") * 1.5 + 0.3).to(torch.bfloat16)
    w = torch.randn(N, K, generator=g, device="cuda") * (2.0 / K) ** 0.5
    return x, w, g


@pytest.mark.parametrize("K,N", SHAPES)
@pytest.mark.parametrize("M", [128 * 37 + 78, 2048])
def test_pw_conv_train_mode(K, N, M):
    """raw output rows + batch statistics, with and without the BatchNorm+ReLU6 prologue.  bf16 operands, fp32
    accumulation: rows within 1e-2 relative L2 of the fp32 product of the SAME bf16 operands rounded once (bf16
    output rounding, 2^-9), statistics = sums of exactly the stored rows (1e-6)."""
    from src import ops
    if not ops.pw_conv_supported(K, N, M):
        pytest.skip("shape not supported")
    pack = ops._pw_pack_factor(K, N)
    x, w, g = _inputs(M, K, N, K * 1000 + N)
    wb = ops.pw_conv_weight(w, pack)
    sc = torch.rand(K, generator=g, device="cuda") + 0.5
    sh = torch.randn(K, generator=g, device="cuda") * 0.5
    wf = w.to(torch.bfloat16).float()
    for pro in (None, (sc, sh, 2), (sc, sh, 1)):
        out, stats = ops.pw_conv_fwd(x, wb, pack, pro=pro, want_stats=True)
        a = x.float()
        if pro is not None:
            a = _act(a * sc + sh, pro[2]).to(torch.bfloat16).float()        # the operand tile is rounded to bf16
        ref = a @ wf.t()
        assert out.shape == (M, N) and out.dtype == torch.bfloat16
        assert _l2(out.float(), ref) < 1e-2, (K, N, pro is not None, _l2(out.float(), ref))
        assert (out.float() - ref).abs().max().item() < 0.05 * ref.abs().max().item()
        o = out.double()
        assert torch.allclose(stats[0], o.sum(0), rtol=1e-6, atol=1e-3), (K, N)
        assert torch.allclose(stats[1], (o * o).sum(0), rtol=1e-6, atol=1e-3), (K, N)
        plain = ops.pw_conv_fwd(x, wb, pack, pro=pro)
        assert torch.equal(plain, out)                                      # statistics do not change the rows


@pytest.mark.parametrize("K,N", SHAPES)
def test_pw_conv_eval_mode(K, N):
    """folded BatchNorm + activation (+ shortcut) in the epilogue: the inference (teacher) path."""
    from src import ops
    M = 128 * 21 + 4
    if not ops.pw_conv_supported(K, N, M):
        pytest.skip("shape not supported")
    pack = ops._pw_pack_factor(K, N)
    x, w, g = _inputs(M, K, N, K * 999 + N)
    wb = ops.pw_conv_weight(w, pack)
    es = torch.rand(N, generator=g, device="cuda") + 0.5
    eh = torch.randn(N, generator=g, device="cuda") * 0.3
    res = torch.randn(M, N, generator=g, device="cuda").to(torch.bfloat16)
    z = x.float() @ w.to(torch.bfloat16).float().t()
    for act in (0, 1, 2):
        for r in (None, res):
            out = ops.pw_conv_fwd(x, wb, pack, epi=(es, eh, act), residual=r)
            ref = _act(z * es + eh, act)
            if r is not None:
                ref = ref + r.float()
            assert _l2(out.float(), ref) < 1e-2, (K, N, act, r is not None, _l2(out.float(), ref))


def test_pw_conv_small_and_empty():
    from src import ops
    w = torch.randn(64, 64, device="cuda")
    wb = ops.pw_conv_weight(w)
    for M in (0, 1, 7, 128, 129):
        x = torch.randn(M, 64, device="cuda").to(torch.bfloat16)
        out, stats = ops.pw_conv_fwd(x, wb, want_stats=True)
        ref = x.float() @ w.to(torch.bfloat16).float().t()
        assert out.shape == (M, 64)
        if M:
            assert _l2(out.float(), ref) < 1e-2
        assert torch.allclose(stats[0], out.double().sum(0), rtol=1e-6, atol=1e-4)
    with pytest.raises(RuntimeError, match="multiple of 64"):
        ops.pw_conv_fwd(torch.zeros(4, 48, device="cuda", dtype=torch.bfloat16), torch.zeros(32, 48, device="cuda", dtype=torch.bfloat16))
