"""torch-facing wrappers (and autograd glue) over the kdfusion_b200 C ABI.

Everything here only moves pointers: tensors are allocated by torch on the
current CUDA stream, the arithmetic happens in libkdfusion_b200.so.  CUDA tensors
only -- there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import native as _n
from .native import call, dtype_code, lib, ptr, require_cuda, stream_ptr

__all__ = [
    "bn_act", "run_fused", "fpn_merge",
    "bev_range_constants", "bev_index", "bev_project", "BevProjectFn", "range_index", "range_project",
    "pw_conv_fwd", "fused_fusion", "kd_loss_fwd_bwd", "KDLossFn", "confusion_matrix_", "adamw_flat_",
]


# ----------------------------------------------------------------------------- BatchNorm (+act) over rows
_ACT = {None: 0, "none": 0, "relu": 1, "relu6": 2}
class _RowBNActFn(torch.autograd.Function):
    """y = act(x*scale + shift) [+ residual] over rows [M,C]; scale/shift/mean/invstd are the fp32
    per-channel vectors prepared by ``bn_act`` (batch statistics in training, running statistics in
    eval).  Backward = kdf_rowbn_bwd (reduce + apply, chaining through the batch statistics)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, scale, shift, mean, invstd, act, batch_stats, residual, pre_bias):
        M, C = x.shape
        y = torch.empty_like(x)
        call("kdf_rowbn_apply_fwd", ptr(x), ptr(residual), dtype_code(x), M, C, ptr(scale), ptr(shift), act, ptr(y),
             stream_ptr(x.device))
        ctx.save_for_backward(x, scale, shift, mean, invstd)
        ctx.act, ctx.batch_stats = act, batch_stats
        ctx.has_res = residual is not None
        ctx.affine = (gamma is not None, beta is not None, pre_bias is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, scale, shift, mean, invstd = ctx.saved_tensors
        M, C = x.shape
        g = g.contiguous()
        if g.dtype != x.dtype:
            g = g.to(x.dtype)
        dx = torch.empty_like(x)
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
        nb = lib.kdf_rowbn_bwd_workspace_bytes(C)
        ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
        call("kdf_rowbn_bwd", ptr(g), ptr(x), dtype_code(x), M, C, ptr(scale), ptr(shift), ptr(mean), ptr(invstd),
             ctx.act, int(ctx.batch_stats), ptr(dx), ptr(dgamma), ptr(dbeta), ptr(ws), stream_ptr(x.device))
        dbias = None
        if ctx.affine[2]:
            # BN(x + b): batch statistics cancel b exactly (zero gradient); with running statistics
            # b acts like a shift in front of the scale
            dbias = torch.zeros_like(dbeta) if ctx.batch_stats else dbeta * scale
        return (dx, dgamma if ctx.affine[0] else None, dbeta if ctx.affine[1] else None,
                None, None, None, None, None, None, g if ctx.has_res else None, dbias)


def _eval_affine(bn, pre_bias=None):
    """(scale, shift, mean, invstd) of a BatchNorm in eval mode, cached until its tensors change.
    ``pre_bias`` (the producing layer's bias, applied as BN(x + bias)) is folded into the shift."""
    key = (_n.cache_generation(bn),) + tuple((t.data_ptr(), t._version) for t in
                                             (bn.weight, bn.bias, bn.running_mean, bn.running_var, pre_bias) if t is not None)
    cache = getattr(bn, "_kdf_eval_cache", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    with torch.no_grad():
        mean = bn.running_mean.float()
        invstd = torch.rsqrt(bn.running_var.float() + bn.eps)
        scale = invstd * bn.weight.float() if bn.weight is not None else invstd.clone()
        shift = (bn.bias.float() if bn.bias is not None else torch.zeros_like(mean)) - mean * scale
        if pre_bias is not None:
            shift = shift + pre_bias.float() * scale
        vals = tuple(t.contiguous() for t in (scale, shift, mean, invstd))
    bn._kdf_eval_cache = (key, vals)
    return vals


def _bn_prepare(rows: torch.Tensor, bn, pre_bias=None):
    """fp32 per-channel (scale, shift, mean, invstd, batch_stats) for a BatchNorm module over rows [M,C]:
    batch statistics (one ``kdf_rowbn_stats`` launch, which also advances the running statistics
    exactly like nn.BatchNorm) in training, cached running-statistics affine in eval."""
    dev = rows.device
    M, C = rows.shape
    use_batch = bn.training or bn.running_mean is None
    if not use_batch:
        return (*_eval_affine(bn, pre_bias), False)
    f32 = dict(dtype=torch.float32, device=dev)
    mean, invstd, scale, shift = (torch.empty(C, **f32) for _ in range(4))
    track = bn.training and bn.track_running_stats and bn.running_mean is not None
    mom = 0.0
    if track:
        mom = _n.bump_batch_counter(bn)
    ws = torch.empty(lib.kdf_rowbn_workspace_bytes(C), dtype=torch.uint8, device=dev)
    with torch.no_grad():
        call("kdf_rowbn_stats", ptr(rows), dtype_code(rows), M, C, ptr(bn.weight), ptr(bn.bias),
             ptr(pre_bias.detach().float()) if pre_bias is not None else None, float(bn.eps), float(mom),
             ptr(bn.running_mean) if track else None, ptr(bn.running_var) if track else None,
             ptr(mean), ptr(invstd), ptr(scale), ptr(shift), ptr(ws), stream_ptr(dev))
    return scale, shift, mean, invstd, True


def bn_act(x: torch.Tensor, bn, act: Optional[str] = None, residual: Optional[torch.Tensor] = None,
           pre_bias: Optional[torch.Tensor] = None, col_sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``act(bn(x [+ pre_bias])) [+ residual]`` with the BatchNorm module's parameters / buffers (updated
    like nn.BatchNorm does in training).  ``x`` is a channels-last [B,C,H,W] tensor or rows [M,C].
    ``pre_bias`` is the bias of the layer that produced ``x``: instead of adding it to every element
    (and reducing its gradient over all rows) it is folded into the normalisation, which in training
    cancels it exactly."""
    dev = require_cuda(x, residual)
    four_d = x.dim() == 4
    if four_d:
        B, C, H, W = x.shape
        rows = x.permute(0, 2, 3, 1).reshape(B * H * W, C)        # free for channels-last, one copy otherwise
        res = None if residual is None else residual.permute(0, 2, 3, 1).reshape(B * H * W, C)
    elif x.dim() == 2:
        rows, res, C = x.contiguous(), (None if residual is None else residual.contiguous()), x.shape[1]
    else:
        raise ValueError(f"bn_act expects [B,C,H,W] or [M,C], got {tuple(x.shape)}")
    if not rows.is_contiguous():
        rows = rows.contiguous()
    if res is not None and (res.dtype != rows.dtype or not res.is_contiguous()):
        res = res.to(rows.dtype).contiguous()
    if col_sums is not None and (bn.training or bn.running_mean is None):
        # the producing kernel already reduced sum / sum of squares of these rows: finalise only (no statistics pass)
        from .point_mlp import bn_finalize
        track = bn.training and bn.track_running_stats and bn.running_mean is not None
        with torch.no_grad():
            mean, invstd, scale, shift = bn_finalize(col_sums, rows.shape[0], bn,
                                                     pre_bias.detach().float().contiguous() if pre_bias is not None else None, track)
        use_batch = True
    else:
        scale, shift, mean, invstd, use_batch = _bn_prepare(rows, bn, pre_bias)
    y = _RowBNActFn.apply(rows, bn.weight, bn.bias, scale, shift, mean, invstd, _ACT[act], use_batch, res, pre_bias)
    return y.view(B, H, W, C).permute(0, 3, 1, 2) if four_d else y


# ----------------------------------------------------------------------------- depthwise 3x3 convolution
def _nhwc_rows(x: torch.Tensor) -> Optional[torch.Tensor]:
    """[B,C,H,W] -> its [B,H,W,C] memory if x is dense channels-last (no copy), else None."""
    if x.dim() != 4:
        return None
    y = x.permute(0, 2, 3, 1)
    return y if y.is_contiguous() else None


class _DwConv3x3Fn(torch.autograd.Function):
    """nn.Conv2d(C, C, 3, stride, padding=1, groups=C, bias=False) over a dense channels-last map."""

    @staticmethod
    def forward(ctx, x, weight, stride, want_stats):
        B, C, H, W = x.shape
        OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
        w = weight.detach().reshape(C, 9).float().contiguous()
        out = torch.empty(B, OH, OW, C, dtype=x.dtype, device=x.device)
        stats = torch.empty(2, C, dtype=torch.float64, device=x.device) if want_stats else None
        call("kdf_dwconv3x3_fwd", ptr(_nhwc_rows(x)), ptr(w), dtype_code(x), B, H, W, C, stride, 0, ptr(out), ptr(stats),
             stream_ptr(x.device))
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.wshape, ctx.wdtype = stride, weight.shape, weight.dtype
        ctx.set_materialize_grads(False)        # no zero tensor for the statistics output's (absent) gradient: a launch per layer
        if want_stats:
            ctx.mark_non_differentiable(stats)
        return out.permute(0, 3, 1, 2), stats

    @staticmethod
    def backward(ctx, g, _gs=None):
        if g is None:
            return None, None, None, None
        x, w = ctx.saved_tensors
        B, C, H, W = x.shape
        if g.dtype != x.dtype:
            g = g.to(x.dtype)
        gr = _nhwc_rows(g)
        if gr is None:
            g = g.contiguous(memory_format=torch.channels_last)
            gr = g.permute(0, 2, 3, 1)
        dev, st = x.device, stream_ptr(x.device)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty(B, H, W, C, dtype=x.dtype, device=dev)
            call("kdf_dwconv3x3_bwd_data", ptr(gr), ptr(w), dtype_code(x), B, H, W, C, ctx.stride, ptr(gx), st)
            gx = gx.permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            gw = torch.empty(C, 9, dtype=torch.float32, device=dev)
            call("kdf_dwconv3x3_bwd_weight", ptr(_nhwc_rows(x)), ptr(gr), dtype_code(x), B, H, W, C, ctx.stride, ptr(gw), st)
            gw = gw.view(ctx.wshape).to(ctx.wdtype)
        return gx, gw, None, None


def dwconv3x3(conv, x: torch.Tensor, want_stats: bool = False, post=None):
    """Depthwise 3x3 ``nn.Conv2d`` on the hand-written kernels when it is one (3x3, padding 1, stride 1|2, no bias,
    groups == channels) and ``x`` is a dense channels-last CUDA map; None otherwise (the caller runs the module).
    With ``want_stats`` returns (out, stats f64 [2,C]): the column sums the BatchNorm that follows needs.
    ``post`` = (scale, shift, act code): the eval-mode BatchNorm + activation that follow, applied in the same kernel
    (inference only: no autograd through it)."""
    import torch.nn as nn
    if not (isinstance(conv, nn.Conv2d) and x.is_cuda and x.dim() == 4):
        return None
    C = conv.in_channels
    if not (conv.groups == C == conv.out_channels and C > 1 and conv.kernel_size == (3, 3) and conv.padding == (1, 1)
            and conv.dilation == (1, 1) and conv.stride in ((1, 1), (2, 2)) and conv.bias is None
            and conv.padding_mode == "zeros"):
        return None
    if torch.is_autocast_enabled() and x.dtype == torch.float32:
        x = x.to(torch.get_autocast_dtype("cuda"))
    if x.dtype not in (torch.float32, torch.bfloat16) or C % 8 or C > 1024 or _nhwc_rows(x) is None:
        return None
    if post is not None:
        scale, shift, act = post
        B, _, H, W = x.shape
        stride = conv.stride[0]
        OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
        w = conv.weight.detach().reshape(C, 9).float().contiguous()
        out = torch.empty(B, OH, OW, C, dtype=x.dtype, device=x.device)
        call("kdf_dwconv3x3_affine_fwd", ptr(_nhwc_rows(x)), ptr(w), dtype_code(x), B, H, W, C, stride, ptr(scale), ptr(shift),
             act, ptr(out), stream_ptr(x.device))
        return out.permute(0, 3, 1, 2)
    out, stats = _DwConv3x3Fn.apply(x, conv.weight, conv.stride[0], want_stats)
    return (out, stats) if want_stats else out


# ----------------------------------------------------------------------------- the camera stem
class _StemConvFn(torch.autograd.Function):
    """nn.Conv2d(3, 32, 3, stride 2, padding 1, bias=False) from the fp32 NCHW image to bf16 channels-last rows, with the
    column sums of the BatchNorm that follows (camera_encoder.py:63-67); backward = the weight gradient only."""

    @staticmethod
    def forward(ctx, x, weight):
        B, _, H, W = x.shape
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        w = weight.detach().float().contiguous()
        out = torch.empty(B, OH, OW, 32, dtype=torch.bfloat16, device=x.device)
        stats = torch.empty(2, 32, dtype=torch.float64, device=x.device)
        call("kdf_stem_conv_fwd", ptr(x), ptr(w), B, H, W, None, None, 0, ptr(out), ptr(stats), stream_ptr(x.device))
        ctx.save_for_backward(x)
        ctx.wshape, ctx.wdtype = weight.shape, weight.dtype
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(stats)
        return out.permute(0, 3, 1, 2), stats

    @staticmethod
    def backward(ctx, g, _gs=None):
        if g is None or not ctx.needs_input_grad[1]:
            return None, None
        (x,) = ctx.saved_tensors
        B, _, H, W = x.shape
        if g.dtype != torch.bfloat16:
            g = g.to(torch.bfloat16)
        gr = _nhwc_rows(g)
        if gr is None:
            gr = g.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)
        gw = torch.empty(32, 27, dtype=torch.float32, device=x.device)
        call("kdf_stem_conv_bwd_weight", ptr(x), ptr(gr), B, H, W, ptr(gw), stream_ptr(x.device))
        return None, gw.view(ctx.wshape).to(ctx.wdtype)


def stem_conv(seq, x: torch.Tensor) -> Optional[torch.Tensor]:
    """The camera stem ``Sequential(Conv2d(3, 32, 3, s2, p1, bias=False), BatchNorm2d, ReLU6)`` on the stem kernel when the
    image is a dense fp32 NCHW CUDA tensor under bf16 autocast (what the loader delivers); None otherwise (the caller then
    takes the library path).  Training: rows + batch statistics from the kernel, BatchNorm + ReLU6 by the row kernel;
    inference: convolution + folded BatchNorm + ReLU6 in one kernel."""
    import torch.nn as nn
    mods = list(seq)
    if not (2 <= len(mods) <= 3 and isinstance(mods[0], nn.Conv2d) and isinstance(mods[1], nn.BatchNorm2d)):
        return None
    conv, bn = mods[0], mods[1]
    act = None
    if len(mods) == 3:
        if not isinstance(mods[2], (nn.ReLU, nn.ReLU6)):
            return None
        act = "relu6" if isinstance(mods[2], nn.ReLU6) else "relu"
    if not (conv.in_channels == 3 and conv.out_channels == 32 and conv.kernel_size == (3, 3) and conv.stride == (2, 2)
            and conv.padding == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1 and conv.bias is None
            and conv.padding_mode == "zeros" and conv.weight.dtype == torch.float32 and bn.affine):
        return None
    if not (x.is_cuda and x.dim() == 4 and x.shape[1] == 3 and x.dtype == torch.float32 and x.is_contiguous()
            and not x.requires_grad and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16):
        return None
    batch = bn.training or bn.running_mean is None
    if batch:
        z, stats = _StemConvFn.apply(x, conv.weight)
        return bn_act(z, bn, act, None, col_sums=stats)
    if torch.is_grad_enabled() and (conv.weight.requires_grad or bn.weight.requires_grad):
        return None                                              # eval-mode statistics with gradients: the library path
    B, _, H, W = x.shape
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    scale, shift, _, _ = _eval_affine(bn)
    out = torch.empty(B, OH, OW, 32, dtype=torch.bfloat16, device=x.device)
    call("kdf_stem_conv_fwd", ptr(x), ptr(conv.weight.detach().contiguous()), B, H, W, ptr(scale), ptr(shift), _ACT[act], ptr(out), None,
         stream_ptr(x.device))
    return out.permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------- the head's classifier
class _ClsConvFn(torch.autograd.Function):
    """nn.Conv2d(32, K, 1) (+ bias) over channels-last bf16 maps -> planar bf16 logits [B,K,H,W]; backward in one kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        B, C, H, W = x.shape
        K = weight.shape[0]
        rows = _nhwc_rows(x)
        w = weight.detach().reshape(K, C).float().contiguous()
        out = torch.empty(B, K, H, W, dtype=torch.bfloat16, device=x.device)
        call("kdf_cls_conv_fwd", ptr(rows), ptr(w), ptr(bias.detach().float().contiguous()) if bias is not None else None,
             B * H * W, C, K, H * W, ptr(out), stream_ptr(x.device))
        ctx.save_for_backward(x, w)
        ctx.wshape, ctx.has_bias = weight.shape, bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        B, C, H, W = x.shape
        K = w.shape[0]
        g = g.to(torch.bfloat16).contiguous()
        dev = x.device
        gx = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev) if ctx.needs_input_grad[0] else None
        gw = torch.empty(K, C, dtype=torch.float32, device=dev)
        gb = torch.empty(K, dtype=torch.float32, device=dev) if ctx.has_bias else None
        call("kdf_cls_conv_bwd", ptr(_nhwc_rows(x)), ptr(g), ptr(w), B * H * W, C, K, H * W, ptr(gx), ptr(gw), ptr(gb), stream_ptr(dev))
        return (gx.permute(0, 3, 1, 2) if gx is not None else None), gw.view(ctx.wshape), gb


def cls_conv(conv, x: torch.Tensor) -> Optional[torch.Tensor]:
    """The head's 1x1 classifier ``nn.Conv2d(32, K <= 4, 1)`` on the classifier kernel for dense channels-last bf16 CUDA maps
    under bf16 autocast (planar bf16 logits, as the autocast convolution returns them in value); None otherwise."""
    import torch.nn as nn
    if not (isinstance(conv, nn.Conv2d) and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.padding == (0, 0)
            and conv.groups == 1 and conv.in_channels == 32 and 1 <= conv.out_channels <= 4 and conv.weight.dtype == torch.float32):
        return None
    if not (x.is_cuda and x.dim() == 4 and x.dtype == torch.bfloat16 and _nhwc_rows(x) is not None
            and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16):
        return None
    return _ClsConvFn.apply(x, conv.weight, conv.bias)


def _conv_frozen(m, x: torch.Tensor) -> torch.Tensor:
    """``m(x)``; for a convolution run without autograd under bf16 autocast (the frozen teacher) the bf16 copy of its
    weight is cached until the weight changes, instead of being re-cast by autocast on every step."""
    import torch.nn as nn
    import torch.nn.functional as F
    if not (isinstance(m, nn.Conv2d) and x.is_cuda and not torch.is_grad_enabled() and torch.is_autocast_enabled()
            and torch.get_autocast_dtype("cuda") == torch.bfloat16 and m.weight.dtype == torch.float32
            and m.padding_mode == "zeros" and not isinstance(m.padding, str)):
        return m(x)
    key = (_n.cache_generation(m), m.weight.data_ptr(), m.weight._version, None if m.bias is None else m.bias._version)
    cache = getattr(m, "_kdf_w16", None)
    if cache is None or cache[0] != key:
        cache = (key, m.weight.detach().to(torch.bfloat16), None if m.bias is None else m.bias.detach().to(torch.bfloat16))
        m._kdf_w16 = cache
    return F.conv2d(x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16), cache[1], cache[2], m.stride, m.padding,
                    m.dilation, m.groups)


def _is_pw_conv(m) -> bool:
    import torch.nn as nn
    return (isinstance(m, nn.Conv2d) and m.kernel_size == (1, 1) and m.stride == (1, 1) and m.groups == 1
            and m.dilation == (1, 1) and m.padding in ((0, 0), 0) and m.padding_mode == "zeros")


def _pw_weight_cached(conv, pack: int) -> torch.Tensor:
    """bf16 (block-diagonal when packed) operand of a 1x1 convolution.  A parameter owned by ``FlatAdamW`` has a bf16
    shadow that the optimizer kernel keeps current: the operand is a VIEW of it (no per-step cast); a packed layer
    copies the view into the two diagonal blocks of a persistent buffer.  Other weights (the frozen teacher) are cast
    once and cached until they change."""
    w = conv.weight
    shadow = getattr(w, "_kdf_shadow", None)
    if shadow is not None:
        N = w.shape[0]
        w16 = shadow().view(N, -1)
        if pack == 1:
            return w16
        K = w16.shape[1]
        buf = getattr(conv, "_kdf_pw_packed", None)
        if buf is None or buf.device != w16.device or tuple(buf.shape) != (N * pack, K * pack):
            buf = torch.zeros(N * pack, K * pack, dtype=torch.bfloat16, device=w16.device)
            conv._kdf_pw_packed = buf
        for i in range(pack):
            buf[i * N:(i + 1) * N, i * K:(i + 1) * K].copy_(w16)
        return buf
    key = (_n.cache_generation(conv), w.data_ptr(), w._version, pack)
    cache = getattr(conv, "_kdf_pw_w", None)
    if cache is None or cache[0] != key:
        cache = (key, pw_conv_weight(w, pack))
        conv._kdf_pw_w = cache
    return cache[1]


class _PwConvFn(torch.autograd.Function):
    """z = x . W^T over pixel rows on the fused layer kernel, with the batch statistics of z from its epilogue.
    Backward: data and weight gradients as plain library GEMMs over the same rows."""

    @staticmethod
    def forward(ctx, x_rows, weight, wb, pack):
        z, stats = pw_conv_fwd(x_rows, wb, pack, want_stats=True)
        ctx.save_for_backward(x_rows, wb)
        ctx.wshape = weight.shape
        ctx.set_materialize_grads(False)        # no zero tensor for the statistics output's (absent) gradient
        ctx.mark_non_differentiable(stats)
        return z, stats

    @staticmethod
    def backward(ctx, g, _gs):
        if g is None:
            return None, None, None, None
        x_rows, wb = ctx.saved_tensors
        g = g.contiguous()
        N, K = ctx.wshape[0], x_rows.shape[1]
        w16 = wb[:N, :K]                                      # the bf16 operand the forward used (first diagonal block when packed)
        gx = torch.mm(g, w16) if ctx.needs_input_grad[0] else None
        # fp32 straight out of the GEMM (aten::mm.dtype): no bf16 rounding of the weight gradient and no conversion launch
        gw = torch.mm(g.t(), x_rows, out_dtype=torch.float32).view(ctx.wshape) if ctx.needs_input_grad[1] else None
        return gx, gw, None, None


def _pw_usable(conv, x: torch.Tensor) -> bool:
    """The fused 1x1 layer serves dense channels-last bf16 CUDA maps (the autocast path)."""
    if not (_is_pw_conv(conv) and x.is_cuda and x.dim() == 4 and x.dtype == torch.bfloat16 and _nhwc_rows(x) is not None):
        return False
    B, C, H, W = x.shape
    return pw_conv_supported(conv.in_channels, conv.out_channels, B * H * W)


def pw_project_rows(conv, bn, rows: torch.Tensor):
    """rows [M,K] -> (pre-BatchNorm rows [M,N], column sums f64 [2,N] | None) of a bias-free 1x1 convolution whose
    BatchNorm ``bn`` follows in a consumer kernel: on the fused layer kernel (statistics from its epilogue) for bf16
    rows in training, as a library GEMM otherwise."""
    import torch.nn.functional as F
    M, K = rows.shape
    N = conv.out_channels
    if (rows.is_cuda and rows.dtype == torch.bfloat16 and conv.bias is None and _is_pw_conv(conv) and bn.training
            and torch.is_grad_enabled() and pw_conv_supported(K, N, M)):
        pack = _pw_pack_factor(K, N)
        return _PwConvFn.apply(rows.contiguous(), conv.weight, _pw_weight_cached(conv, pack), pack)
    return F.linear(rows, conv.weight.flatten(1), conv.bias), None


def run_fused(seq, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Run an ``nn.Sequential`` of conv / BatchNorm / ReLU(6) layers with every BatchNorm(+activation)
    group executed by the fused row kernels; ``residual`` is added after the LAST BatchNorm group.
    The Sequential keeps its layers (and state_dict keys); only the execution is fused.

    bf16 maps: a 1x1 convolution runs on ``kdf_pw_conv_fwd`` -- with the BatchNorm behind it folded into the kernel's
    epilogue in inference (running statistics, no autograd), or with that BatchNorm's batch statistics coming out of
    the epilogue in training (no separate statistics pass)."""
    import torch.nn as nn
    mods = list(seq)
    last_bn = max((i for i, m in enumerate(mods) if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d))), default=-1)
    i, col_sums = 0, None
    if torch.is_autocast_enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and _is_pw_conv(mods[0]):
        x = x.to(torch.get_autocast_dtype("cuda"))
    while i < len(mods):
        m = mods[i]
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            act, step = None, 1
            if i + 1 < len(mods) and isinstance(mods[i + 1], (nn.ReLU, nn.ReLU6)):
                act, step = ("relu6" if isinstance(mods[i + 1], nn.ReLU6) else "relu"), 2
            x = bn_act(x, m, act, residual if i == last_bn else None, col_sums=col_sums)
            col_sums = None
            i += step
        else:
            # depthwise 3x3 on the stencil kernels (which also reduce the statistics of a train-mode BatchNorm that
            # follows), 1x1 on the fused tensor-core layer kernel, everything else on the library
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            fuse_stats = isinstance(nxt, nn.BatchNorm2d) and (nxt.training or nxt.running_mean is None)
            inference = isinstance(nxt, nn.BatchNorm2d) and not fuse_stats and not torch.is_grad_enabled()
            act, step = None, 2
            if inference and i + 2 < len(mods) and isinstance(mods[i + 2], (nn.ReLU, nn.ReLU6)):
                act, step = ("relu6" if isinstance(mods[i + 2], nn.ReLU6) else "relu"), 3
            if inference and _pw_usable(m, x):
                # inference: conv + running-statistics BatchNorm (+ activation, + shortcut) = one kernel
                B, _, H, W = x.shape
                pack = _pw_pack_factor(m.in_channels, m.out_channels)
                scale, shift, _, _ = _eval_affine(nxt, m.bias)
                res = None
                if residual is not None and i + 1 == last_bn:
                    res = _nhwc_rows(residual)
                    res = None if res is None else res.reshape(B * H * W, -1)
                if residual is None or i + 1 != last_bn or (res is not None and res.dtype == torch.bfloat16):
                    rows = pw_conv_fwd(_nhwc_rows(x).reshape(B * H * W, -1), _pw_weight_cached(m, pack), pack,
                                       epi=(scale, shift, _ACT[act]), residual=res)
                    x = rows.view(B, H, W, -1).permute(0, 3, 1, 2)
                    i += step
                    continue
            if inference and not (residual is not None and i + 1 == last_bn):
                # inference: the running-statistics BatchNorm (+ activation) rides in the stencil kernel's epilogue
                scale, shift, _, _ = _eval_affine(nxt)
                y = dwconv3x3(m, x, post=(scale, shift, _ACT[act]))
                if y is not None:
                    x = y
                    i += step
                    continue
            if fuse_stats and m.bias is None and _pw_usable(m, x) and torch.is_grad_enabled():
                # training: rows + the batch statistics of the BatchNorm that follows, from one kernel
                B, _, H, W = x.shape
                pack = _pw_pack_factor(m.in_channels, m.out_channels)
                rows, col_sums = _PwConvFn.apply(_nhwc_rows(x).reshape(B * H * W, -1), m.weight, _pw_weight_cached(m, pack), pack)
                x = rows.view(B, H, W, -1).permute(0, 3, 1, 2)
                i += 1
                continue
            y = dwconv3x3(m, x, want_stats=fuse_stats)
            if y is None:
                x = _conv_frozen(m, x)
            elif fuse_stats:
                x, col_sums = y
            else:
                x = y
            i += 1
    if residual is not None and last_bn < 0:
        x = x + residual
    return x


# ----------------------------------------------------------------------------- (1) projection
def bev_range_constants(point_cloud_range: Sequence) -> Tuple[float, float, float, float]:
    """fp32 (x0, xspan, y0, yspan) with the reference's promotion rules: the range
    buffers are int64 when the entries are Python ints (lidar_encoder.py:38-39), so
    the span is an exact integer difference before it becomes fp32 (:47-48)."""
    out = []
    for lo, hi in ((point_cloud_range[0], point_cloud_range[3]), (point_cloud_range[1], point_cloud_range[4])):
        if isinstance(lo, int) and isinstance(hi, int):
            x0 = torch.tensor(lo, dtype=torch.float32)
            span = torch.tensor(hi - lo, dtype=torch.float32)
        else:
            x0 = torch.tensor(lo, dtype=torch.float32)
            span = torch.tensor(hi, dtype=torch.float32) - x0
        out += [float(x0), float(span)]
    return tuple(out)


def _check_points(points: torch.Tensor) -> Tuple[int, int, int]:
    require_cuda(points)
    if points.dim() != 3 or points.shape[-1] < 2:
        raise ValueError(f"points must be [B, N, >=2], got {tuple(points.shape)}")
    if points.dtype != torch.float32:
        raise TypeError("points must stay float32 (index math is fp32 by contract)")
    return points.shape


def bev_index(points: torch.Tensor, geom: Tuple[float, float, float, float], grid_size: Tuple[int, int],
              want_rank: bool = False):
    """Cell id per point (int32, -1 outside) and per-cell occupancy (int32 [B,H*W]).
    Bit-exact with SpatialLiDAREncoder.points_to_bev_coords + the truncation of
    forward_vectorized (lidar_encoder.py:42-55, 69-71)."""
    B, N, D = _check_points(points)
    points = points.contiguous()
    H, W = grid_size
    cell = torch.empty(B, N, dtype=torch.int32, device=points.device)
    count = torch.empty(B, H * W, dtype=torch.int32, device=points.device)
    rank = torch.empty(B, N, dtype=torch.int32, device=points.device) if want_rank else None
    call("kdf_bev_index", ptr(points), B, N, D, *geom, H, W, ptr(cell), ptr(rank), ptr(count),
                            stream_ptr(points.device))
    return (cell, count, rank) if want_rank else (cell, count)


class BevProjectFn(torch.autograd.Function):
    """grid[B,C,H,W] (NHWC memory) = per-cell max/mean of point-major feats[B,N,C]."""

    @staticmethod
    def forward(ctx, points, feats, geom, grid_size, reduce):
        B, N, D = _check_points(points)
        require_cuda(points, feats)
        if feats.dim() != 3 or feats.shape[0] != B or feats.shape[1] != N:
            raise ValueError(f"feats must be [B={B}, N={N}, C], got {tuple(feats.shape)}")
        points = points.contiguous()
        feats = feats.contiguous()
        C = feats.shape[2]
        H, W = grid_size
        dev = points.device
        grid = torch.empty(B, H, W, C, dtype=feats.dtype, device=dev)
        count = torch.empty(B, H * W, dtype=torch.int32, device=dev)
        cell = torch.empty(B, N, dtype=torch.int32, device=dev)
        need_grad = feats.requires_grad
        # rows of 8/16/32 16-byte lanes: pure-maximum forward, the backward counts the ties itself
        lanes, rem = divmod(C * feats.element_size(), 16)
        need_ties = reduce == _n.REDUCE_MAX and need_grad and not (rem == 0 and lanes in (8, 16, 32))
        ties = torch.empty(B, H * W, C, dtype=torch.int32, device=dev) if need_ties else None
        # the cell ordering is kept for the (cell-major) backward
        order = torch.empty(B, N, dtype=torch.int32, device=dev) if need_grad else None
        offsets = torch.empty(B, H * W + 1, dtype=torch.int32, device=dev) if need_grad else None
        ws_bytes = lib.kdf_bev_workspace_bytes(B, N, H, W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        call("kdf_bev_project_fwd", ptr(points), D, ptr(feats), dtype_code(feats), B, N, C, *geom, H, W, reduce,
             ptr(grid), ptr(count), ptr(cell), ptr(ties), ptr(order), ptr(offsets), ptr(ws), ws_bytes, stream_ptr(dev))
        ctx.reduce, ctx.dims = reduce, (B, N, C, H, W)
        if reduce == _n.REDUCE_MAX:
            ctx.save_for_backward(feats, grid, ties, cell, order, offsets)
        else:
            ctx.save_for_backward(count, cell, order, offsets)
        ctx.mark_non_differentiable(count, cell)
        return grid.permute(0, 3, 1, 2), count, cell      # lidar_encoder.py:99: [B,C,H,W] view of NHWC memory

    @staticmethod
    def backward(ctx, grad_grid, _gc, _gi):
        B, N, C, H, W = ctx.dims
        gg = grad_grid.permute(0, 2, 3, 1).contiguous()
        if ctx.reduce == _n.REDUCE_MAX:
            feats, grid, ties, cell, order, offsets = ctx.saved_tensors
            count = None
        else:
            count, cell, order, offsets = ctx.saved_tensors
            feats = grid = ties = None
        if gg.dtype != (feats.dtype if feats is not None else gg.dtype):
            gg = gg.to(feats.dtype)
        out = torch.empty(B, N, C, dtype=gg.dtype, device=gg.device)
        call("kdf_bev_project_bwd", ptr(gg), ptr(feats), ptr(grid), ptr(ties), ptr(count), ptr(cell), ptr(order), ptr(offsets),
             dtype_code(gg), B, N, C, H, W, ctx.reduce, ptr(out), stream_ptr(gg.device))
        return None, out, None, None, None


def range_index(points: torch.Tensor, grid_size: Tuple[int, int], fov_deg: Tuple[float, float] = (3.0, -25.0)):
    """Range-image cell id per point (int32, -1 invalid) and per-cell occupancy (int32 [B,H*W]): the spherical
    counterpart of ``bev_index`` (``kdf_range_index``; convention in include/kdfusion_b200.h).  ``fov_deg`` = (up, down)
    vertical field of view in degrees; points f32 [B,N,>=3]."""
    import math
    B, N, D = _check_points(points)
    if D < 3:
        raise ValueError("the range view needs (x, y, z) points")
    points = points.contiguous()
    H, W = grid_size
    cell = torch.empty(B, N, dtype=torch.int32, device=points.device)
    count = torch.empty(B, H * W, dtype=torch.int32, device=points.device)
    call("kdf_range_index", ptr(points), B, N, D, math.radians(fov_deg[0]), math.radians(fov_deg[1]), H, W, ptr(cell), ptr(count),
         stream_ptr(points.device))
    return cell, count


class RangeProjectFn(torch.autograd.Function):
    """range image [B,C,H,W] (NHWC memory) = per-cell max/mean of point-major feats [B,N,C]; the backward is the BEV one
    (``kdf_bev_project_bwd`` only sees cells)."""

    @staticmethod
    def forward(ctx, points, feats, fov, grid_size, reduce):
        import math
        B, N, D = _check_points(points)
        require_cuda(points, feats)
        if D < 3 or feats.dim() != 3 or feats.shape[0] != B or feats.shape[1] != N:
            raise ValueError(f"range_project: points [B,N,>=3] and feats [B,N,C] expected, got {tuple(points.shape)}, {tuple(feats.shape)}")
        points, feats = points.contiguous(), feats.contiguous()
        C = feats.shape[2]
        H, W = grid_size
        dev = points.device
        i32 = dict(dtype=torch.int32, device=dev)
        grid = torch.empty(B, H, W, C, dtype=feats.dtype, device=dev)
        count, cell = torch.empty(B, H * W, **i32), torch.empty(B, N, **i32)
        lanes, rem = divmod(C * feats.element_size(), 16)
        need_ties = reduce == _n.REDUCE_MAX and feats.requires_grad and not (rem == 0 and lanes in (8, 16, 32))
        ties = torch.empty(B, H * W, C, **i32) if need_ties else None
        order, offsets = torch.empty(B, N, **i32), torch.empty(B, H * W + 1, **i32)
        nb = lib.kdf_bev_workspace_bytes(B, N, H, W)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        call("kdf_range_project_fwd", ptr(points), D, ptr(feats), dtype_code(feats), B, N, C, math.radians(fov[0]), math.radians(fov[1]),
             H, W, reduce, ptr(grid), ptr(count), ptr(cell), ptr(ties), ptr(order), ptr(offsets), ptr(ws), nb, stream_ptr(dev))
        ctx.reduce, ctx.dims = reduce, (B, N, C, H, W)
        if reduce == _n.REDUCE_MAX:
            ctx.save_for_backward(feats, grid, ties, cell, order, offsets)
        else:
            ctx.save_for_backward(count, cell, order, offsets)
        ctx.mark_non_differentiable(count, cell)
        return grid.permute(0, 3, 1, 2), count, cell

    backward = staticmethod(BevProjectFn.backward)


def range_project(points: torch.Tensor, feats: torch.Tensor, grid_size, fov_deg=(3.0, -25.0), reduce: str = "max"):
    """-> (range image [B,C,H,W] over NHWC memory, count int32 [B,H*W], cell int32 [B,N])."""
    code = {"max": _n.REDUCE_MAX, "amax": _n.REDUCE_MAX, "mean": _n.REDUCE_MEAN}[reduce]
    return RangeProjectFn.apply(points, feats, tuple(fov_deg), tuple(grid_size), code)


def bev_project(points: torch.Tensor, feats: torch.Tensor, geom, grid_size, reduce: str = "max"):
    """-> (grid [B,C,H,W] over NHWC memory, count int32 [B,H*W], cell int32 [B,N])."""
    code = {"max": _n.REDUCE_MAX, "amax": _n.REDUCE_MAX, "mean": _n.REDUCE_MEAN}[reduce]
    return BevProjectFn.apply(points, feats, tuple(geom), tuple(grid_size), code)


# ----------------------------------------------------------------------------- (2) fusion
_MODES = {"add": 0, "concat": 1, "weighted": 2}


class _FusedFusionFn(torch.autograd.Function):
    """BN-apply + ReLU of both projection branches fused with the fusion itself.

    Inputs are pre-BatchNorm rows [M,C]; (mean, invstd) are the statistics the
    BatchNorm uses (batch statistics when ``batch_stats`` else running), computed
    by the caller.  Backward chains through the batch statistics analytically."""

    @staticmethod
    def forward(ctx, cam_pre, lid_pre, cam_g, cam_b, lid_g, lid_b, csc, csh, cam_mean, cam_invstd,
                lsc, lsh, lid_mean, lid_invstd, batch_stats, mode, w1, b1, w2, b2):
        dev = require_cuda(cam_pre, lid_pre, cam_g, lid_g, w1)
        cam_pre, lid_pre = cam_pre.contiguous(), lid_pre.contiguous()
        if cam_pre.shape != lid_pre.shape or cam_pre.dtype != lid_pre.dtype or cam_pre.dim() != 2:
            raise ValueError("fusion expects two [M,C] row tensors of the same shape and dtype")
        M, C = cam_pre.shape
        f32 = torch.float32
        dt, st = dtype_code(cam_pre), stream_ptr(dev)
        attn = None
        if mode == 2:
            w1c = w1.reshape(C, 2 * C).to(f32).contiguous()
            w2c = w2.reshape(2, C).to(f32).contiguous()
            b1c, b2c = b1.to(f32).contiguous(), b2.to(f32).contiguous()
            out = torch.empty(M, C, dtype=cam_pre.dtype, device=dev)
            attn = torch.empty(M, 2, dtype=f32, device=dev)
            call("kdf_fusion_weighted_fwd", ptr(cam_pre), ptr(lid_pre), dt, M, C, ptr(csc), ptr(csh), ptr(lsc), ptr(lsh),
                                              ptr(w1c), ptr(b1c), ptr(w2c), ptr(b2c), ptr(out), ptr(attn), st)
            ctx.save_for_backward(cam_pre, lid_pre, csc, csh, lsc, lsh, cam_mean, cam_invstd, lid_mean, lid_invstd,
                                  w1c, b1c, w2c, b2c, attn)
        else:
            out = torch.empty(M, C * (2 if mode == 1 else 1), dtype=cam_pre.dtype, device=dev)
            call("kdf_fusion_affine_relu_pair_fwd", ptr(cam_pre), ptr(lid_pre), dt, M, C, ptr(csc), ptr(csh), ptr(lsc),
                                                      ptr(lsh), mode, ptr(out), st)
            ctx.save_for_backward(cam_pre, lid_pre, csc, csh, lsc, lsh, cam_mean, cam_invstd, lid_mean, lid_invstd)
        ctx.mode, ctx.batch_stats = mode, batch_stats
        ctx.set_materialize_grads(False)        # the attention map output carries no gradient: do not materialise zeros for it
        ctx.param_dtypes = (cam_g.dtype, w1.dtype if w1 is not None else None)
        ctx.w_shapes = (w1.shape, w2.shape) if mode == 2 else None
        if attn is not None:
            ctx.mark_non_differentiable(attn)
            return out, attn
        return out, None

    @staticmethod
    def backward(ctx, grad_out, _ga):
        if grad_out is None:
            return (None,) * 20
        saved = ctx.saved_tensors
        cam_pre, lid_pre, csc, csh, lsc, lsh, cam_mean, cam_invstd, lid_mean, lid_invstd = saved[:10]
        M, C = cam_pre.shape
        dev = cam_pre.device
        f32 = torch.float32
        grad_out = grad_out.contiguous().to(cam_pre.dtype)
        g_cam, g_lid = torch.empty_like(cam_pre), torch.empty_like(lid_pre)
        gaff = torch.empty(4, C, dtype=f32, device=dev)
        dt, st = dtype_code(cam_pre), stream_ptr(dev)
        gw1 = gb1 = gw2 = gb2 = None
        if ctx.mode == 2:
            w1c, b1c, w2c, b2c, attn = saved[10:]
            gw1 = torch.empty(C, 2 * C, dtype=f32, device=dev)
            gb1 = torch.empty(C, dtype=f32, device=dev)
            gw2 = torch.empty(2, C, dtype=f32, device=dev)
            gb2 = torch.empty(2, dtype=f32, device=dev)
            call("kdf_fusion_weighted_bwd", ptr(grad_out), ptr(cam_pre), ptr(lid_pre), dt, M, C, ptr(csc), ptr(csh),
                                              ptr(lsc), ptr(lsh), ptr(w1c), ptr(b1c), ptr(w2c), ptr(b2c), ptr(attn),
                                              ptr(g_cam), ptr(g_lid), ptr(gaff), ptr(gw1), ptr(gb1), ptr(gw2), ptr(gb2), st)
            gw1, gw2 = gw1.view(ctx.w_shapes[0]), gw2.view(ctx.w_shapes[1])
        else:
            call("kdf_fusion_affine_relu_pair_bwd", ptr(grad_out), ptr(cam_pre), ptr(lid_pre), dt, M, C, ptr(csc),
                                                      ptr(csh), ptr(lsc), ptr(lsh), ctx.mode, ptr(g_cam), ptr(g_lid),
                                                      ptr(gaff), st)
        grads_gb = []
        for (x, g, scale, mean, invstd, s1, s0) in ((cam_pre, g_cam, csc, cam_mean, cam_invstd, gaff[0], gaff[1]),
                                                    (lid_pre, g_lid, lsc, lid_mean, lid_invstd, gaff[2], gaff[3])):
            dgamma = invstd * (s1 - mean * s0)          # sum dy * xhat
            dbeta = s0
            if ctx.batch_stats:                         # chain through the batch mean / variance
                bc = (-(scale * invstd * dgamma) / M).to(f32).contiguous()
                ac = (-(scale * s0) / M - bc * mean).to(f32).contiguous()
                call("kdf_rows_axpb", ptr(g), ptr(x), dt, M, C, ptr(bc), ptr(ac), st)      # g += bc*x + ac, one pass
            grads_gb.append((dgamma, dbeta))
        pd = ctx.param_dtypes[0]
        return (g_cam, g_lid, grads_gb[0][0].to(pd), grads_gb[0][1].to(pd), grads_gb[1][0].to(pd), grads_gb[1][1].to(pd),
                None, None, None, None, None, None, None, None, None, None, gw1, gb1, gw2, gb2)


def _bn_prepare_from_sums(rows: torch.Tensor, bn, col_sums):
    """``_bn_prepare`` when the kernel that produced ``rows`` already reduced their column sums (f64 [2,C])."""
    if col_sums is None or not (bn.training or bn.running_mean is None):
        return _bn_prepare(rows, bn)
    from .point_mlp import bn_finalize
    track = bn.training and bn.track_running_stats and bn.running_mean is not None
    with torch.no_grad():
        mean, invstd, scale, shift = bn_finalize(col_sums, rows.shape[0], bn, None, track)
    return scale, shift, mean, invstd, True


def fused_fusion(cam_pre, lid_pre, cam_bn, lid_bn, mode: str, attention=None, cam_sums=None, lid_sums=None):
    """Fused fusion over pre-BN rows.  ``cam_bn`` / ``lid_bn`` are the BatchNorm2d
    modules of the two Conv1x1 blocks (their running statistics are updated here
    exactly like nn.BatchNorm2d does in training); ``attention`` is the
    ``nn.Sequential(conv, relu, conv, softmax)`` of WeightedFusion; ``cam_sums`` / ``lid_sums`` are the column sums
    of the rows when the projection kernel already produced them (no statistics pass then).
    Returns (rows [M,C] or [M,2C], attn [M,2] | None)."""
    cs = _bn_prepare_from_sums(cam_pre.contiguous(), cam_bn, cam_sums)
    ls = _bn_prepare_from_sums(lid_pre.contiguous(), lid_bn, lid_sums)
    if cs[4] != ls[4]:
        raise RuntimeError("the two projection BatchNorms must be in the same mode")
    if mode == "weighted":
        w1, b1, w2, b2 = attention[0].weight, attention[0].bias, attention[2].weight, attention[2].bias
    else:
        w1 = b1 = w2 = b2 = None
    return _FusedFusionFn.apply(cam_pre, lid_pre, cam_bn.weight, cam_bn.bias, lid_bn.weight, lid_bn.bias,
                                cs[0], cs[1], cs[2], cs[3], ls[0], ls[1], ls[2], ls[3], cs[4], _MODES[mode],
                                w1, b1, w2, b2)


# ----------------------------------------------------------------------------- FPN-lite merge
class _FpnMergeFn(torch.autograd.Function):
    """out = base + bilinear(lo_a) [+ bilinear(lo_b)] (align_corners=False), NHWC; backward = the incoming
    gradient for ``base`` and ONE adjoint resize shared by the low-resolution inputs (exact 2x)."""

    @staticmethod
    def forward(ctx, base, lo_a, lo_b):
        B, C, H, W = base.shape
        h, w = lo_a.shape[-2:]
        out = torch.empty(B, H, W, C, dtype=base.dtype, device=base.device)
        call("kdf_fpn_merge_fwd", ptr(_nhwc_rows(base)), ptr(_nhwc_rows(lo_a)), ptr(_nhwc_rows(lo_b)) if lo_b is not None else None,
             dtype_code(base), B, H, W, h, w, C, ptr(out), stream_ptr(base.device))
        ctx.dims = (B, C, h, w)
        ctx.two = lo_b is not None
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        B, C, h, w = ctx.dims
        gr = _nhwc_rows(g)
        if gr is None:
            g = g.contiguous(memory_format=torch.channels_last)
            gr = g.permute(0, 2, 3, 1)
        glo = torch.empty(B, h, w, C, dtype=g.dtype, device=g.device)
        call("kdf_fpn_up2_bwd", ptr(gr), dtype_code(g), B, h, w, C, ptr(glo), stream_ptr(g.device))
        glo = glo.permute(0, 3, 1, 2)
        return g, glo, (glo if ctx.two else None)


def fpn_merge(base: torch.Tensor, lows: Sequence[torch.Tensor]) -> Optional[torch.Tensor]:
    """Fused FPN-lite merge when the layout allows it (dense channels-last CUDA maps of one dtype, one or two
    low-resolution maps at exactly half the base resolution, C a multiple of 8); None otherwise (the caller
    then composes F.interpolate + adds)."""
    if not (1 <= len(lows) <= 2) or not base.is_cuda:
        return None
    B, C, H, W = base.shape
    for t in lows:
        if t.dtype != base.dtype or t.shape[0] != B or t.shape[1] != C or tuple(t.shape[-2:]) != (H // 2, W // 2):
            return None
    if H % 2 or W % 2 or C % 8 or base.dtype not in (torch.float32, torch.bfloat16):
        return None
    if any(_nhwc_rows(t) is None for t in (base, *lows)):
        return None
    if len(lows) == 2 and tuple(lows[0].shape) != tuple(lows[1].shape):
        return None
    return _FpnMergeFn.apply(base, lows[0], lows[1] if len(lows) == 2 else None)


# ----------------------------------------------------------------------------- (3) distillation loss
def _dense_like(s: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """teacher tensor laid out exactly like the student one (so one flat pass pairs them)."""
    if t.shape != s.shape:
        raise ValueError(f"mimic tap shapes differ: {tuple(s.shape)} vs {tuple(t.shape)}")
    if t.dtype != s.dtype:
        t = t.to(s.dtype)
    if t.stride() != s.stride():
        t = torch.empty_like(s).copy_(t)
    return t


def _dense(s: torch.Tensor) -> torch.Tensor:
    if s.is_contiguous() or s.is_contiguous(memory_format=torch.channels_last):
        return s
    if s.dim() == 4 and s.permute(0, 2, 3, 1).is_contiguous():
        return s
    return s.contiguous()


def kd_label_count(labels: torch.Tensor, num_classes: int, ignore_index: int = -1) -> torch.Tensor:
    """Histogram of the valid labels into a fresh loss workspace (``kdf_kd_label_count``), on the current stream.  It
    depends on the labels only, so a training step takes it when the batch arrives (next to the teacher's forward on
    its side stream) and hands the workspace to ``kd_loss_fwd_bwd(..., counted_ws=...)``."""
    dev = require_cuda(labels)
    if labels.dtype != torch.int64 or labels.dim() != 3:
        raise ValueError(f"labels must be int64 [B,H,W], got {labels.dtype} {tuple(labels.shape)}")
    labels = labels.contiguous()
    B = labels.shape[0]
    ws = torch.empty(lib.kdf_kd_loss_workspace_bytes(), dtype=torch.uint8, device=dev)
    call("kdf_kd_label_count", ptr(labels), B, int(num_classes), labels.numel() // max(B, 1), int(ignore_index), ptr(ws),
         stream_ptr(dev))
    return ws


def kd_loss_fwd_bwd(student_logits, teacher_logits, labels, class_weights=None,
                    student_feats: Sequence[torch.Tensor] = (), teacher_feats: Sequence[torch.Tensor] = (),
                    T: float = 4.0, alpha: float = 0.5, beta: float = 1.0, ignore_index: int = -1,
                    grad_scale: float = 1.0, counted_ws: Optional[torch.Tensor] = None):
    """One pass: loss terms + gradients.  -> (scalars f32[8] = loss, ce, kl, mse, wsum, mse0, mse1, n_valid;
    d_logits like student_logits; [d_feat like each student feat]).  No host sync.  ``counted_ws``: a workspace that
    ``kd_label_count`` already filled for THESE labels (same class count and ignore index)."""
    dev = require_cuda(student_logits, teacher_logits, labels, class_weights, *student_feats, *teacher_feats)
    if student_logits.dim() != 4:
        raise ValueError("logits must be [B,K,H,W]")
    if len(student_feats) != len(teacher_feats) or len(student_feats) > 2:
        raise ValueError("0, 1 or 2 (student, teacher) mimic taps are supported")
    zs = student_logits.contiguous()
    zt = None if teacher_logits is None else teacher_logits.to(zs.dtype).contiguous()
    if zt is not None and zt.shape != zs.shape:
        raise ValueError("teacher and student logits differ in shape")
    B, K, H, W = zs.shape
    if labels.dtype != torch.int64 or tuple(labels.shape) != (B, H, W):
        raise ValueError(f"labels must be int64 [B,H,W]={B, H, W}, got {labels.dtype} {tuple(labels.shape)}")
    labels = labels.contiguous()
    cw = None if class_weights is None else class_weights.to(torch.float32).contiguous()
    if cw is not None and cw.numel() != K:
        raise ValueError(f"class_weights has {cw.numel()} entries, logits have {K} classes")
    s_list = [_dense(s) for s in student_feats]
    t_list = [_dense_like(s, t) for s, t in zip(s_list, teacher_feats)]
    if len({s.dtype for s in s_list}) > 1:
        raise TypeError("mimic taps must share one dtype")
    d_list = [torch.empty_like(s) for s in s_list]
    fd = dtype_code(s_list[0]) if s_list else _n.KDF_F32
    taps = []
    for i in range(2):
        if i < len(s_list):
            taps += [ptr(s_list[i]), ptr(t_list[i]), ptr(d_list[i]), s_list[i].numel()]
        else:
            taps += [None, None, None, 0]
    d_logits = torch.empty_like(zs)
    scalars = torch.empty(8, dtype=torch.float32, device=dev)
    ws = counted_ws if counted_ws is not None else torch.empty(lib.kdf_kd_loss_workspace_bytes(), dtype=torch.uint8, device=dev)
    call("kdf_kd_loss_fwd_bwd_counted" if counted_ws is not None else "kdf_kd_loss_fwd_bwd",
         ptr(zs), ptr(zt), ptr(labels), ptr(cw), B, K, H * W, dtype_code(zs),
                                  float(T), float(alpha), float(beta), int(ignore_index), *taps, fd, float(grad_scale),
                                  ptr(d_logits), ptr(scalars), ptr(ws), stream_ptr(dev))
    return scalars, d_logits, d_list


class KDLossFn(torch.autograd.Function):
    """Autograd face of the fused loss: returns the scalar loss; gradients were
    produced in the forward pass and are only scaled by the incoming grad."""

    @staticmethod
    def forward(ctx, student_logits, teacher_logits, labels, class_weights, T, alpha, beta, ignore_index, *feats):
        n = len(feats) // 2
        s_feats, t_feats = feats[:n], feats[n:]
        scalars, d_logits, d_feats = kd_loss_fwd_bwd(student_logits, teacher_logits, labels, class_weights,
                                                     s_feats, t_feats, T, alpha, beta, ignore_index)
        ctx.save_for_backward(d_logits, *d_feats)
        ctx.n = n
        ctx.mark_non_differentiable(scalars)
        return scalars[0].clone(), scalars

    @staticmethod
    def backward(ctx, g_loss, _gs):
        d_logits, *d_feats = ctx.saved_tensors
        return (d_logits * g_loss.to(d_logits.dtype), None, None, None, None, None, None, None,
                *[d * g_loss.to(d.dtype) for d in d_feats], *([None] * ctx.n))


# ----------------------------------------------------------------------------- step helpers
def confusion_matrix_(conf: torch.Tensor, logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -1):
    """conf[t,p] += 1 (int64 [K,K], on device) -- SegmentationMetrics.update without the host loop."""
    dev = require_cuda(conf, logits, labels)
    B, K = logits.shape[0], logits.shape[1]
    if conf.dtype != torch.int64 or tuple(conf.shape) != (K, K) or not conf.is_contiguous():
        raise ValueError("conf must be a contiguous int64 [K,K] tensor")
    logits, labels = logits.contiguous(), labels.contiguous()
    if labels.dtype != torch.int64:
        labels = labels.long()
    HW = logits.numel() // max(B * K, 1)
    call("kdf_confusion_matrix", ptr(logits), ptr(labels), B, K, HW, dtype_code(logits), int(ignore_index),
                                   ptr(conf), stream_ptr(dev))
    return conf


def adamw_flat_(param, grad, exp_avg, exp_avg_sq, hyper, beta1, beta2, eps, weight_decay, grad_scale=1.0, shadow=None):
    """In-place AdamW step over flat fp32 buffers; hyper = device tensor [lr, step]; ``shadow`` (bf16, same length)
    receives the updated parameters rounded to bf16."""
    dev = require_cuda(param, grad, exp_avg, exp_avg_sq, hyper)
    for t in (param, grad, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != param.numel():
            raise ValueError("adamw_flat_ needs contiguous float32 buffers of one size")
    if shadow is not None and (shadow.dtype != torch.bfloat16 or shadow.numel() != param.numel() or not shadow.is_contiguous()):
        raise ValueError("adamw_flat_: the shadow must be a contiguous bf16 buffer of the parameters' length")
    call("kdf_adamw_flat", ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), ptr(hyper),
                             float(beta1), float(beta2), float(eps), float(weight_decay), float(grad_scale),
                             ptr(shadow), stream_ptr(dev))


# ----------------------------------------------------------------------------- fused point-MLP layers (tcgen05)
def mlp_layer_fwd(mode: int, x: torch.Tensor, pro_a: torch.Tensor, pro_b: torch.Tensor, weight_bf16: torch.Tensor):
    """One tensor-core MLP layer: z [M,128] bf16 = prologue(x) @ W^T and the fp64 column sums of z and z^2.
    mode 0: x = raw points f32 [M,4] (first layer recomputed in the prologue, pro_a = q [64,4], pro_b = r [64]);
    mode 1: x = previous z bf16 [M,128] (prologue relu(z*pro_a + pro_b))."""
    dev = require_cuda(x, pro_a, pro_b, weight_bf16)
    x = x.contiguous()
    M = x.shape[0]
    Nout, Kin = weight_bf16.shape
    if weight_bf16.dtype != torch.bfloat16 or not weight_bf16.is_contiguous():
        raise TypeError("weight must be a contiguous bf16 [Nout, Kin] tensor")
    z = torch.empty(M, Nout, dtype=torch.bfloat16, device=dev)
    stats = torch.empty(2, Nout, dtype=torch.float64, device=dev)
    call("kdf_mlp_layer_fwd", mode, ptr(x), M, ptr(pro_a.float().contiguous()), ptr(pro_b.float().contiguous()),
         ptr(weight_bf16), Kin, Nout, ptr(z), ptr(stats), stream_ptr(dev))
    return z, stats


def mlp_layer_bwd(mode: int, dy: torch.Tensor, z: torch.Tensor, gs: torch.Tensor, ga: torch.Tensor, gb: torch.Tensor,
                  x: torch.Tensor, pro_a: torch.Tensor, pro_b: torch.Tensor, weight_bf16: torch.Tensor,
                  row_cell: Optional[torch.Tensor] = None):
    """The same layer backwards (one kernel): dz = gs*dy + ga + gb*z, dW = dz^T @ a_in, and
    mode 1: (dy_prev bf16 [M,128], sums f64 [2,128], dW f32 [128,128]);
    mode 0: (None, sums f64 [5,64] = sum dy1 * (1, x, y, z, i), dW f32 [128,64]).
    ``row_cell`` (i32 [M]): rows with a negative entry are treated as dy == 0 without trusting their contents."""
    dev = require_cuda(dy, z, gs, ga, gb, x, pro_a, pro_b, weight_bf16)
    M = dy.shape[0]
    Nout, Kin = weight_bf16.shape
    if dy.dtype != torch.bfloat16 or z.dtype != torch.bfloat16 or weight_bf16.dtype != torch.bfloat16:
        raise TypeError("dy, z and the weight must be bf16")
    if tuple(dy.shape) != (M, Nout) or tuple(z.shape) != (M, Nout) or x.shape[0] != M:
        raise ValueError("mlp_layer_bwd: shape mismatch")
    dy, z, x = dy.contiguous(), z.contiguous(), x.contiguous()
    dy_prev = torch.empty(M, Kin, dtype=torch.bfloat16, device=dev) if mode == 1 else None
    sums = torch.empty((2, Kin) if mode == 1 else (5, Kin), dtype=torch.float64, device=dev)
    dW = torch.empty(Nout, Kin, dtype=torch.float32, device=dev)
    f = lambda t: t.float().contiguous()
    call("kdf_mlp_layer_bwd", mode, ptr(dy), ptr(z), ptr(f(gs)), ptr(f(ga)), ptr(f(gb)), ptr(x), M, ptr(f(pro_a)), ptr(f(pro_b)),
         ptr(weight_bf16.contiguous()), Kin, ptr(dy_prev), ptr(sums), ptr(dW), ptr(row_cell), stream_ptr(dev))
    return dy_prev, sums, dW


# ----------------------------------------------------------------------------- pointwise (1x1) convolution layers (tcgen05)
def _pw_pack_factor(K: int, N: int) -> int:
    """Rows are handed to the tensor cores in 64-channel panels and the accumulator wants N % 32 == 0.  Narrow
    layers (K = 32: the first two blocks of the camera encoder) are run on PAIRS of pixel rows: x [M,32] is the same
    memory as [M/2, 64], and a block-diagonal weight [[W,0],[0,W]] makes the GEMM produce the pair's two output rows
    side by side -- [M/2, 2N] is the same memory as [M, N].  (The zero blocks cost flops the tensor cores do not
    notice; bytes moved are unchanged.)"""
    f = 1
    while (K * f) % 64 or (N * f) % 32:
        f *= 2
        if f > 8:
            raise ValueError(f"pw_conv: cannot pack K={K}, N={N} into 64-channel panels")
    return f


def pw_conv_supported(K: int, N: int, M: int) -> bool:
    try:
        f = _pw_pack_factor(K, N)
    except ValueError:
        return False
    return M % f == 0 and K * f <= 1024


def pw_conv_weight(weight: torch.Tensor, pack: int = 1) -> torch.Tensor:
    """bf16 [N*pack, K*pack] operand of ``pw_conv_fwd`` from a conv weight [N, K(,1,1)] (block-diagonal when packed)."""
    w = weight.detach().reshape(weight.shape[0], -1).to(torch.bfloat16)
    if pack > 1:
        w = torch.block_diag(*([w] * pack))
    return w.contiguous()


def pw_conv_fwd(x: torch.Tensor, weight_bf16: torch.Tensor, pack: int = 1, pro=None, epi=None,
                residual: Optional[torch.Tensor] = None, want_stats: bool = False):
    """One fused 1x1-convolution layer over pixel rows (``kdf_pw_conv_fwd``).
    x bf16 [M, K]; ``weight_bf16`` from ``pw_conv_weight`` (same ``pack``); ``pro`` = (scale f32[K], shift f32[K], act code)
    of the BatchNorm + activation in FRONT of the convolution, applied on the fly; ``epi`` = (scale f32[N], shift f32[N],
    act code) of a folded (running-statistics) BatchNorm + activation BEHIND it, with optional ``residual`` [M, N];
    ``want_stats``: also return the fp64 column sums [2, N] of the stored rows (the batch statistics of the
    BatchNorm behind the convolution).  -> out bf16 [M, N] (or (out, stats))."""
    dev = require_cuda(x, weight_bf16, residual)
    if x.dtype != torch.bfloat16 or weight_bf16.dtype != torch.bfloat16:
        raise TypeError("pw_conv_fwd runs on bf16 rows and weights")
    M, K = x.shape
    Np, Kp = weight_bf16.shape
    if Kp != K * pack or Np % pack or M % pack:
        raise ValueError(f"pw_conv_fwd: weight {tuple(weight_bf16.shape)} does not match rows {tuple(x.shape)} at pack {pack}")
    N = Np // pack
    x = x.contiguous()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    stats = torch.empty(2, Np, dtype=torch.float64, device=dev) if want_stats else None
    tile = (lambda v: v.float().repeat(pack).contiguous()) if pack > 1 else (lambda v: v.float().contiguous())
    ps = psh = es = esh = None
    pact = eact = 0
    if pro is not None:
        ps, psh, pact = tile(pro[0]), tile(pro[1]), int(pro[2])
    if epi is not None:
        es, esh, eact = tile(epi[0]), tile(epi[1]), int(epi[2])
    if residual is not None:
        residual = residual.contiguous()
        if residual.dtype != torch.bfloat16 or tuple(residual.shape) != (M, N):
            raise ValueError("pw_conv_fwd: residual must be bf16 [M, N]")
    call("kdf_pw_conv_fwd", ptr(x), M // pack, Kp, Np, ptr(weight_bf16), ptr(ps), ptr(psh), pact, ptr(es), ptr(esh), eact,
         ptr(residual), ptr(out), ptr(stats), stream_ptr(dev))
    if want_stats:
        if pack > 1:
            stats = stats.view(2, pack, N).sum(1)
        return out, stats
    return out
