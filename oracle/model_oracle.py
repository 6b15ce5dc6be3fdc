"""Functional fp32 torch oracle of the reference network (TEST INFRASTRUCTURE ONLY).

The whole camera + LiDAR segmentation model of the reference, restated as
pure functions over a reference-format ``state_dict`` (the key names are the
reference's) so that it can be checked against the
reference's ``nn.Module`` graph weight-for-weight, and so that the product's
CUDA path can be checked against it on a box where ``/root/reference`` does
not exist.  Runs on CPU (or any device the tensors live on) in eager fp32.

Reference files restated:
  src/models/camera_encoder.py:9-115     inverted residual stages + stem
  src/models/lidar_encoder.py:25-99      point MLP + BEV amax scatter
  src/models/fusion_module.py:8-64       Conv1x1 / DWSeparableConv / FPN-lite
  src/models/fusion_module.py:70-136     concat / minimal / weighted fusion
  src/models/fusion_module.py:142-173    the two heads
  src/models/fusion_module.py:234-263    CompleteSegmentationModel.forward

Pinned against the reference in ``tests/test_oracle_vs_reference.py`` and via
``tests/golden/model_*.npz``.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

BN_EPS = 1e-5       # nn.BatchNorm default (the reference never overrides it)
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------- helpers
def _bn(x: Tensor, sd: StateDict, p: str, train: bool) -> Tensor:
    """nn.BatchNorm{1,2}d with the tensors stored under prefix ``p``.

    In train mode the running statistics in ``sd`` are updated in place exactly
    as the module would (momentum 0.1, unbiased variance)."""
    rm, rv = sd.get(p + ".running_mean"), sd.get(p + ".running_var")
    out = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"],
                       training=train, momentum=BN_MOMENTUM, eps=BN_EPS)
    if train and (p + ".num_batches_tracked") in sd:
        sd[p + ".num_batches_tracked"] += 1
    return out


def _conv2d(x, sd, p, stride=1, padding=0, groups=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride,
                    padding=padding, groups=groups)


def _pw_bn_relu(x, sd, p, train):
    """``Conv1x1`` block: 1x1 conv (no bias) + BN + ReLU  (fusion_module.py:8-17)."""
    return F.relu(_bn(_conv2d(x, sd, p + ".conv.0"), sd, p + ".conv.1", train))


def _dwsep(x, sd, p, train):
    """``DWSeparableConv``: DW3x3+BN+ReLU, PW1x1+BN+ReLU  (fusion_module.py:20-34)."""
    c = x.shape[1]
    x = F.relu(_bn(_conv2d(x, sd, p + ".net.0", padding=1, groups=c), sd, p + ".net.1", train))
    return F.relu(_bn(_conv2d(x, sd, p + ".net.3"), sd, p + ".net.4", train))


# --------------------------------------------------------------------------- camera
def _inverted_residual(x, sd, p, stride, expand, residual, train):
    """camera_encoder.py:9-51.  Sequential indices differ with/without expansion."""
    y, i = x, 0
    if expand:
        y = F.relu6(_bn(_conv2d(y, sd, f"{p}.conv.{i}"), sd, f"{p}.conv.{i+1}", train))
        i += 3
    c = y.shape[1]
    y = F.relu6(_bn(_conv2d(y, sd, f"{p}.conv.{i}", stride=stride, padding=1, groups=c),
                    sd, f"{p}.conv.{i+1}", train))
    i += 3
    y = _bn(_conv2d(y, sd, f"{p}.conv.{i}"), sd, f"{p}.conv.{i+1}", train)
    return x + y if residual else y


def camera_encoder(images: Tensor, sd: StateDict, prefix="camera_encoder", train=True):
    """``TwinLiteEncoder.forward`` with ``return_multiscale=True``
    (camera_encoder.py:95-112) -> dict stage2..stage5."""
    p = prefix
    x = F.relu6(_bn(_conv2d(images, sd, p + ".stem.0", stride=2, padding=1), sd, p + ".stem.1", train))
    x1 = _inverted_residual(x, sd, p + ".stage1", 1, False, True, train)
    x2 = _inverted_residual(x1, sd, p + ".stage2", 2, True, False, train)
    x3 = _inverted_residual(x2, sd, p + ".stage3", 1, True, True, train)
    x4 = _inverted_residual(x3, sd, p + ".stage4", 2, True, False, train)
    x5 = _inverted_residual(x4, sd, p + ".stage5", 1, True, True, train)
    return {"stage2": x2, "stage3": x3, "stage4": x4, "stage5": x5}


def camera_fpn(feats: Dict[str, Tensor], sd: StateDict, stages: Sequence[str],
               prefix="camera_fpn", train=True) -> Tensor:
    """``CameraFPNLite.forward`` (fusion_module.py:51-64)."""
    sizes = [feats[s].shape[-2:] for s in stages]
    H, W = max(sizes, key=lambda hw: hw[0] * hw[1])
    acc = None
    for s in stages:
        x = _pw_bn_relu(feats[s], sd, f"{prefix}.laterals.{s}", train)
        if tuple(x.shape[-2:]) != (H, W):
            x = F.interpolate(x, size=(H, W), mode="bilinear", align_corners=False)
        acc = x if acc is None else acc + x
    return _dwsep(acc, sd, prefix + ".post", train)


# --------------------------------------------------------------------------- lidar
def point_mlp(points: Tensor, sd: StateDict, prefix="lidar_encoder.encoder.point_mlp", train=True):
    """Conv1d(4,64)+BN+ReLU, Conv1d(64,128)+BN+ReLU, Conv1d(128,C)+BN+ReLU
    on [B,4,N] (lidar_encoder.py:25-35,66) -> [B,C,N]."""
    x = points.transpose(1, 2)
    for conv, bn in ((0, 1), (3, 4), (6, 7)):
        x = F.conv1d(x, sd[f"{prefix}.{conv}.weight"], sd[f"{prefix}.{conv}.bias"])
        x = F.relu(_bn(x, sd, f"{prefix}.{bn}", train))
    return x


def bev_flat_index(points: Tensor, sd: StateDict, grid_size: Tuple[int, int],
                   prefix="lidar_encoder.encoder"):
    """lidar_encoder.py:42-55,69-79 with the module's own buffers
    (``x_range``/``y_range`` int64, ``grid_tensor`` fp32)."""
    H, W = grid_size
    xr, yr, gt = sd[prefix + ".x_range"], sd[prefix + ".y_range"], sd[prefix + ".grid_tensor"]
    xn = (points[..., 0] - xr[0]) / (xr[1] - xr[0])
    yn = (points[..., 1] - yr[0]) / (yr[1] - yr[0])
    valid = (xn >= 0) & (xn <= 1) & (yn >= 0) & (yn <= 1)
    g = (torch.stack([xn, yn], dim=-1) * gt).long()
    col = g[..., 0].clamp(0, W - 1)
    row = g[..., 1].clamp(0, H - 1)
    B, N = points.shape[:2]
    b = torch.arange(B, device=points.device).view(B, 1).expand(B, N)
    flat = b * (H * W) + row * W + col
    return flat, valid


def lidar_encoder(points: Tensor, sd: StateDict, grid_size=(64, 64),
                  prefix="lidar_encoder.encoder", train=True) -> Tensor:
    """``SpatialLiDAREncoder.forward_vectorized`` (lidar_encoder.py:57-99).
    Returns the [B,C,H,W] view over NHWC memory, like the reference."""
    H, W = grid_size
    B, N, _ = points.shape
    flat, valid = bev_flat_index(points, sd, grid_size, prefix)
    feats = point_mlp(points, sd, prefix + ".point_mlp", train)       # [B,C,N]
    C = feats.shape[1]
    sel = feats.permute(0, 2, 1)[valid]
    out = torch.zeros(B * H * W, C, dtype=points.dtype, device=points.device)
    if sel.numel() > 0:
        out.scatter_reduce_(0, flat[valid].unsqueeze(1).expand(-1, C), sel,
                            reduce="amax", include_self=False)
    return out.view(B, H, W, C).permute(0, 3, 1, 2)


# --------------------------------------------------------------------------- fusion / head
def fusion(cam_feat: Tensor, lidar_feat: Tensor, sd: StateDict, fusion_type: str,
           prefix="fusion", train=True):
    """The inline fusion of ``CompleteSegmentationModel.forward``
    (fusion_module.py:242-256).  Returns (pre_fusion, fused, extras)."""
    extras = {}
    if fusion_type == "concat":
        cp = _pw_bn_relu(cam_feat, sd, prefix + ".camera_proj", train)
        lp = _pw_bn_relu(lidar_feat, sd, prefix + ".lidar_proj", train)
        pre = torch.cat([cp, lp], dim=1)
        c = pre.shape[1]
        y = F.relu(_bn(_conv2d(pre, sd, prefix + ".fuse.0", padding=1, groups=c), sd, prefix + ".fuse.1", train))
        fused = F.relu(_bn(_conv2d(y, sd, prefix + ".fuse.3"), sd, prefix + ".fuse.4", train))
    elif fusion_type in ("weighted", "minimal"):
        cp = _pw_bn_relu(cam_feat, sd, prefix + ".cam_proj", train)
        lp = _pw_bn_relu(lidar_feat, sd, prefix + ".lidar_proj", train)
        if fusion_type == "weighted":
            a = F.relu(_conv2d(torch.cat([cp, lp], dim=1), sd, prefix + ".attention.0"))
            w = torch.softmax(_conv2d(a, sd, prefix + ".attention.2"), dim=1)
            pre = cp * w[:, 0:1] + lp * w[:, 1:2]
            extras["attention"] = w
        else:
            pre = cp + lp
        fused = pre
    else:
        raise ValueError(f"Unknown fusion_type: {fusion_type}")
    extras["cam_proj"], extras["lidar_proj"] = cp, lp
    return pre, fused, extras


def head(x: Tensor, sd: StateDict, output_mode="same", prefix="head", train=True) -> Tensor:
    """fusion_module.py:142-173."""
    if output_mode == "same":
        x = _dwsep(x, sd, prefix + ".block.0", train)
        x = _dwsep(x, sd, prefix + ".block.1", train)
        return _conv2d(x, sd, prefix + ".cls")
    if output_mode == "x4":
        for up in ("up1", "up2"):
            x = F.conv_transpose2d(x, sd[f"{prefix}.{up}.0.weight"], None, stride=2, padding=1)
            x = F.relu(_bn(x, sd, f"{prefix}.{up}.1", train))
        return _conv2d(x, sd, prefix + ".cls", padding=1)
    raise ValueError(f"Unknown output_mode: {output_mode}")


# --------------------------------------------------------------------------- whole model
def model_forward(images: Tensor, points: Tensor, sd: StateDict, *, fusion_type="weighted",
                  grid_size=(64, 64), fpn_stages=("stage3", "stage4", "stage5"),
                  output_mode="same", train=True):
    """``CompleteSegmentationModel.forward(..., return_intermediates=True)``
    (fusion_module.py:234-263) -> (logits, intermediates)."""
    cam_raw = camera_encoder(images, sd, "camera_encoder", train)
    cam_feat = camera_fpn(cam_raw, sd, list(fpn_stages), "camera_fpn", train)
    lid_feat = lidar_encoder(points, sd, grid_size, "lidar_encoder.encoder", train)
    if cam_feat.shape[-2:] != lid_feat.shape[-2:]:
        lid_feat = F.interpolate(lid_feat, size=cam_feat.shape[-2:], mode="bilinear", align_corners=False)
    pre, fused, extras = fusion(cam_feat, lid_feat, sd, fusion_type, "fusion", train)
    logits = head(fused, sd, output_mode, "head", train)
    mids = {"camera_feat": cam_feat, "lidar_feat": lid_feat, "pre_fusion": pre,
            "post_fusion": fused, "logits": logits}
    mids.update({"_" + k: v for k, v in extras.items()})
    return logits, mids


def clone_state(sd: StateDict, requires_grad: bool = False) -> StateDict:
    """Detached copy; float tensors optionally become autograd leaves."""
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and "running_" not in k and not k.endswith("grid_tensor"):
            t.requires_grad_(True)
        out[k] = t
    return out


# --------------------------------------------------------------------------- shapes
def state_dict_spec(fusion_type="weighted", num_classes=2, fusion_out_channels: Optional[int] = None,
                    feature_dim=128, fpn_channels=128, fpn_stages=("stage3", "stage4", "stage5"),
                    output_mode="same", base=32, grid_size=(64, 64),
                    point_cloud_range=(-50, -50, -5, 50, 50, 3)):
    """Ordered {key: (shape, kind)} of the reference state_dict for one model
    configuration; kind in {"conv","bias","bn_w","bn_b","bn_rm","bn_rv","bn_n","buf"}.
    Derived from the constructors in the three reference model files."""
    spec: Dict[str, Tuple[Tuple[int, ...], str]] = {}

    def conv(k, shape, bias=False):
        spec[k + ".weight"] = (tuple(shape), "conv")
        if bias:
            spec[k + ".bias"] = ((shape[0],), "bias")

    def bn(k, c):
        spec[k + ".weight"] = ((c,), "bn_w"); spec[k + ".bias"] = ((c,), "bn_b")
        spec[k + ".running_mean"] = ((c,), "bn_rm"); spec[k + ".running_var"] = ((c,), "bn_rv")
        spec[k + ".num_batches_tracked"] = ((), "bn_n")

    # camera (camera_encoder.py:63-82)
    conv("camera_encoder.stem.0", (base, 3, 3, 3)); bn("camera_encoder.stem.1", base)

    def ir(p, cin, cout, expand):
        hid, i = cin * expand, 0
        if expand != 1:
            conv(f"{p}.conv.0", (hid, cin, 1, 1)); bn(f"{p}.conv.1", hid); i = 3
        conv(f"{p}.conv.{i}", (hid, 1, 3, 3)); bn(f"{p}.conv.{i+1}", hid)
        conv(f"{p}.conv.{i+3}", (cout, hid, 1, 1)); bn(f"{p}.conv.{i+4}", cout)

    ir("camera_encoder.stage1", base, base, 1)
    ir("camera_encoder.stage2", base, 2 * base, 6)
    ir("camera_encoder.stage3", 2 * base, 2 * base, 6)
    ir("camera_encoder.stage4", 2 * base, 4 * base, 6)
    ir("camera_encoder.stage5", 4 * base, 4 * base, 6)
    # lidar (lidar_encoder.py:25-40)
    p = "lidar_encoder.encoder"
    dims = [4, 64, 128, feature_dim]
    # a module's own buffers precede its children in state_dict order
    spec[p + ".x_range"] = ((2,), "buf"); spec[p + ".y_range"] = ((2,), "buf")
    spec[p + ".grid_tensor"] = ((2,), "buf")
    for li, (ci, co) in enumerate(zip(dims[:-1], dims[1:])):
        conv(f"{p}.point_mlp.{3*li}", (co, ci, 1), bias=True); bn(f"{p}.point_mlp.{3*li+1}", co)

    def c1x1(k, ci, co):
        conv(k + ".conv.0", (co, ci, 1, 1)); bn(k + ".conv.1", co)

    def dws(k, ci, co):
        conv(k + ".net.0", (ci, 1, 3, 3)); bn(k + ".net.1", ci)
        conv(k + ".net.3", (co, ci, 1, 1)); bn(k + ".net.4", co)

    # fusion (fusion_module.py:70-136, 213-222)
    stage_ch = {"stage2": 2 * base, "stage3": 2 * base, "stage4": 4 * base, "stage5": 4 * base}
    if fusion_type == "concat":
        oc = 256 if fusion_out_channels is None else fusion_out_channels
        c1x1("fusion.camera_proj", fpn_channels, fpn_channels)
        c1x1("fusion.lidar_proj", feature_dim, feature_dim)
        cat = fpn_channels + feature_dim
        conv("fusion.fuse.0", (cat, 1, 3, 3)); bn("fusion.fuse.1", cat)
        conv("fusion.fuse.3", (oc, cat, 1, 1)); bn("fusion.fuse.4", oc)
        head_in = oc
    else:
        c1x1("fusion.cam_proj", fpn_channels, fpn_channels)
        c1x1("fusion.lidar_proj", feature_dim, fpn_channels)
        if fusion_type == "weighted":
            conv("fusion.attention.0", (fpn_channels, 2 * fpn_channels, 1, 1), bias=True)
            conv("fusion.attention.2", (2, fpn_channels, 1, 1), bias=True)
        head_in = fpn_channels
    # registration order in CompleteSegmentationModel.__init__: camera_encoder,
    # lidar_encoder, camera_fpn, fusion, head  (fusion_module.py:191-232)
    fus = {k: v for k, v in spec.items() if k.startswith("fusion.")}
    for k in fus:
        del spec[k]
    for s in fpn_stages:
        c1x1(f"camera_fpn.laterals.{s}", stage_ch[s], fpn_channels)
    dws("camera_fpn.post", fpn_channels, fpn_channels)
    spec.update(fus)
    if output_mode == "same":
        dws("head.block.0", head_in, 64); dws("head.block.1", 64, 32)
        conv("head.cls", (num_classes, 32, 1, 1), bias=True)
    else:
        spec["head.up1.0.weight"] = ((head_in, 64, 4, 4), "conv"); bn("head.up1.1", 64)
        spec["head.up2.0.weight"] = ((64, 16, 4, 4), "conv"); bn("head.up2.1", 16)
        conv("head.cls", (num_classes, 16, 3, 3), bias=True)
    return spec
