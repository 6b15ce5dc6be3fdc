"""FPN-lite, the three camera-LiDAR fusion ablations, the segmentation heads and
the complete model (drop-in for the reference's ``src/models/fusion_module.py``:
same classes, constructor arguments, attribute names and state_dict keys).

The fusion blocks do not run the reference's chain of eager ops
(fusion_module.py:242-256: 2x [conv1x1, BN, ReLU], cat, conv, ReLU, conv, softmax,
2x mul, add).  The two 1x1 projections are row GEMMs over pixel-major features and
everything after them -- BatchNorm apply, ReLU, concat, attention MLP, softmax,
blend -- is one fused CUDA kernel per direction (``ops.fused_fusion``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


# ----------------------------------------------------------------------------- building blocks
class Conv1x1(nn.Module):
    """1x1 conv (no bias by default) + BatchNorm + ReLU under ``.conv`` (fusion_module.py:8-17)."""

    def __init__(self, in_ch, out_ch, bias=False):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, kernel_size=1, bias=bias),
                                  nn.BatchNorm2d(out_ch), nn.ReLU())

    def forward(self, x):
        return ops.run_fused(self.conv, x)


class DWSeparableConv(nn.Module):
    """depthwise 3x3 + BN + ReLU, pointwise 1x1 + BN + ReLU under ``.net`` (fusion_module.py:20-34)."""

    def __init__(self, in_ch, out_ch, stride=1):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv2d(in_ch, in_ch, kernel_size=3, stride=stride, padding=1, groups=in_ch, bias=False),
            nn.BatchNorm2d(in_ch), nn.ReLU(),
            nn.Conv2d(in_ch, out_ch, kernel_size=1, bias=False),
            nn.BatchNorm2d(out_ch), nn.ReLU())

    def forward(self, x):
        return ops.run_fused(self.net, x)


class CameraFPNLite(nn.Module):
    """Lateral 1x1 per stage -> bilinear resize to the largest map -> sum -> DW-separable
    smoothing (fusion_module.py:37-64)."""

    def __init__(self, in_channels_by_stage: Dict[str, int], target_channels: int = 128,
                 stages_to_use: Optional[List[str]] = None, target_size: Optional[Tuple[int, int]] = None):
        super().__init__()
        self.stages_to_use = stages_to_use or list(in_channels_by_stage.keys())
        self.laterals = nn.ModuleDict({s: Conv1x1(in_channels_by_stage[s], target_channels)
                                       for s in self.stages_to_use})
        self.post = DWSeparableConv(target_channels, target_channels)
        self.target_size = target_size

    def forward(self, feats: Dict[str, torch.Tensor]) -> torch.Tensor:
        if self.target_size is not None:
            size = tuple(self.target_size)
        else:
            size = tuple(max((feats[s].shape[-2:] for s in self.stages_to_use), key=lambda hw: hw[0] * hw[1]))
        lats = [self.laterals[s](feats[s]) for s in self.stages_to_use]
        # one streaming pass when the pyramid is "one full-resolution map + one or two maps at half resolution,
        # in that order" (the reference configuration: stage3 @64x64, stage4/5 @32x32)
        if len(lats) >= 2 and tuple(lats[0].shape[-2:]) == size:
            merged = ops.fpn_merge(lats[0], lats[1:])
            if merged is not None:
                return self.post(merged)
        total = None
        for lat in lats:
            if tuple(lat.shape[-2:]) != size:
                lat = F.interpolate(lat, size=size, mode="bilinear", align_corners=False)
            total = lat if total is None else total + lat
        return self.post(total)


# ----------------------------------------------------------------------------- fusion
def _rows(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> pixel-major rows [B*H*W, C]; free when x is channels-last (which both
    branches produce), one copy otherwise."""
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B * H * W, C)


def _maps(rows: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """rows [B*H*W, C] -> [B,C,H,W] view over the same NHWC memory."""
    return rows.view(B, H, W, rows.shape[-1]).permute(0, 3, 1, 2)


class _PairFusion(nn.Module):
    """Shared machinery: project both branches with the 1x1 convs of two Conv1x1
    blocks as row GEMMs, then hand the pre-BatchNorm rows to the fused kernel."""

    _mode = "add"
    _cam_attr, _lid_attr = "cam_proj", "lidar_proj"

    def _project(self, cam_feat, lidar_feat):
        if not (cam_feat.is_cuda and lidar_feat.is_cuda):
            raise RuntimeError("fusion runs on CUDA tensors only (no CPU fallback)")
        if cam_feat.shape[-2:] != lidar_feat.shape[-2:]:
            lidar_feat = F.interpolate(lidar_feat, size=cam_feat.shape[-2:], mode="bilinear", align_corners=False)
        cam_blk, lid_blk = getattr(self, self._cam_attr).conv, getattr(self, self._lid_attr).conv
        # bf16 training: the fused 1x1 layer kernel, whose epilogue also reduces the batch statistics the fusion kernel needs
        cam_pre, cam_sums = ops.pw_project_rows(cam_blk[0], cam_blk[1], _rows(cam_feat))
        lid_pre, lid_sums = ops.pw_project_rows(lid_blk[0], lid_blk[1], _rows(lidar_feat))
        if cam_pre.dtype != lid_pre.dtype:
            lid_pre, lid_sums = lid_pre.to(cam_pre.dtype), None
        return cam_pre, lid_pre, cam_blk[1], lid_blk[1], cam_sums, lid_sums

    def fuse_rows(self, cam_feat, lidar_feat):
        cam_pre, lid_pre, cam_bn, lid_bn, cam_sums, lid_sums = self._project(cam_feat, lidar_feat)
        rows, attn = ops.fused_fusion(cam_pre, lid_pre, cam_bn, lid_bn, self._mode, getattr(self, "attention", None),
                                      cam_sums, lid_sums)
        B, _, H, W = cam_feat.shape
        return _maps(rows, B, H, W), attn


class ConcatenationFusion(_PairFusion):
    """proj both -> concat -> DW3x3+BN+ReLU -> PW1x1+BN+ReLU (fusion_module.py:70-91)."""

    _mode = "concat"
    _cam_attr, _lid_attr = "camera_proj", "lidar_proj"

    def __init__(self, camera_channels=128, lidar_channels=128, out_channels=256):
        super().__init__()
        self.camera_proj = Conv1x1(camera_channels, camera_channels)
        self.lidar_proj = Conv1x1(lidar_channels, lidar_channels)
        cat = camera_channels + lidar_channels
        self.fuse = nn.Sequential(
            nn.Conv2d(cat, cat, kernel_size=3, padding=1, groups=cat, bias=False),
            nn.BatchNorm2d(cat), nn.ReLU(),
            nn.Conv2d(cat, out_channels, kernel_size=1, bias=False),
            nn.BatchNorm2d(out_channels), nn.ReLU())

    def forward_with_pre(self, cam_feat, lidar_feat):
        if self.camera_proj.conv[0].out_channels != self.lidar_proj.conv[0].out_channels:
            # unequal branch widths (the reference accepts any pair, fusion_module.py:74-76): the pair kernel wants one
            # row width, so each projection block runs on its own and the rows are concatenated
            if not (cam_feat.is_cuda and lidar_feat.is_cuda):
                raise RuntimeError("fusion runs on CUDA tensors only (no CPU fallback)")
            if cam_feat.shape[-2:] != lidar_feat.shape[-2:]:
                lidar_feat = F.interpolate(lidar_feat, size=cam_feat.shape[-2:], mode="bilinear", align_corners=False)
            cam_p, lid_p = self.camera_proj(cam_feat), self.lidar_proj(lidar_feat)
            pre = torch.cat([cam_p, lid_p.to(cam_p.dtype)], dim=1).contiguous(memory_format=torch.channels_last)
            return pre, ops.run_fused(self.fuse, pre)
        pre, _ = self.fuse_rows(cam_feat, lidar_feat)
        return pre, ops.run_fused(self.fuse, pre)

    def forward(self, cam_feat, lidar_feat):
        return self.forward_with_pre(cam_feat, lidar_feat)[1]


class MinimalFusion(_PairFusion):
    """proj both -> add (fusion_module.py:94-104)."""

    _mode = "add"

    def __init__(self, cam_ch=128, lidar_ch=128, out_ch=128):
        super().__init__()
        self.cam_proj = Conv1x1(cam_ch, out_ch)
        self.lidar_proj = Conv1x1(lidar_ch, out_ch)

    def forward_with_pre(self, cam_feat, lidar_feat):
        pre, _ = self.fuse_rows(cam_feat, lidar_feat)
        return pre, pre

    def forward(self, cam_feat, lidar_feat):
        return self.fuse_rows(cam_feat, lidar_feat)[0]


class WeightedFusion(_PairFusion):
    """proj both -> per-pixel 2-way softmax attention -> convex blend (fusion_module.py:107-136)."""

    _mode = "weighted"

    def __init__(self, cam_ch=128, lidar_ch=128, out_ch=128):
        super().__init__()
        self.cam_proj = Conv1x1(cam_ch, out_ch)
        self.lidar_proj = Conv1x1(lidar_ch, out_ch)
        self.attention = nn.Sequential(nn.Conv2d(out_ch * 2, out_ch, kernel_size=1), nn.ReLU(),
                                       nn.Conv2d(out_ch, 2, kernel_size=1), nn.Softmax(dim=1))

    def forward_with_pre(self, cam_feat, lidar_feat):
        pre, attn = self.fuse_rows(cam_feat, lidar_feat)
        self.last_attention = attn
        return pre, pre

    def forward(self, cam_feat, lidar_feat):
        return self.forward_with_pre(cam_feat, lidar_feat)[0]


# ----------------------------------------------------------------------------- heads
class LightweightSegmentationHead(nn.Module):
    """x4 upsampling head: two stride-2 transposed convs + 3x3 classifier (fusion_module.py:142-159)."""

    def __init__(self, in_channels=256, num_classes=2):
        super().__init__()
        self.up1 = nn.Sequential(nn.ConvTranspose2d(in_channels, 64, kernel_size=4, stride=2, padding=1, bias=False),
                                 nn.BatchNorm2d(64), nn.ReLU())
        self.up2 = nn.Sequential(nn.ConvTranspose2d(64, 16, kernel_size=4, stride=2, padding=1, bias=False),
                                 nn.BatchNorm2d(16), nn.ReLU())
        self.cls = nn.Conv2d(16, num_classes, kernel_size=3, padding=1)

    def forward(self, x):
        return self.cls(ops.run_fused(self.up2, ops.run_fused(self.up1, x)))


class SameResolutionSegmentationHead(nn.Module):
    """BEV-resolution head: DWSep(in,64) -> DWSep(64,32) -> 1x1 classifier (fusion_module.py:162-173)."""

    def __init__(self, in_channels=256, num_classes=2):
        super().__init__()
        self.block = nn.Sequential(DWSeparableConv(in_channels, 64), DWSeparableConv(64, 32))
        self.cls = nn.Conv2d(32, num_classes, kernel_size=1)

    def forward(self, x):
        h = self.block(x)
        y = ops.cls_conv(self.cls, h)              # 32 -> K classes: one kernel, planar logits (None: not the bf16 path)
        return self.cls(h) if y is None else y


# ----------------------------------------------------------------------------- complete model
_FUSIONS = {"concat": ConcatenationFusion, "minimal": MinimalFusion, "weighted": WeightedFusion}
_HEADS = {"x4": LightweightSegmentationHead, "same": SameResolutionSegmentationHead}


class CompleteSegmentationModel(nn.Module):
    """camera encoder (+FPN-lite) | LiDAR encoder -> fusion -> head (fusion_module.py:179-286)."""

    def __init__(self, camera_encoder: nn.Module, lidar_encoder: nn.Module, num_classes: int = 2,
                 fusion_type: str = "concat", fusion_out_channels: int = 256,
                 camera_fpn_stages: Optional[List[str]] = None, camera_fpn_channels: int = 128,
                 output_mode: str = "same"):
        super().__init__()
        self.camera_encoder = camera_encoder
        self.lidar_encoder = lidar_encoder
        self.fusion_type = fusion_type
        self.output_mode = output_mode

        self.use_multiscale = getattr(camera_encoder, "return_multiscale", False)
        self.camera_fpn = None
        if self.use_multiscale:
            self.camera_fpn = CameraFPNLite(camera_encoder.get_feature_info(), target_channels=camera_fpn_channels,
                                            stages_to_use=camera_fpn_stages)
            cam_ch = camera_fpn_channels
        else:
            cam_ch = getattr(camera_encoder, "out_channels", 128)
        lid_ch = getattr(getattr(lidar_encoder, "encoder", lidar_encoder), "feature_dim", 128)

        if fusion_type not in _FUSIONS:
            raise ValueError(f"Unknown fusion_type: {fusion_type}")
        if fusion_type == "concat":
            self.fusion = ConcatenationFusion(cam_ch, lid_ch, fusion_out_channels)
            head_in = fusion_out_channels
        else:
            self.fusion = _FUSIONS[fusion_type](cam_ch=cam_ch, lidar_ch=lid_ch, out_ch=cam_ch)
            head_in = cam_ch
        if output_mode not in _HEADS:
            raise ValueError(f"Unknown output_mode: {output_mode}")
        self.head = _HEADS[output_mode](in_channels=head_in, num_classes=num_classes)

    def forward(self, images: torch.Tensor, points: torch.Tensor, return_intermediates: bool = False):
        cam = self.camera_encoder(images)
        cam_feat = self.camera_fpn(cam) if isinstance(cam, dict) else cam
        lidar_feat = self.lidar_encoder(points)
        if cam_feat.shape[-2:] != lidar_feat.shape[-2:]:
            lidar_feat = F.interpolate(lidar_feat, size=cam_feat.shape[-2:], mode="bilinear", align_corners=False)
        if lidar_feat.dtype != cam_feat.dtype:
            lidar_feat = lidar_feat.to(cam_feat.dtype)
        pre_fusion, fused = self.fusion.forward_with_pre(cam_feat, lidar_feat)
        logits = self.head(fused)
        if return_intermediates:
            return logits, {"camera_feat": cam_feat, "lidar_feat": lidar_feat, "pre_fusion": pre_fusion,
                            "post_fusion": fused, "logits": logits}
        return logits

    def get_architecture_summary(self):
        def n(m):
            return sum(p.numel() for p in m.parameters())
        fusion = n(self.fusion) + (n(self.camera_fpn) if self.camera_fpn is not None else 0)
        return {"camera_params": f"{n(self.camera_encoder):,}", "lidar_params": f"{n(self.lidar_encoder):,}",
                "fusion_params": f"{fusion:,}", "head_params": f"{n(self.head):,}", "total_params": f"{n(self):,}",
                "fusion_type": self.fusion_type, "output_mode": self.output_mode,
                "use_multiscale": self.use_multiscale}
