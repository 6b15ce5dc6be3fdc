"""Camera branch: TwinLite-style MobileNetV2 encoder (drop-in for the reference's
``src/models/camera_encoder.py``; same class names, constructor arguments,
attributes and state_dict keys -- ``stem.{0,1}``, ``stageK.conv.{i}``).

The convolutions are dense contractions / depthwise stencils and stay on the
library tensor-core path (cuDNN / cuBLAS); what this implementation changes is the
execution shape: the whole branch runs channels-last (NHWC) so that the 1x1
convolutions are plain GEMMs over pixel rows and the feature map handed to the
fusion kernel is already pixel-major, and it is autocast-friendly (bf16
activations, fp32 parameters and BatchNorm statistics).
"""
from __future__ import annotations

from typing import Dict, List, Union

import torch
import torch.nn as nn

from ..ops import run_fused, stem_conv


def _conv_bn(cin: int, cout: int, kernel: int, stride: int = 1, groups: int = 1, act: bool = True) -> List[nn.Module]:
    """conv (no bias) -> BatchNorm2d [-> ReLU6], as a flat list so callers control Sequential indices."""
    layers: List[nn.Module] = [
        nn.Conv2d(cin, cout, kernel_size=kernel, stride=stride, padding=kernel // 2, groups=groups, bias=False),
        nn.BatchNorm2d(cout),
    ]
    if act:
        layers.append(nn.ReLU6())
    return layers


class InvertedResidual(nn.Module):
    """expand 1x1 (skipped when expansion_ratio == 1) -> depthwise 3x3 -> linear 1x1 projection,
    identity shortcut when the block keeps shape (reference camera_encoder.py:9-51)."""

    def __init__(self, in_channels, out_channels, stride=1, expansion_ratio=6):
        super().__init__()
        hidden = int(round(in_channels * expansion_ratio))
        self.use_residual = stride == 1 and in_channels == out_channels
        seq: List[nn.Module] = []
        if expansion_ratio != 1:
            seq += _conv_bn(in_channels, hidden, 1)
        seq += _conv_bn(hidden, hidden, 3, stride=stride, groups=hidden)
        seq += _conv_bn(hidden, out_channels, 1, act=False)
        self.conv = nn.Sequential(*seq)

    def forward(self, x):
        # conv layers on cuDNN; every BatchNorm(+ReLU6) group and the shortcut add on the fused row kernels
        return run_fused(self.conv, x, residual=x if self.use_residual else None)


class TwinLiteEncoder(nn.Module):
    """stem (3x3 s2) + five inverted-residual stages; multiscale dict on request
    (reference camera_encoder.py:56-123)."""

    # (name, cin multiple, cout multiple, stride, expansion)
    _STAGES = (("stage1", 1, 1, 1, 1), ("stage2", 1, 2, 2, 6), ("stage3", 2, 2, 1, 6),
               ("stage4", 2, 4, 2, 6), ("stage5", 4, 4, 1, 6))

    def __init__(self, in_channels=3, base_channels=32, return_multiscale=False):
        super().__init__()
        self.return_multiscale = return_multiscale
        self.stem = nn.Sequential(*_conv_bn(in_channels, base_channels, 3, stride=2))
        for name, ci, co, stride, expand in self._STAGES:
            setattr(self, name, InvertedResidual(base_channels * ci, base_channels * co, stride=stride,
                                                 expansion_ratio=expand))
        self.feature_channels = {"stage2": base_channels * 2, "stage3": base_channels * 2,
                                 "stage4": base_channels * 4, "stage5": base_channels * 4}
        self.out_channels = base_channels * 4

    def forward(self, x) -> Union[torch.Tensor, Dict[str, torch.Tensor]]:
        if not x.is_cuda:
            raise RuntimeError("TwinLiteEncoder runs on CUDA tensors only (no CPU fallback)")
        y = stem_conv(self.stem, x)                                  # fp32 NCHW image -> bf16 NHWC rows in one kernel
        if y is None:
            x = x.contiguous(memory_format=torch.channels_last)     # NHWC end to end
            y = run_fused(self.stem, x)
        x = y
        feats = {}
        for name, *_ in self._STAGES:
            x = getattr(self, name)(x)
            feats[name] = x
        if self.return_multiscale:
            return {k: feats[k] for k in ("stage2", "stage3", "stage4", "stage5")}
        return x

    def get_feature_info(self):
        return self.feature_channels

    def count_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)
