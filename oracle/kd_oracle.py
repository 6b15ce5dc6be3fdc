"""Frozen knowledge-distillation loss oracle (TEST INFRASTRUCTURE ONLY).

The reference contains NO distillation loss: the only
loss it has is ``nn.CrossEntropyLoss(ignore_index=-1, weight=w)``
(``src/training/trainer.py:55,88``).  The KL and feature-mimic terms asked for
by the north star are therefore PARITY UNPINNED by the reference; this file is
the frozen specification from SURVEY.md section 8c, written once as a plain
torch composition and never changed.  The CE term keeps the reference's
semantics exactly and is pinned against it in
``tests/test_oracle_vs_reference.py``.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

DEFAULT_T = 4.0
DEFAULT_ALPHA = 0.5
DEFAULT_BETA = 1.0
MIMIC_TAPS = ("lidar_feat", "camera_feat")     # fusion_module.py:260-262 intermediates


def ce_loss(logits: torch.Tensor, labels: torch.Tensor,
            class_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """trainer.py:55,88 -- weighted mean over pixels with label != -1."""
    return F.cross_entropy(logits, labels, weight=class_weights, ignore_index=-1)


def kd_loss(student_logits: torch.Tensor, teacher_logits: torch.Tensor, labels: torch.Tensor,
            class_weights: Optional[torch.Tensor],
            student_feats: Sequence[torch.Tensor] = (), teacher_feats: Sequence[torch.Tensor] = (),
            T: float = DEFAULT_T, alpha: float = DEFAULT_ALPHA, beta: float = DEFAULT_BETA
            ) -> Dict[str, torch.Tensor]:
    """loss = (1-alpha)*ce + alpha*kl + beta*mse   (SURVEY.md section 8c).

    kl  = T^2 * sum_pixels KL(softmax(z_t/T) || softmax(z_s/T)) / (B*H*W)
    mse = sum over taps of mean((s - t)^2)
    """
    B, K, H, W = student_logits.shape
    ce = ce_loss(student_logits, labels, class_weights)
    kl = (T * T) * F.kl_div(F.log_softmax(student_logits / T, dim=1),
                            F.softmax(teacher_logits / T, dim=1), reduction="sum") / (B * H * W)
    mse = student_logits.new_zeros(())
    for s, t in zip(student_feats, teacher_feats):
        mse = mse + F.mse_loss(s, t)
    loss = (1.0 - alpha) * ce + alpha * kl + beta * mse
    return {"loss": loss, "ce": ce, "kl": kl, "mse": mse}


def confusion_matrix(logits: torch.Tensor, labels: torch.Tensor, num_classes: int = 2,
                     ignore_index: int = -1) -> torch.Tensor:
    """``SegmentationMetrics.update`` (trainer.py:18-26) without the Python loop:
    conf[t, p] += 1 for every pixel with t != ignore and both in range."""
    pred = logits.argmax(dim=1).reshape(-1)
    tgt = labels.reshape(-1)
    ok = (tgt != ignore_index) & (tgt >= 0) & (tgt < num_classes) & (pred >= 0) & (pred < num_classes)
    idx = tgt[ok] * num_classes + pred[ok]
    return torch.bincount(idx, minlength=num_classes * num_classes).view(num_classes, num_classes)


def miou(conf: torch.Tensor):
    """``SegmentationMetrics.compute`` (trainer.py:28-37)."""
    conf = conf.double()
    ious = []
    for i in range(conf.shape[0]):
        tp = conf[i, i]
        denom = conf[:, i].sum() + conf[i, :].sum() - tp
        ious.append(float(tp / denom) if denom > 0 else 0.0)
    return {"class_iou": ious, "miou": float(sum(ious) / len(ious))}
