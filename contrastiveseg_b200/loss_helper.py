"""Segmentation cross-entropy used inside ContrastCELoss (lib/loss/loss_helper.py:169-212, :215-280).

OUT OF THE HOT PATH: kept as plain PyTorch ops so the wrappers are drop-in.
RMI / Lovasz variants (contrast.use_rmi / use_lovasz) are not re-implemented: pass the reference's own
module through ``seg_criterion=`` if needed.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _ce_params(configer):
    weight, reduction, ignore_index = None, "mean", -1
    if configer is not None and configer.exists("loss", "params"):
        p = configer.get("loss", "params")
        if "ce_weight" in p:
            weight = torch.tensor(p["ce_weight"], dtype=torch.float32)
        if "ce_reduction" in p:
            reduction = {"elementwise_mean": "mean"}.get(p["ce_reduction"], p["ce_reduction"])
        if "ce_ignore_index" in p:
            ignore_index = p["ce_ignore_index"]
    return weight, reduction, ignore_index


def _scale_target(target: torch.Tensor, size) -> torch.Tensor:
    t = target.clone().unsqueeze(1).float()
    return F.interpolate(t, size=size, mode="nearest").squeeze(1).long()


class FSCELoss(nn.Module):
    def __init__(self, configer=None):
        super().__init__()
        weight, reduction, ignore_index = _ce_params(configer)
        self.ce_loss = nn.CrossEntropyLoss(weight=weight, ignore_index=ignore_index, reduction=reduction)

    def forward(self, inputs, *targets, weights=None, **kwargs):
        if isinstance(inputs, (tuple, list)):
            weights = weights or [1.0] * len(inputs)
            loss = 0.0
            for i, x in enumerate(inputs):
                tgt = targets[i] if len(targets) > 1 else targets[0]
                loss = loss + weights[i] * self.ce_loss(x, _scale_target(tgt, (x.size(2), x.size(3))))
            return loss
        return self.ce_loss(inputs, _scale_target(targets[0], (inputs.size(2), inputs.size(3))))


class FSAuxCELoss(nn.Module):
    """seg_loss_weight * CE(seg) + aux_loss_weight * CE(aux)  (lib/loss/loss_helper.py:301-313)."""

    def __init__(self, configer=None):
        super().__init__()
        self.configer = configer
        self.ce_loss = FSCELoss(configer)

    def forward(self, inputs, targets, **kwargs):
        aux_out, seg_out = inputs
        w = self.configer.get("network", "loss_weights")
        return w["seg_loss"] * self.ce_loss(seg_out, targets) + w["aux_loss"] * self.ce_loss(aux_out, targets)
