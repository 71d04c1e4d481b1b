"""ProjectionHead with the engine's L2-normalise (lib/models/modules/projection.py:8-24).

The 1x1 convolutions / BatchNorm stay cuDNN / PyTorch (not on the replaced path); only the
final F.normalize(p=2, dim=1) runs on the engine kernel.  ``normalize=False`` returns the raw projection for the
fused path (PixelContrastLoss(..., normalize=True) then normalises only the sampled columns)."""
from __future__ import annotations

import torch.nn as nn

from .functional import l2_normalize


class ProjectionHead(nn.Module):
    def __init__(self, dim_in, proj_dim=256, proj="convmlp", bn_type="torchsyncbn", normalize=True):
        super().__init__()
        if proj == "linear":
            self.proj = nn.Conv2d(dim_in, proj_dim, kernel_size=1)
        elif proj == "convmlp":
            bn = nn.SyncBatchNorm(dim_in) if bn_type == "torchsyncbn" else nn.BatchNorm2d(dim_in)
            self.proj = nn.Sequential(nn.Conv2d(dim_in, dim_in, kernel_size=1), nn.Sequential(bn, nn.ReLU()),
                                      nn.Conv2d(dim_in, proj_dim, kernel_size=1))
        else:
            raise ValueError(f"unknown projection type {proj!r}")
        self.normalize = normalize

    def forward(self, x):
        y = self.proj(x)
        return l2_normalize(y) if self.normalize else y
