"""Stand-in trainable networks for the benchmarks, examples and tests.

The trainable part of PPNN stays plain PyTorch (north_star; SURVEY.md 2.1 #7 marks
``src/models.py`` out of scope), and the reference's own ``models.py`` cannot travel to
the GPU box.  These modules have the layer shapes of the reference's ``cnn2d`` /
``cnn2dRich`` (models.py:77-124: two strided 6x6 convolutions down to 64x64, three
48-channel 5x5 residual blocks with LayerNorm, PixelShuffle(4), a 5x5 head, and a
rank-one learned embedding of the scalar parameter), so a rollout step around them
costs what a reference rollout step costs.  Weights are random; nothing here is on
the PDE-preserving hot path.
"""
from __future__ import annotations

import torch
from torch import nn


class ResBlock(nn.Module):
    def __init__(self, channels: int, ksize: int, grid):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(channels, channels, ksize, padding=ksize // 2), nn.ReLU(),
                                  nn.Conv2d(channels, channels, ksize, padding=ksize // 2))
        self.norm = nn.LayerNorm([channels] + list(grid))

    def forward(self, x):
        return self.norm(self.body(x)) + x


class ConvResNet(nn.Module):
    """``uses_pde=True``: inputs (u0 [B,2,256,256], mu [B,1,1,1], pdeu [B,2,256,256]) like
    cnn2dRich; ``uses_pde=False``: (u0, mu) like cnn2d.  Output [B,2,256,256]."""

    def __init__(self, uses_pde: bool = True, grid: int = 256, width: int = 48):
        super().__init__()
        self.uses_pde = uses_pde
        cin = 5 if uses_pde else 3
        g = grid // 4
        self.net = nn.Sequential(
            nn.Conv2d(cin, 12, 6, stride=2, padding=2), nn.ReLU(),
            nn.Conv2d(12, width, 6, stride=2, padding=2), nn.ReLU(),
            ResBlock(width, 5, (g, g)), ResBlock(width, 5, (g, g)), ResBlock(width, 5, (g, g)),
            nn.PixelShuffle(4),
            nn.Conv2d(width // 16, 2, 5, padding=2))
        self.col = nn.Parameter(torch.randn(1, 1, 1, grid))
        self.row = nn.Parameter(torch.randn(1, 1, grid, 1))

    def forward(self, u0, mu, pdeu=None):
        emb = mu * self.row @ self.col
        parts = (u0, emb, pdeu) if self.uses_pde else (u0, emb)
        return self.net(torch.cat(parts, dim=1))
