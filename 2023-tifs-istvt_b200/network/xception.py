"""Parameter tree of the Xception backbone + the B200 entry-flow forward.

Mirrors the attribute names of the reference's `network/xception.py` (`Xception` :104-149, `Block` :52-101,
`SeparableConv2d` :39-49) so that checkpoints written by the reference load with `strict=True` (the key
list is SURVEY.md Appendix A).  The entry flow (`low_level_features`, reference :193-206) is on the
ISTVT path; the whole backbone (`features` / `logits` / `forward`, reference :161-226) serves the per-frame
'xception' baseline of the same evaluation script (train_CNN.py:924-929) on the same kernels, inference only.

The arithmetic lives in the CUDA library (see `engine.py`); nothing here computes with torch ops.
"""
from __future__ import annotations

import torch
from torch import nn


class SeparableConv2d(nn.Module):
    """depthwise kxk (`conv1`) + pointwise 1x1 (`pointwise`), both bias-free."""

    def __init__(self, cin: int, cout: int, kernel_size: int = 1, stride: int = 1, padding: int = 0):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cin, kernel_size, stride, padding, groups=cin, bias=False)
        self.pointwise = nn.Conv2d(cin, cout, 1, bias=False)


def _rep_layers(cin: int, cout: int, reps: int, stride: int, start_with_relu: bool, grow_first: bool) -> list:
    """[ReLU, Sep, BN] * reps (+ MaxPool) with the channel growth placed first or last."""
    widths = []
    c = cin
    for i in range(reps):
        grow_here = (i == 0) if grow_first else (i == reps - 1)
        nxt = cout if grow_here else c
        widths.append((c, nxt))
        c = nxt
    layers: list = []
    for a, b in widths:
        layers += [nn.ReLU(inplace=False), SeparableConv2d(a, b, 3, 1, 1), nn.BatchNorm2d(b)]
    if not start_with_relu:
        layers = layers[1:]
    if stride != 1:
        layers.append(nn.MaxPool2d(3, stride, 1))
    return layers


class Block(nn.Module):
    def __init__(self, cin: int, cout: int, reps: int, stride: int = 1, start_with_relu: bool = True,
                 grow_first: bool = True):
        super().__init__()
        if cout != cin or stride != 1:
            self.skip = nn.Conv2d(cin, cout, 1, stride=stride, bias=False)
            self.skipbn = nn.BatchNorm2d(cout)
        else:
            self.skip = None
        self.rep = nn.Sequential(*_rep_layers(cin, cout, reps, stride, start_with_relu, grow_first))
        self.start_with_relu = start_with_relu


class Xception(nn.Module):
    """Creation order follows the reference (conv1, bn1, conv2, bn2, block1..12, conv3, bn3, conv4, bn4, fc)
    so that a seeded construction draws the same random initial weights."""

    def __init__(self, num_classes: int = 1000):
        super().__init__()
        self.num_classes = num_classes
        self.conv1 = nn.Conv2d(3, 32, 3, 2, 0, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        self.conv2 = nn.Conv2d(32, 64, 3, bias=False)
        self.bn2 = nn.BatchNorm2d(64)
        self.block1 = Block(64, 128, 2, 2, start_with_relu=False)
        self.block2 = Block(128, 256, 2, 2)
        self.block3 = Block(256, 728, 2, 2)
        for i in range(4, 12):
            setattr(self, f"block{i}", Block(728, 728, 3, 1))
        self.block12 = Block(728, 1024, 2, 2, grow_first=False)
        self.conv3 = SeparableConv2d(1024, 1536, 3, 1, 1)
        self.bn3 = nn.BatchNorm2d(1536)
        self.conv4 = SeparableConv2d(1536, 2048, 3, 1, 1)
        self.bn4 = nn.BatchNorm2d(2048)
        fc = nn.Linear(2048, num_classes)  # drawn for RNG parity with the reference, replaced by the caller
        self.last_linear = fc

    # ---- the only forward on the ISTVT path ----
    def low_level_features(self, x: torch.Tensor) -> torch.Tensor:
        """x: fp32 NCHW [n, 3, H, W] on CUDA -> fp32 NCHW [n, 728, h, w] (reference layout, for parity checks).

        The fused model does not call this: `XceptionVidTr.forward` keeps the NHWC result inside the token
        buffer.  Here the NHWC tensor is returned as a permuted view.
        """
        from ..engine import entry_flow_features
        return entry_flow_features(self, x)

    # ---- the per-frame baseline: whole backbone ----
    precision = "bf16"
    input_norm = ((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))      # used for uint8 [N, H, W, 3] inputs only

    def _xengine(self):
        from ..engine import XceptionEngine
        eng = self.__dict__.get("_xeng")
        if eng is None:
            eng = XceptionEngine()
            self.__dict__["_xeng"] = eng
        return eng

    def features(self, x: torch.Tensor) -> torch.Tensor:
        """reference :161-191 — fp32 NCHW [n, 2048, h', w'] (bn4 output), for parity with the reference layout."""
        feats = self._xengine().forward(self, x, precision=self.precision, features_only=True)
        return feats.float().permute(0, 3, 1, 2)

    def logits(self, features: torch.Tensor) -> torch.Tensor:
        """reference :208-221 on an NCHW feature tensor: ReLU, global average pool, last_linear."""
        from .. import ops
        fc = self.last_linear[-1] if isinstance(self.last_linear, nn.Sequential) else self.last_linear
        x = features.permute(0, 2, 3, 1).contiguous()
        return ops.pool_linear(x, fc.weight.detach().float().contiguous(), fc.bias.detach().float().contiguous())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """reference :223-226 — frames [N, 3, H, W] fp32 (or uint8 [N, H, W, 3]) on CUDA -> logits fp32 [N, classes]."""
        return self._xengine().forward(self, x, precision=self.precision)

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_xeng", None)
        return super()._apply(fn, *args, **kwargs)
