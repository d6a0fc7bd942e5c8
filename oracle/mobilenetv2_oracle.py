"""CPU oracle for the reference's MobileNetV2 eval forward.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Functional, state_dict-driven restatement (torch CPU ops, fp32) of /root/reference/models/mobilenetv2.py:164-178 with the
module tree of :8-64 / :118-161.  Pinned against the reference itself by ``tests/golden/generate_golden_mnv2.py`` (imports
/root/reference/models, asserts this file reproduces it, commits ``tests/golden/mobilenetv2.npz``); nothing under
``uit_mobile_b200/`` imports it.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import uit_oracle as O

Tensor = torch.Tensor
SETTING = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1]]   # mobilenetv2.py:106-116


def _bn(x: Tensor, sd: Dict[str, Tensor], p: str) -> Tensor:
    """Eval BatchNorm2d, eps 1e-5 (nn.BatchNorm2d defaults; mobilenetv2.py:18-19)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def _conv_bn_relu6(x: Tensor, sd: Dict[str, Tensor], p: str, stride: int = 1, groups: int = 1) -> Tensor:
    """_ConvBNReLU (mobilenetv2.py:8-28): padding (k - 1) // 2, bias=False."""
    w = sd[p + ".0.weight"]
    return F.relu6(_bn(F.conv2d(x, w, None, stride, (w.shape[-1] - 1) // 2, 1, groups), sd, p + ".1"))


def features(x: Tensor, sd: Dict[str, Tensor], trace: Optional[List[Tensor]] = None) -> Tensor:
    """self.features without the pooling (mobilenetv2.py:118-142): [B, 1, 64, T] -> [B, 1280, 2, T']."""
    h = _conv_bn_relu6(x, sd, "features.0", stride=2)
    f, inp = 1, 32
    for t, c, n, s in SETTING:
        for i in range(n):
            stride, hidden, p = (s if i == 0 else 1), inp * t, f"features.{f}.conv"
            y, j = h, 0
            if t != 1:
                y, j = _conv_bn_relu6(y, sd, f"{p}.0"), 1                                   # pw expand
            y = _conv_bn_relu6(y, sd, f"{p}.{j}", stride=stride, groups=hidden)             # depthwise
            y = _bn(F.conv2d(y, sd[f"{p}.{j + 1}.weight"]), sd, f"{p}.{j + 2}")              # pw linear
            h = h + y if (stride == 1 and inp == c) else y                                  # mobilenetv2.py:60-64
            if trace is not None:
                trace.append(h)
            inp, f = c, f + 1
    return _conv_bn_relu6(h, sd, f"features.{f}")


def head(h: Tensor, sd: Dict[str, Tensor]) -> Tensor:
    """AdaptiveAvgPool2d((1, None)), flatten, Linear, sigmoid, mean over time (mobilenetv2.py:142, 174-177); Dropout is identity in eval."""
    h = h.mean(dim=2, keepdim=True).flatten(-2).transpose(1, 2)
    return torch.sigmoid(F.linear(h, sd["classifier.1.weight"], sd["classifier.1.bias"])).mean(1)


def forward(sd: Dict[str, Tensor], wav: Tensor) -> Tensor:
    """MobileNetV2.forward, eval branch.  wav [B, L] fp32 -> [B, outputdim]."""
    db = O.logmel(wav, sd["front_end.0.spectrogram.window"], sd["front_end.0.mel_scale.fb"])
    return head(features(db.unsqueeze(1), sd), sd)
