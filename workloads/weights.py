"""Seeded random-init weights of the reference architecture.

* `default_state_dict(seed)` -- the reference's own construction order under `torch.manual_seed`
  (torchvision `resnet18(weights=None)`, then `nn.Conv1d(1000, 512, S)`, then the two `nn.Linear`s:
  playaid/models/cnn_action_detector.py:16-27), i.e. exactly the tensors the reference would hold with
  no checkpoint. This net is degenerate (one predicted class, top-2 margins ~0.01).
* `calibrated_state_dict(seed)` -- same init, then the calibration recipe that makes label parity
  non-vacuous: conv / linear weights rounded to bf16-representable values (shared by oracle and
  GPU path), BatchNorm running stats set from one train-mode pass over a seeded synthetic
  calibration batch (`momentum=None`), last layer centred on the calibration logits and scaled
  (by a power of two) to a logit std of ~4.

Workload generation only: torch CPU ops are used here to *make* weights, never to classify.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import synthetic

NUM_ACTIONS = 63
SEQ = 7


class _Net(nn.Module):
    def __init__(self, num_actions, seq):
        super().__init__()
        from torchvision.models import resnet18

        self.cnn2d = resnet18(weights=None)
        self.cnn1d = nn.Sequential(nn.Conv1d(1000, 512, kernel_size=seq, stride=1), nn.ReLU())
        self.classifier = nn.Sequential(nn.Linear(512, 128), nn.ReLU(), nn.Linear(128, num_actions))

    def forward(self, x):  # [B,S,3,H,W] -> logits [B,A]
        B, S, C, H, W = x.shape
        f = self.cnn2d(x.view(B * S, C, H, W)).view(B, S, -1).permute(0, 2, 1)
        return self.classifier(self.cnn1d(f).view(B, -1))


def _to_state_dict(net: nn.Module) -> dict:
    return {"model." + k: v.detach().clone() for k, v in net.state_dict().items()}


def default_state_dict(seed: int = 0, num_actions: int = NUM_ACTIONS, seq: int = SEQ) -> dict:
    torch.manual_seed(seed)
    return _to_state_dict(_Net(num_actions, seq))


def calibration_windows(n_windows: int = 48, seq: int = SEQ, seed: int = 99) -> torch.Tensor:
    """[n,S,3,128,128] float in [0,1]: area-downsampled boxes of small synthetic frames (statistics
    only -- exact crop semantics are irrelevant for calibration)."""
    n_frames = n_windows + 6 * (seq // 2) ** 2 // 3
    rng = np.random.default_rng(seed)
    Hc, Wc = 540, 960
    cx = np.clip(0.5 + np.cumsum(rng.normal(0, 0.01, (n_frames, 2)), 0), 0.25, 0.75)
    cy = np.clip(0.55 + np.cumsum(rng.normal(0, 0.005, (n_frames, 2)), 0), 0.35, 0.65)
    boxes_px = np.stack([cx * Wc, cy * Hc, np.full_like(cx, 125), np.full_like(cx, 145)], -1).astype(np.int64)
    frames = synthetic.synth_frames(np.arange(n_frames), boxes_px, H=Hc, W=Wc, device="cpu", seed=seed)
    crops = []
    for i in range(n_frames):
        x, y = int(boxes_px[i, 0, 0]), int(boxes_px[i, 0, 1])
        win = frames[i, y - 96 : y + 96, x - 96 : x + 96].permute(2, 0, 1)[None].float()
        crops.append(F.interpolate(win, size=(128, 128), mode="area")[0].flip(0) / 255.0)
    crops = torch.stack(crops)  # [n_frames,3,128,128] RGB
    mid = seq // 2
    idx = torch.tensor([[min(max(i + int(math.copysign(3 * (k - mid) ** 2, k - mid)), 0), n_frames - 1) for k in range(seq)]
                        for i in range(n_windows)])
    return crops[idx]


_CACHE: dict = {}


@torch.no_grad()
def calibrated_state_dict(seed: int = 0, num_actions: int = NUM_ACTIONS, seq: int = SEQ, logit_std: float = 4.0,
                          round_bf16: bool = True, calib: torch.Tensor | None = None) -> dict:
    """Cached per process (treat the returned tensors as read-only)."""
    key = (seed, num_actions, seq, logit_std, round_bf16)
    if calib is None and key in _CACHE:
        return _CACHE[key]
    sd = _calibrated_state_dict(seed, num_actions, seq, logit_std, round_bf16, calib)
    if calib is None:
        _CACHE[key] = sd
    return sd


def _calibrated_state_dict(seed, num_actions, seq, logit_std, round_bf16, calib) -> dict:
    torch.manual_seed(seed)
    net = _Net(num_actions, seq)
    if round_bf16:
        for m in net.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv1d, nn.Linear)):
                m.weight.copy_(m.weight.to(torch.bfloat16).float())
    if calib is None:
        calib = calibration_windows(seq=seq)
    bns = [m for m in net.modules() if isinstance(m, nn.BatchNorm2d)]
    for bn in bns:
        bn.reset_running_stats()
        bn.momentum = None  # cumulative average over the single pass
    net.train()
    net(calib)
    net.eval()
    logits = net(calib)
    last = net.classifier[2]
    last.bias.sub_(logits.mean(0))
    s = float((logits - logits.mean(0)).std())
    scale = 2.0 ** round(math.log2(logit_std / max(s, 1e-12)))
    last.weight.mul_(scale)
    last.bias.mul_(scale)
    for bn in bns:
        bn.momentum = 0.1
    return _to_state_dict(net)
