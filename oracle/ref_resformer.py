"""CPU restatement of the reference's second detector, `ResnetTransformerDetector` / `ResFormer`
(playaid/models/resnet_transformer_detector.py:17-143) -- SURVEY 8f rank 2, the model `action_detector.py` trains.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.

What the reference computes (file:line refer to that module):
  * frames [B,S,3,H,W] -> (B*S) crops through timm's "resnet50" with num_classes=0, i.e. the torchvision v1.5
    bottleneck ResNet up to global average pooling: 2048 features per crop (:34, :71-76);
  * Linear(2048, 247) (:38, :77);
  * nine time-encoding features per slot appended: t = linspace(0,1,S); [t, cos(pi t 2^i), sin(pi t 2^i)] for
    i = 0..3 (:18-23, :40-46, :82-85) -> d_model = 256;
  * three post-norm `nn.TransformerEncoderLayer`s (8 heads of 32, feed-forward 2048, ReLU, eps 1e-5) (:50-57, :87).
    They are built with the default `batch_first=False`, and the tensor handed to them is [B,S,256]: the encoder
    therefore treats the B windows of the batch as the *sequence* and the S slots as the batch -- attention mixes
    the windows of a batch slot by slot, and a window's output depends on which other windows share its batch
    (measured: 1.7e-2 in log-prob between batch 1 and batch 2 on the seed-0 net). A drop-in keeps that;
  * Linear(256, A) per token and `log_softmax` over the classes (:90-96, :142-143) -> [B,S,A].
`resnet_classifier` (:61) takes no part in the forward.

`RefResFormer` composes the same torch modules with the reference's parameter names (so a Lightning checkpoint's
`state_dict` loads with the "model." prefix); `forward_explicit` re-derives the encoder with plain matmuls as the
spec a CUDA implementation follows, and tests pin both against goldens generated from the reference itself
(oracle/gen_golden.py::gen_resformer).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def time_encoding(sequence_length: int, num_freq: int = 4) -> torch.Tensor:
    t = torch.linspace(0, 1, sequence_length).reshape(-1, 1)
    cols = [t]
    for i in range(num_freq):
        cols += [torch.cos(math.pi * t * (2 ** i)), torch.sin(math.pi * t * (2 ** i))]
    return torch.cat(cols, dim=1)   # [S, 1 + 2*num_freq]


class RefResFormer(nn.Module):
    def __init__(self, num_actions: int = 63, sequence_length: int = 7, hidden_dim: int = 247, num_heads: int = 8,
                 num_layers: int = 3):
        super().__init__()
        import torchvision

        self.num_actions, self.sequence_length, self.hidden_dim = num_actions, sequence_length, hidden_dim
        self.num_heads, self.num_layers = num_heads, num_layers
        self.resnet = torchvision.models.resnet50(weights=None)
        self.resnet.fc = nn.Identity()
        self.resnet_ffn = nn.Linear(2048, hidden_dim)
        self.register_buffer("freq_encoding", time_encoding(sequence_length))
        self.d_model = hidden_dim + self.freq_encoding.shape[1]
        self.encoder_layer = nn.TransformerEncoderLayer(d_model=self.d_model, nhead=num_heads)
        self.transformer = nn.TransformerEncoder(self.encoder_layer, num_layers=num_layers, enable_nested_tensor=False)
        self.resnet_classifier = nn.Linear(hidden_dim, num_actions)   # present in checkpoints, unused in forward
        self.classifier = nn.Linear(self.d_model, num_actions)

    def tokens(self, frames: torch.Tensor) -> torch.Tensor:
        B, S = frames.shape[:2]
        feat = self.resnet(frames.reshape(B * S, *frames.shape[2:]))
        x = self.resnet_ffn(feat).reshape(B, S, self.hidden_dim)
        return torch.cat([x, self.freq_encoding.unsqueeze(0).expand(B, -1, -1)], dim=2)   # [B,S,256]

    def forward(self, frames: torch.Tensor) -> torch.Tensor:
        """[B,S,3,H,W] float in [0,1] -> log-probs [B,S,A]."""
        y = self.transformer(self.tokens(frames))            # batch_first=False: B is the sequence axis
        return F.log_softmax(self.classifier(y), dim=2)

    @torch.no_grad()
    def forward_explicit(self, frames: torch.Tensor) -> torch.Tensor:
        """The same forward with the encoder written out (eval mode, no dropout)."""
        x = self.tokens(frames)                               # [L=B, N=S, E]
        L, N, E = x.shape
        h, dh = self.num_heads, E // self.num_heads
        for layer in self.transformer.layers:
            att = layer.self_attn
            qkv = x @ att.in_proj_weight.T + att.in_proj_bias                    # [L,N,3E]
            q, k, v = [t.reshape(L, N, h, dh) for t in qkv.split(E, dim=2)]
            # scores over the L axis (the windows of the batch), separately per slot n and head
            s = torch.einsum("lnhd,mnhd->nhlm", q, k) / math.sqrt(dh)
            p = torch.softmax(s, dim=-1)
            o = torch.einsum("nhlm,mnhd->lnhd", p, v).reshape(L, N, E)
            o = o @ att.out_proj.weight.T + att.out_proj.bias
            x = F.layer_norm(x + o, (E,), layer.norm1.weight, layer.norm1.bias, layer.norm1.eps)
            f = torch.relu(x @ layer.linear1.weight.T + layer.linear1.bias) @ layer.linear2.weight.T + layer.linear2.bias
            x = F.layer_norm(x + f, (E,), layer.norm2.weight, layer.norm2.bias, layer.norm2.eps)
        return F.log_softmax(x @ self.classifier.weight.T + self.classifier.bias, dim=2)


class RefResnetTransformerDetector(nn.Module):
    """Same constructor arguments, attributes and `state_dict` keys ("model.*") as the reference LightningModule."""

    def __init__(self, actions: list, batch_size: int = 64, sequence_length: int = 4, learning_rate: float = 2e-4,
                 num_samples: int = 1024, freeze_encoder=False, **kwargs):
        super().__init__()
        self.actions, self.num_actions, self.sequence_length = actions, len(actions), sequence_length
        self.model = RefResFormer(self.num_actions, sequence_length)

    def forward(self, frames):
        return self.model(frames)
