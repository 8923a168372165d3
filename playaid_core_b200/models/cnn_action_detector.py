"""`CNNActionDetector` on B200: drop-in for reference playaid/models/cnn_action_detector.py.

Same constructor arguments, attributes (`actions`, `num_actions`, `sequence_length`, `model`),
`load_from_checkpoint(path, actions=...)` (Lightning checkpoint: `ckpt["state_dict"]`, keys
`model.cnn2d.*`, `model.cnn1d.0.*`, `model.classifier.{0,2}.*`), `eval()`, and
`forward(x[B,S,3,128,128] in [0,1], RGB) -> [B, len(actions)]` log-probabilities
(reference :46-92). The arithmetic runs in the hand-written sm_100a kernels behind
`pa_features` / `pa_head`; inference only (training steps and dataloaders are out of scope).

Beyond the reference surface, `features()` / `head()` expose the two halves separately so that
callers can compute ResNet-18 features once per (frame, fighter) and reuse them across the seven
windows that contain a frame (the reference recomputes them 7x, playaid/ai_runner.py:493-520).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _lib

# name -> (native precision id, activations split hi+lo, IEEE half instead of bfloat16)
#   bf16 / f16       one tensor-core product per k-step (f16: 11-bit significand, same rate as bf16)
#   bf16x2 / f16x2   activations carry a rounding-residual plane: 2 products (f16x2 ~ fp32 accuracy)
#   bf16x3 / f16x3   weights split as well: 3 products, for checkpoints that are not 16-bit exact
PRECISIONS = {
    "bf16": (_lib.PREC_BF16, False, False), "bf16x2": (_lib.PREC_BF16X2, True, False), "bf16x3": (_lib.PREC_BF16X3, True, False),
    "f16": (_lib.PREC_F16, False, True), "f16x2": (_lib.PREC_F16X2, True, True), "f16x3": (_lib.PREC_F16X3, True, True),
}


class SpatialStreamCNN:
    """Holds the weights (host fp32 state_dict) and the native model handle."""

    def __init__(self, num_actions: int, sequence_length: int):
        self.num_actions = num_actions
        self.sequence_length = sequence_length
        self._state: dict[str, torch.Tensor] = {}

    def state_dict(self):
        return dict(self._state)


class CNNActionDetector:
    def __init__(
        self,
        actions: list,
        batch_size: int = 64,
        sequence_length: int = 4,
        learning_rate: float = 2e-4,
        num_samples: int = 1024,
        freeze_encoder=False,
        precision: str = "f16x2",
        device=None,
        **kwargs,
    ):
        self.learning_rate = learning_rate
        self.batch_size = batch_size
        self.actions = list(actions)
        self.num_actions = len(self.actions)
        self.num_samples = num_samples
        self.sequence_length = sequence_length
        self.dataset_kwargs = kwargs
        self.model = SpatialStreamCNN(self.num_actions, self.sequence_length)
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision = precision
        self.training = False
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._ctx = None
        self._handle = None
        self._ws: torch.Tensor | None = None
        self._ws_head: torch.Tensor | None = None

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict, strict: bool = True):
        """Accepts the reference's keys with or without the Lightning `model.` prefix."""
        sd = {}
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked") or "accuracy" in k:
                continue
            sd[k[6:] if k.startswith("model.") else k] = v.detach().to("cpu", torch.float32).contiguous()
        self.model._state = sd
        # The two-product modes are fp32-exact only when the WEIGHTS are representable in the 16-bit operand format (the
        # activations carry a second plane, the weights do not). A trained checkpoint is not: give the weights a residual
        # plane too (three products per k-step) instead of silently rounding them.
        if self.precision in ("f16x2", "bf16x2"):
            dt = torch.float16 if self.half else torch.bfloat16
            # "representable" up to 2^-20 of the tensor's largest weight: values below the format's normal range round
            # with an absolute error (<= 2^-25 for half) that is far below fp32 noise; a trained fp32 tensor misses by 2^-12
            inexact = [k for k, v in sd.items() if v.ndim >= 2 and
                       float((v.to(dt).float() - v).abs().max()) > max(2.0 ** -20 * float(v.abs().max()), 2.0 ** -24)]
            if inexact:
                import warnings

                self.precision = self.precision[:-1] + "3"
                warnings.warn(f"{len(inexact)} weight tensors are not exactly representable in {dt}: using precision "
                              f"'{self.precision}' (weights split as well) to keep fp32 parity", stacklevel=2)
        self._finalize()
        return self

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, actions=None, map_location=None, **kwargs):
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(kwargs)
        if actions is not None:
            hp["actions"] = actions
        sd = ckpt["state_dict"]
        if "sequence_length" not in hp:
            key = "model.cnn1d.0.weight" if "model.cnn1d.0.weight" in sd else "cnn1d.0.weight"
            hp["sequence_length"] = int(sd[key].shape[2])
        obj = cls(**hp)
        obj.load_state_dict(sd)
        return obj

    def state_dict(self):
        return {"model." + k: v for k, v in self.model._state.items()}

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("playaid_core_b200 is inference-only (training is out of scope)")
        return self.eval()

    def to(self, device):
        self._device = torch.device(device)
        if self.model._state:
            self._finalize()
        return self

    def _finalize(self):
        if self._device.type != "cuda":
            raise _lib.PlayaidLibraryError("CNNActionDetector needs a CUDA device (no CPU path)")
        self._ctx = _lib.Context.get(self._device)
        lib = self._ctx.lib
        if self._handle is not None:
            lib.pa_model_destroy(self._handle)
            self._handle = None
        h = ctypes.c_void_p()
        _lib.check(lib.pa_model_create(self._ctx.handle, self.num_actions, self.sequence_length, ctypes.byref(h)),
                   self._ctx.handle, "pa_model_create")
        for k, v in self.model._state.items():
            a = v.numpy()
            shape = (ctypes.c_int64 * max(a.ndim, 1))(*a.shape)
            _lib.check(lib.pa_model_set_tensor(h, k.encode(), a.ctypes.data, shape, a.ndim), self._ctx.handle, f"set_tensor {k}")
        with torch.cuda.device(self._device):
            _lib.check(lib.pa_model_finalize(h, PRECISIONS[self.precision][0]), self._ctx.handle, "pa_model_finalize")
        self._handle = h

    def __del__(self):
        try:
            if self._handle is not None and self._ctx is not None:
                self._ctx.lib.pa_model_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ native halves
    @property
    def split(self) -> bool:
        return PRECISIONS[self.precision][1]

    @property
    def half(self) -> bool:
        return PRECISIONS[self.precision][2]

    @property
    def act_dtype(self) -> torch.dtype:
        return torch.float16 if self.half else torch.bfloat16

    @property
    def crop_dtype(self) -> int:
        """pa_preprocess out_dtype that matches this model's arithmetic."""
        if self.half:
            return _lib.DTYPE_F16X2 if self.split else _lib.DTYPE_F16
        return _lib.DTYPE_BF16X2 if self.split else _lib.DTYPE_BF16

    def _workspace(self, n: int) -> torch.Tensor:
        need = ctypes.c_size_t()
        _lib.check(self._ctx.lib.pa_model_workspace_bytes(self._handle, n, ctypes.byref(need)), self._ctx.handle, "workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = torch.empty((need.value,), dtype=torch.uint8, device=self._device)
        return self._ws

    @property
    def crop_dtype_u8(self) -> int:
        """pa_preprocess out_dtype for byte-valued crops (`features(..., u8=True)`): the resampled byte itself, exact in
        16 bits, one plane in every precision; the stem applies the 1/255 of ai_runner.py:463 in fp32."""
        return _lib.DTYPE_F16_U8 if self.half else _lib.DTYPE_BF16_U8

    def features(self, crops: torch.Tensor, out: torch.Tensor | None = None, u8: bool = False) -> torch.Tensor:
        """crops: 16-bit (self.act_dtype) CUDA NHWC4P [n,128,136,4] -- 4 channels (last zero), 4 zero
        pixels either side of every row -- or [2,n,128,136,4] hi/lo planes in split precision, as written by
        `preprocess_crops(dtype=self.crop_dtype, layout=LAYOUT_NHWC4P)` -> fp32 [n,1000].
        `u8=True`: crops hold byte values 0..255 (`dtype=self.crop_dtype_u8`, always one plane)."""
        if self._handle is None:
            raise _lib.PlayaidLibraryError("no weights loaded: call load_state_dict / load_from_checkpoint first")
        want = 5 if (self.split and not u8) else 4
        if crops.dtype != self.act_dtype or crops.ndim != want or not crops.is_contiguous() or tuple(crops.shape[-3:]) != (128, 136, 4):
            raise ValueError(f"crops must be contiguous {self.act_dtype} NHWC4P [n,128,136,4] ([2,n,...] planes in split precision)")
        n = int(crops.shape[-4])
        if out is None:
            out = torch.empty((n, 1000), dtype=torch.float32, device=crops.device)
        ws = self._workspace(n)
        with torch.cuda.device(crops.device):
            fn = self._ctx.lib.pa_features_u8 if u8 else self._ctx.lib.pa_features
            rc = fn(self._handle, crops.data_ptr(), n, out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.current_stream_ptr(crops.device))
        _lib.check(rc, self._ctx.handle, "pa_features")
        return out

    def head(self, feat: torch.Tensor, win_idx: torch.Tensor, status: torch.Tensor | None = None):
        """feat fp32 [n_feat,1000], win_idx int32 [n_win,S] rows of feat ->
        (logp [n_win,A] fp32, label [n_win] int32, prob [n_win] fp32). `status` int32 [n_feat]: per-crop status
        of the feature rows (pa_preprocess); windows that touch a crop which is not OK get label -1 / prob 0."""
        if self._handle is None:
            raise _lib.PlayaidLibraryError("no weights loaded")
        if feat.dtype != torch.float32 or not feat.is_contiguous() or feat.shape[-1] != 1000:
            raise ValueError("feat must be contiguous fp32 [n,1000]")
        if win_idx.dtype != torch.int32 or not win_idx.is_contiguous() or win_idx.shape[-1] != self.sequence_length:
            raise ValueError("win_idx must be contiguous int32 [n_win, sequence_length]")
        n_feat, n_win = int(feat.shape[0]), int(win_idx.shape[0])
        if status is not None and (status.dtype != torch.int32 or not status.is_contiguous() or status.numel() != n_feat):
            raise ValueError("status must be contiguous int32 [n_feat]")
        logp = torch.empty((n_win, self.num_actions), dtype=torch.float32, device=feat.device)
        label = torch.empty((n_win,), dtype=torch.int32, device=feat.device)
        prob = torch.empty((n_win,), dtype=torch.float32, device=feat.device)
        # its own workspace: the head of one chunk may run on a side stream while pa_features of the next chunk uses the other
        need = ctypes.c_size_t()
        _lib.check(self._ctx.lib.pa_model_workspace_bytes(self._handle, max(n_feat, 1), ctypes.byref(need)), self._ctx.handle, "workspace_bytes")
        if self._ws_head is None or self._ws_head.numel() < need.value:
            self._ws_head = torch.empty((need.value,), dtype=torch.uint8, device=self._device)
        ws = self._ws_head
        with torch.cuda.device(feat.device):
            rc = self._ctx.lib.pa_head(self._handle, feat.data_ptr(), n_feat, status.data_ptr() if status is not None else None,
                                       win_idx.data_ptr(), n_win, logp.data_ptr(),
                                       label.data_ptr(), prob.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _lib.current_stream_ptr(feat.device))
        _lib.check(rc, self._ctx.handle, "pa_head")
        return logp, label, prob

    # ------------------------------------------------------------------ reference surface
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,S,3,H,W] float in [0,1] (RGB, as built at ai_runner.py:461-463) -> log-probs [B,A]."""
        B, S, C, H, W = x.shape
        if S != self.sequence_length or C != 3 or H != 128 or W != 128:
            raise ValueError(f"expected [B,{self.sequence_length},3,128,128], got {tuple(x.shape)}")
        x = x.to(self._device, torch.float32).reshape(B * S, 3, H, W).permute(0, 2, 3, 1)
        x4 = torch.zeros((B * S, H, W + 8, 4), dtype=torch.float32, device=self._device)
        x4[:, :, 4 : W + 4, :3] = x
        hi = x4.to(self.act_dtype)
        if self.split:
            lo = (x4 - hi.float()).to(self.act_dtype)
            crops = torch.stack([hi, lo]).contiguous()
        else:
            crops = hi.contiguous()
        feat = self.features(crops)
        idx = torch.arange(B * S, dtype=torch.int32, device=self._device).reshape(B, S).contiguous()
        logp, _, _ = self.head(feat, idx)
        return logp

    __call__ = forward
