"""`ResnetTransformerDetector` on B200: drop-in for reference playaid/models/resnet_transformer_detector.py
(SURVEY 8f rank 2 -- the model `action_detector.py` trains).

Same constructor arguments and attributes (`actions`, `num_actions`, `sequence_length`, `model`), Lightning
checkpoint loading (`ckpt["state_dict"]`, keys `model.resnet.*`, `model.resnet_ffn.*`, `model.freq_encoding`,
`model.transformer.layers.{i}.*`, `model.classifier.*`), `eval()`, and
`forward(frames[B,S,3,128,128] in [0,1], RGB) -> [B,S,len(actions)]` log-probabilities (reference :136-143).
The arithmetic runs behind `pa_resformer_forward`: fused stem, tcgen05 implicit-GEMM ResNet-50 bottlenecks, every
Linear of the encoder as a tensor-core GEMM over the B*S tokens, attention / LayerNorm / softmax in small fp32 kernels.

The reference builds its encoder with `batch_first=False` yet feeds it [B,S,256]: attention runs across the B windows
of a batch (per slot), not across the S frames of a window. That is kept -- `forward` on a batch reproduces the
reference's output for that same batch, including its dependence on batch composition.
Inference only; `resnet_classifier` (unused by the reference's forward) is accepted and ignored.
"""
from __future__ import annotations

import ctypes

import torch

from .. import _lib
from .cnn_action_detector import PRECISIONS


class ResFormer:
    """Holds the host fp32 state_dict (the reference's `self.model`)."""

    def __init__(self, num_actions: int, sequence_length: int):
        self.num_actions = num_actions
        self.sequence_length = sequence_length
        self.hidden_dim = 247
        self._state: dict[str, torch.Tensor] = {}

    def state_dict(self):
        return dict(self._state)


class ResnetTransformerDetector:
    def __init__(self, actions: list, batch_size: int = 64, sequence_length: int = 4, learning_rate: float = 2e-4,
                 num_samples: int = 1024, freeze_encoder=False, precision: str = "f16x2", device=None, **kwargs):
        self.learning_rate = learning_rate
        self.batch_size = batch_size
        self.actions = list(actions)
        self.num_actions = len(self.actions)
        self.num_samples = num_samples
        self.sequence_length = sequence_length
        self.dataset_kwargs = kwargs
        self.model = ResFormer(self.num_actions, self.sequence_length)
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision = precision
        self.training = False
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._ctx = None
        self._handle = None
        self._ws: torch.Tensor | None = None

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict, strict: bool = True):
        sd = {}
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked") or "accuracy" in k:
                continue
            k = k[6:] if k.startswith("model.") else k
            if k.startswith("encoder_layer.") or k.startswith("resnet_classifier."):
                continue   # the template layer the encoder was cloned from, and a head the forward never uses
            sd[k] = v.detach().to("cpu", torch.float32).contiguous()
        self.model._state = sd
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
            key = "model.freq_encoding" if "model.freq_encoding" in sd else "freq_encoding"
            hp["sequence_length"] = int(sd[key].shape[0])
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

    def _finalize(self):
        if self._device.type != "cuda":
            raise _lib.PlayaidLibraryError("ResnetTransformerDetector needs a CUDA device (no CPU path)")
        self._ctx = _lib.Context.get(self._device)
        lib = self._ctx.lib
        if self._handle is not None:
            lib.pa_model_destroy(self._handle)
            self._handle = None
        h = ctypes.c_void_p()
        _lib.check(lib.pa_resformer_create(self._ctx.handle, self.num_actions, self.sequence_length, ctypes.byref(h)),
                   self._ctx.handle, "pa_resformer_create")
        for k, v in self.model._state.items():
            a = v.numpy()
            shape = (ctypes.c_int64 * max(a.ndim, 1))(*a.shape)
            _lib.check(lib.pa_model_set_tensor(h, k.encode(), a.ctypes.data, shape, a.ndim), self._ctx.handle, f"set_tensor {k}")
        with torch.cuda.device(self._device):
            _lib.check(lib.pa_resformer_finalize(h, PRECISIONS[self.precision][0]), self._ctx.handle, "pa_resformer_finalize")
        self._handle = h

    def __del__(self):
        try:
            if self._handle is not None and self._ctx is not None:
                self._ctx.lib.pa_model_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ native forward
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
        if self.half:
            return _lib.DTYPE_F16X2 if self.split else _lib.DTYPE_F16
        return _lib.DTYPE_BF16X2 if self.split else _lib.DTYPE_BF16

    def forward_crops(self, crops: torch.Tensor, n_windows: int) -> torch.Tensor:
        """crops: 16-bit NHWC4P [n_windows*S,128,136,4] ([2,...] planes in split precision), window-major, as
        `preprocess_crops(dtype=self.crop_dtype, layout=LAYOUT_NHWC4P)` writes them -> log-probs [n_windows,S,A]."""
        if self._handle is None:
            raise _lib.PlayaidLibraryError("no weights loaded: call load_state_dict / load_from_checkpoint first")
        want = 5 if self.split else 4
        if crops.dtype != self.act_dtype or crops.ndim != want or not crops.is_contiguous() or tuple(crops.shape[-3:]) != (128, 136, 4):
            raise ValueError(f"crops must be contiguous {self.act_dtype} NHWC4P [n,128,136,4] ([2,n,...] planes in split precision)")
        S = self.sequence_length
        if int(crops.shape[-4]) != n_windows * S:
            raise ValueError("crops must hold n_windows * sequence_length crops")
        need = ctypes.c_size_t()
        lib = self._ctx.lib
        _lib.check(lib.pa_resformer_workspace_bytes(self._handle, n_windows, ctypes.byref(need)), self._ctx.handle, "workspace_bytes")
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = torch.empty((need.value,), dtype=torch.uint8, device=self._device)
        logp = torch.empty((n_windows, S, self.num_actions), dtype=torch.float32, device=crops.device)
        with torch.cuda.device(crops.device):
            rc = lib.pa_resformer_forward(self._handle, crops.data_ptr(), n_windows, logp.data_ptr(), self._ws.data_ptr(),
                                          self._ws.numel(), _lib.current_stream_ptr(crops.device))
        _lib.check(rc, self._ctx.handle, "pa_resformer_forward")
        return logp

    def forward(self, frames: torch.Tensor) -> torch.Tensor:
        """frames [B,S,3,128,128] float in [0,1] (RGB) -> log-probs [B,S,A] (reference :136-143)."""
        B, S, C, H, W = frames.shape
        if S != self.sequence_length or C != 3 or H != 128 or W != 128:
            raise ValueError(f"expected [B,{self.sequence_length},3,128,128], got {tuple(frames.shape)}")
        x = frames.to(self._device, torch.float32).reshape(B * S, 3, H, W).permute(0, 2, 3, 1)
        x4 = torch.zeros((B * S, H, W + 8, 4), dtype=torch.float32, device=self._device)
        x4[:, :, 4 : W + 4, :3] = x
        hi = x4.to(self.act_dtype)
        if self.split:
            lo = (x4 - hi.float()).to(self.act_dtype)
            crops = torch.stack([hi, lo]).contiguous()
        else:
            crops = hi.contiguous()
        return self.forward_crops(crops, B)

    __call__ = forward
