"""TEST INFRASTRUCTURE ONLY -- import shims that let the *real* reference run in this container.

The reference (/root/reference/playaid) pins third-party packages that are absent here
(imutils, addict, pytorch_lightning, torchmetrics, albumentations, dictdiffer, timm).
`install()` registers minimal stand-ins in `sys.modules` and puts /root/reference on sys.path so
that `oracle/gen_golden.py` can import `playaid.fighter`, `playaid.timeline`,
`playaid.dataset_utils` and `playaid.models.cnn_action_detector` unmodified and record golden
vectors under tests/golden/.

Nothing in the product package, the `-m gpu` tests, smoke() or bench.py imports this file:
/root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PLAYAID_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "playaid"))


class _Dict(dict):
    """addict.Dict stand-in: auto-vivifying attribute dict (behaviour the reference relies on:
    missing key -> empty Dict, empty Dict is falsy, `empty += 1` -> 1, `.to_dict()`)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for a in args:
            if not a:
                continue
            for k, v in dict(a).items():
                self[k] = self._wrap(v)
        for k, v in kwargs.items():
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, cls):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return self[k]

    def __setattr__(self, k, v):
        self[k] = v

    def __missing__(self, k):
        child = _Dict()
        object.__setattr__(child, "_parent", (self, k))
        return child

    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        parent = self.__dict__.get("_parent")
        if parent is not None:
            p, pk = parent
            p[pk] = self
            object.__setattr__(self, "_parent", None)

    def __add__(self, other):
        if not self:
            return other
        return NotImplemented

    def to_dict(self):
        out = {}
        for k, v in self.items():
            if isinstance(v, _Dict):
                out[k] = v.to_dict()
            elif isinstance(v, (list, tuple)):
                out[k] = type(v)(x.to_dict() if isinstance(x, _Dict) else x for x in v)
            else:
                out[k] = v
        return out


def _imutils_resize(image, width=None, height=None, inter=None):
    """imutils==0.5.4 `resize` restated (aspect-preserving; `height` ignored when `width` given)."""
    import cv2

    if inter is None:
        inter = cv2.INTER_AREA
    (h, w) = image.shape[:2]
    if width is None and height is None:
        return image
    if width is None:
        r = height / float(h)
        dim = (int(w * r), height)
    else:
        r = width / float(w)
        dim = (width, int(h * r))
    return cv2.resize(image, dim, interpolation=inter)


def install() -> None:
    """Register the stand-in modules and make `import playaid` resolve to the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    import torch

    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    if "imutils" not in sys.modules:
        mod("imutils", resize=_imutils_resize)
    if "addict" not in sys.modules:
        mod("addict", Dict=_Dict)

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    class Accuracy(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            return torch.tensor(0.0)

    def _timm_create_model(name, num_classes=1000, pretrained=False, **kw):
        # reference resnet_transformer_detector.py:34 asks timm for "resnet50" with num_classes=0 (pooled 2048-d
        # features). timm's resnet50 is the torchvision v1.5 architecture with the same parameter names; weights
        # cannot be downloaded here, so the seeded default init stands in.
        import torchvision

        assert name == "resnet50", name
        net = torchvision.models.resnet50(weights=None)
        if num_classes == 0:
            net.fc = torch.nn.Identity()
        return net

    if "timm" not in sys.modules:
        mod("timm", create_model=_timm_create_model)
    if "pytorch_lightning" not in sys.modules:
        mod("pytorch_lightning", LightningModule=LightningModule)
    if "torchmetrics" not in sys.modules:
        mod("torchmetrics", Accuracy=Accuracy)

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return self

        def __getattr__(self, k):
            return _Anything()

    class _AnyModule(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return _Anything

    for name in ("albumentations", "dictdiffer"):
        if name not in sys.modules:
            sys.modules[name] = _AnyModule(name)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # resnet18(pretrained=True) would download: the reference architecture with random init instead.
    import torchvision.models as tvm

    _orig = tvm.resnet18

    def resnet18_noweights(*a, **k):
        k.pop("pretrained", None)
        k.pop("weights", None)
        return _orig(weights=None)

    tvm.resnet18 = resnet18_noweights
