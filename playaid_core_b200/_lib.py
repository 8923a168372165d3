"""ctypes binding of libplayaid_b200.so (include/playaid_b200.h). No fallback: a missing library or a
non-Blackwell device raises -- the product path never runs on the CPU."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PLAYAID_B200_LIB points at another build of the same library (A/B measurements); the default is the in-tree build
LIB_PATH = os.environ.get("PLAYAID_B200_LIB") or os.path.join(_HERE, "libplayaid_b200.so")

PA_OK = 0
CROP_OK, CROP_INVALID, CROP_ZERO_DIV, CROP_TOO_LARGE = 1, 0, -2, -7
DTYPE_U8, DTYPE_BF16, DTYPE_F32, DTYPE_BF16X2, DTYPE_F16, DTYPE_F16X2, DTYPE_BF16_U8, DTYPE_F16_U8 = 0, 1, 2, 3, 4, 5, 6, 7
LAYOUT_NHWC, LAYOUT_NCHW, LAYOUT_NHWC4, LAYOUT_NHWC4P = 0, 1, 2, 3
PREC_BF16, PREC_BF16X2, PREC_BF16X3, PREC_F16, PREC_F16X2, PREC_F16X3 = 0, 1, 2, 3, 4, 5
BOX_STRIDE = 8
LOG_STRIDE = 10

# every symbol include/playaid_b200.h declares
EXPORTS = [
    "pa_abi_version", "pa_status_string", "pa_last_error", "pa_ctx_create", "pa_ctx_destroy", "pa_preprocess", "pa_stage_windows",
    "pa_boxes_from_log",
    "pa_model_create", "pa_model_destroy", "pa_model_set_tensor", "pa_model_finalize", "pa_model_precision",
    "pa_model_workspace_bytes", "pa_features", "pa_features_u8", "pa_crop_elems", "pa_head", "pa_launch_count", "pa_conv2d", "pa_stem",
    "pa_profile_begin", "pa_profile_end", "pa_resformer_create", "pa_resformer_finalize",
    "pa_resformer_workspace_bytes", "pa_resformer_forward",
]


class PlayaidLibraryError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlayaidLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m playaid_core_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i32, i64, sz = c.c_void_p, c.c_int, c.c_int64, c.c_size_t
    lib.pa_abi_version.restype = i32
    lib.pa_status_string.restype = c.c_char_p
    lib.pa_status_string.argtypes = [i32]
    lib.pa_last_error.restype = c.c_char_p
    lib.pa_last_error.argtypes = [vp]
    lib.pa_ctx_create.argtypes = [i32, c.POINTER(vp)]
    lib.pa_ctx_destroy.argtypes = [vp]
    lib.pa_preprocess.argtypes = [vp, vp, i32, i32, i32, i64, i64, vp, i32, i32, i32, i32,
                                  c.POINTER(c.c_float), c.POINTER(c.c_float), vp, i32, i32, vp, vp]
    lib.pa_stage_windows.argtypes = [vp, vp, i32, i32, i32, i64, i64, vp, i32, i32, i32, vp, vp]
    lib.pa_boxes_from_log.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    lib.pa_model_create.argtypes = [vp, i32, i32, c.POINTER(vp)]
    lib.pa_resformer_create.argtypes = [vp, i32, i32, c.POINTER(vp)]
    lib.pa_resformer_finalize.argtypes = [vp, i32]
    lib.pa_resformer_workspace_bytes.argtypes = [vp, i32, c.POINTER(sz)]
    lib.pa_resformer_forward.argtypes = [vp, vp, i32, vp, vp, sz, vp]
    lib.pa_model_destroy.argtypes = [vp]
    lib.pa_model_set_tensor.argtypes = [vp, c.c_char_p, vp, c.POINTER(i64), i32]
    lib.pa_model_finalize.argtypes = [vp, i32]
    lib.pa_model_precision.argtypes = [vp]
    lib.pa_model_workspace_bytes.argtypes = [vp, i32, c.POINTER(sz)]
    lib.pa_features.argtypes = [vp, vp, i32, vp, vp, sz, vp]
    lib.pa_features_u8.argtypes = [vp, vp, i32, vp, vp, sz, vp]
    lib.pa_crop_elems.restype = sz
    lib.pa_crop_elems.argtypes = [i32]
    lib.pa_head.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp, vp, vp, sz, vp]
    fp = c.POINTER(c.c_float)
    lib.pa_conv2d.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp]
    lib.pa_stem.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, vp]
    lib.pa_profile_begin.argtypes = [vp]
    lib.pa_profile_end.argtypes = [vp, c.c_char_p, sz]
    lib.pa_launch_count.restype = i64
    lib.pa_launch_count.argtypes = [vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is c.c_int and name not in ("pa_abi_version",):
            fn.restype = i32
    _lib = lib
    return lib


def check(rc: int, ctx=None, what: str = "") -> None:
    if rc == PA_OK:
        return
    lib = load()
    msg = lib.pa_status_string(rc).decode()
    detail = lib.pa_last_error(ctx).decode() if ctx else ""
    raise PlayaidLibraryError(f"{what or 'playaid_b200'} failed: {msg} ({rc}) {detail}")


class Context:
    """One pa_ctx per (process, device)."""

    _by_device: dict = {}

    def __init__(self, device: int):
        import torch

        if not torch.cuda.is_available():
            raise PlayaidLibraryError("CUDA device required: playaid_core_b200 has no CPU path")
        self.lib = load()
        self.device = int(device)
        h = ctypes.c_void_p()
        check(self.lib.pa_ctx_create(self.device, ctypes.byref(h)), None, f"pa_ctx_create(device={device})")
        self.handle = h

    @classmethod
    def get(cls, device=None) -> "Context":
        import torch

        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        if isinstance(device, torch.device):
            device = device.index if device.index is not None else torch.cuda.current_device()
        if device not in cls._by_device:
            cls._by_device[device] = Context(device)
        return cls._by_device[device]

    def launch_count(self) -> int:
        return int(self.lib.pa_launch_count(self.handle))

    def profile_begin(self) -> None:
        check(self.lib.pa_profile_begin(self.handle), self.handle, "pa_profile_begin")

    def profile_end(self) -> dict:
        """{kernel name: (launches, total_ms)} in first-launch order."""
        buf = ctypes.create_string_buffer(1 << 16)
        check(self.lib.pa_profile_end(self.handle, buf, len(buf)), self.handle, "pa_profile_end")
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split("\t")
            out[name] = (int(n), float(ms))
        return out


def current_stream_ptr(device=None) -> ctypes.c_void_p:
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
