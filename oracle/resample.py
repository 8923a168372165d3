"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/resample.c (see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "resample.c")
_SO = os.path.join(_HERE, "liboracle.so")

OK, INVALID, ERR_ZERO_DIV = 1, 0, -2


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off (OpenCV's area path must not be FMA-contracted)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"]
        )
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        u8p = ctypes.c_void_p
        i = ctypes.c_int
        for name in ("pa_oracle_pil_bicubic", "pa_oracle_pil_pad", "pa_oracle_cv_area"):
            fn = getattr(_lib, name)
            fn.argtypes = [u8p, i, i, i, u8p, i, i]
            fn.restype = i
        _lib.pa_oracle_square_crop.argtypes = [u8p, i, i, i, i, i, i, i, i, i, u8p]
        _lib.pa_oracle_square_crop.restype = i
    return _lib


def _call3(name, img, oh, ow):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    out = np.empty((max(oh, 0), max(ow, 0), 3), np.uint8)
    rc = getattr(lib(), name)(img.ctypes.data, h, w, w * 3, out.ctypes.data, oh, ow)
    return rc, out


def pil_bicubic(img, ow, oh):
    """Image.fromarray(img).resize((ow, oh), BICUBIC)"""
    rc, out = _call3("pa_oracle_pil_bicubic", img, oh, ow)
    if rc != OK:
        raise ValueError(f"pil_bicubic rc={rc}")
    return out


def pil_pad(img, sw, sh):
    """ImageOps.pad(Image.fromarray(img), (sw, sh), color='black')"""
    rc, out = _call3("pa_oracle_pil_pad", img, sh, sw)
    if rc == ERR_ZERO_DIV:
        raise ZeroDivisionError("division by zero")
    if rc != OK:
        raise ValueError(f"pil_pad rc={rc}")
    return out


def cv_area(img, ow, oh):
    """cv2.resize(img, (ow, oh), interpolation=cv2.INTER_AREA)"""
    rc, out = _call3("pa_oracle_cv_area", img, oh, ow)
    if rc != OK:
        raise ValueError(f"cv_area rc={rc}")
    return out


def yolo_pixels(crop, W, H):
    """YoloCrop.yolo_pixels (fighter.py:305-314): int() truncation of float64 products."""
    cx, cy, cw, ch = crop
    return int(cx * W), int(cy * H), int(cw * W), int(ch * H)


def square_crop(image, crop, output_size=128, padding=0):
    """YoloCrop.square_crop (fighter.py:323-381) -> (ok, crop[out,out,3] u8 | None).

    `crop` is the normalised (cx, cy, w, h) tuple. Raises ZeroDivisionError where the reference does.
    """
    image = np.ascontiguousarray(image, dtype=np.uint8)
    H, W = image.shape[:2]
    cx, cy, cw, ch = yolo_pixels(crop, W, H)
    out = np.empty((output_size, output_size, 3), np.uint8)
    rc = lib().pa_oracle_square_crop(image.ctypes.data, H, W, image.strides[0], cx, cy, cw, ch, output_size, padding, out.ctypes.data)
    if rc == OK:
        return True, out
    if rc == ERR_ZERO_DIV:
        raise ZeroDivisionError("division by zero")
    if rc == INVALID:
        return False, None
    raise RuntimeError(f"oracle square_crop rc={rc}")
