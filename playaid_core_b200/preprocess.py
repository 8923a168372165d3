"""Host front-end of the fused crop/resample/normalise kernel (`pa_preprocess`).

Replaces, for whole batches of (frame, fighter) boxes at once, the per-frame loop of reference
playaid/data_gen_scripts/gen_gt_action_detection.py:38-56 (`square_crop(frame, 128, padding=30)`)
followed by the to-tensor step of playaid/ult_action_dataset.py:302,349-359 /
playaid/ai_runner.py:448,461-463 (BGR->RGB, HWC->CHW, `.float()/255`).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .fighter import yolo_pixels_batch

_TORCH_DTYPE = {_lib.DTYPE_U8: torch.uint8, _lib.DTYPE_BF16: torch.bfloat16, _lib.DTYPE_F32: torch.float32,
                _lib.DTYPE_BF16X2: torch.bfloat16, _lib.DTYPE_F16: torch.float16, _lib.DTYPE_F16X2: torch.float16,
                _lib.DTYPE_BF16_U8: torch.bfloat16, _lib.DTYPE_F16_U8: torch.float16}


def crop_records(norm_boxes, frame_index, image_width: int, image_height: int) -> np.ndarray:
    """float64 [n,4] normalised (cx,cy,w,h) + int frame indices -> int32 [n,8] crop records
    `{frame, cx, cy, cw, ch, 0, 0, 0}` with `YoloCrop.yolo_pixels` truncation (fighter.py:305-314)."""
    px = yolo_pixels_batch(np.asarray(norm_boxes, dtype=np.float64).reshape(-1, 4), image_width, image_height)
    rec = np.zeros((px.shape[0], _lib.BOX_STRIDE), dtype=np.int32)
    rec[:, 0] = np.asarray(frame_index, dtype=np.int64).reshape(-1)
    rec[:, 1:5] = px
    return rec


def output_shape(n: int, out_size: int, dtype: int, layout: int):
    planes = 2 if dtype in (_lib.DTYPE_BF16X2, _lib.DTYPE_F16X2) else 1
    if layout == _lib.LAYOUT_NCHW:
        shp = (n, 3, out_size, out_size)
    elif layout == _lib.LAYOUT_NHWC4:
        shp = (n, out_size, out_size, 4)
    elif layout == _lib.LAYOUT_NHWC4P:
        shp = (n, out_size, out_size + 8, 4)
    else:
        shp = (n, out_size, out_size, 3)
    return (planes,) + shp if planes == 2 else shp


def preprocess_crops(
    frames: torch.Tensor,
    records: torch.Tensor,
    output_size: int = 128,
    padding: int = 30,
    swap_rb: bool = True,
    mean=(0.0, 0.0, 0.0),
    std=(1.0, 1.0, 1.0),
    dtype: int = _lib.DTYPE_F32,
    layout: int = _lib.LAYOUT_NCHW,
    out: torch.Tensor | None = None,
    status: torch.Tensor | None = None,
):
    """Launch the fused kernel. `frames` uint8 [N,H,W,3] (rows may be pitched) in device memory -- or in
    PINNED host memory, which the kernel then reads in place over PCIe (only the window bytes cross the
    bus; no staging copy of whole 1080p frames). `records` int32 CUDA [n,8].
    Returns (out, status); status[i] is 1 / 0 / -2 / -7 (see playaid_b200.h)."""
    if not records.is_cuda or not (frames.is_cuda or frames.is_pinned()):
        raise _lib.PlayaidLibraryError("preprocess_crops needs CUDA (or pinned host) frames and CUDA records (no CPU path)")
    if frames.dtype != torch.uint8 or frames.ndim != 4 or frames.shape[3] != 3 or frames.stride(3) != 1 or frames.stride(2) != 3:
        raise ValueError("frames must be uint8 [N,H,W,3] with packed pixels")
    if records.dtype != torch.int32 or records.ndim != 2 or records.shape[1] != _lib.BOX_STRIDE or not records.is_contiguous():
        raise ValueError("records must be contiguous int32 [n,8]")
    dev = records.device
    ctx = _lib.Context.get(dev)
    n = int(records.shape[0])
    N, H, W, _ = frames.shape
    if out is None:
        out = torch.empty(output_shape(n, output_size, dtype, layout), dtype=_TORCH_DTYPE[dtype], device=dev)
    if status is None:
        status = torch.empty((n,), dtype=torch.int32, device=dev)
    mean_a = (ctypes.c_float * 3)(*[float(v) for v in mean])
    std_a = (ctypes.c_float * 3)(*[float(v) for v in std])
    fstride = frames.stride(0) if N > 1 else frames.stride(1) * H
    with torch.cuda.device(dev):
        rc = ctx.lib.pa_preprocess(
            ctx.handle, frames.data_ptr(), N, H, W, frames.stride(1), fstride, records.data_ptr(), n,
            output_size, padding, 1 if swap_rb else 0, mean_a, std_a, out.data_ptr(), dtype, layout,
            status.data_ptr(), _lib.current_stream_ptr(dev),
        )
    _lib.check(rc, ctx.handle, "pa_preprocess")
    return out, status


def stage_windows(host_frames: torch.Tensor, records: torch.Tensor, dev_frames: torch.Tensor, padding: int = 30,
                  frame_base: int = 0) -> None:
    """`pa_stage_windows` on the current stream: copy the window rows of `records` (int32 CUDA [n,8], frame index
    = record.frame - frame_base) from pinned host `host_frames` into `dev_frames` (same shape / strides)."""
    if not host_frames.is_pinned() or not dev_frames.is_cuda or not records.is_cuda:
        raise _lib.PlayaidLibraryError("stage_windows needs pinned host frames, a CUDA frame buffer and CUDA records")
    if host_frames.dtype != torch.uint8 or host_frames.shape != dev_frames.shape or host_frames.stride() != dev_frames.stride():
        raise ValueError("host and device frame batches must have identical uint8 geometry")
    if host_frames.ndim != 4 or host_frames.shape[3] != 3 or host_frames.stride(3) != 1 or host_frames.stride(2) != 3:
        raise ValueError("frames must be uint8 [N,H,W,3] with packed pixels")
    dev = records.device
    ctx = _lib.Context.get(dev)
    N, H, W, _ = host_frames.shape
    fstride = host_frames.stride(0) if N > 1 else host_frames.stride(1) * H
    with torch.cuda.device(dev):
        rc = ctx.lib.pa_stage_windows(ctx.handle, host_frames.data_ptr(), N, H, W, host_frames.stride(1), fstride,
                                      records.data_ptr(), int(records.shape[0]), padding, frame_base,
                                      dev_frames.data_ptr(), _lib.current_stream_ptr(dev))
    _lib.check(rc, ctx.handle, "pa_stage_windows")


def square_crop_single(image, norm_box, output_size=128, padding=0):
    """`YoloCrop.square_crop` contract on one host image: (True, uint8 HWC crop) / (False, None)."""
    image = np.ascontiguousarray(image)
    if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
        raise ValueError("image must be uint8 [H,W,3]")
    H, W = image.shape[:2]
    dev = torch.device("cuda", torch.cuda.current_device())
    frames = torch.from_numpy(image).to(dev)[None]
    rec = torch.from_numpy(crop_records([norm_box], [0], W, H)).to(dev)
    out, status = preprocess_crops(frames, rec, output_size, padding, swap_rb=False, dtype=_lib.DTYPE_U8,
                                   layout=_lib.LAYOUT_NHWC)
    st = int(status.cpu()[0])
    if st == _lib.CROP_OK:
        return True, out[0].cpu().numpy()
    if st == _lib.CROP_ZERO_DIV:
        raise ZeroDivisionError("division by zero")
    if st == _lib.CROP_INVALID:
        return False, None
    raise _lib.PlayaidLibraryError(f"crop status {st}: window exceeds the kernel's staging limits")
