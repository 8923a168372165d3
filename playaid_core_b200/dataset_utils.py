"""Temporal-window sampling for the action classifier (host side).

Mirrors reference playaid/dataset_utils.py:109-138 (`action_sample_from_frame_middle_out`): same
name, arguments, AssertionError on an even window, and the same quadratic middle-out spacing
`offset = |delta * (mid - i)^2|` with clamping to `[min_frame, max_frames - 1]`.

`window_index_table` is the batched form the GPU head consumes: one int32 row of window indices
per frame, identical to calling the scalar function for every frame.
"""
from __future__ import annotations

import math

import numpy as np


def action_sample_from_frame_middle_out(
    middle_frame, num_frames_per_sample, frame_delta, max_frames, min_frame=0, clamp=True
):
    assert num_frames_per_sample % 2 == 1, "num_frames_per_sample must be odd"
    mid = math.floor(num_frames_per_sample / 2)
    out = []
    for i in range(num_frames_per_sample):
        off = abs(frame_delta * ((mid - i) ** 2))
        if i < num_frames_per_sample / 2:
            # includes the middle slot itself (offset 0): the reference clamps it to min_frame too
            v = middle_frame - off
            if clamp:
                v = max(min_frame, v)
        else:
            v = middle_frame + off
            if clamp:
                v = min(max_frames - 1, v)
        out.append(v)
    return out


def window_index_table(
    frames, num_frames_per_sample=7, frame_delta=3, max_frames=None, min_frame=0, clamp=True
) -> np.ndarray:
    """int32 [len(frames), num_frames_per_sample]; row r equals
    `action_sample_from_frame_middle_out(frames[r], ...)`. `frames` may be an int (-> range)."""
    assert num_frames_per_sample % 2 == 1, "num_frames_per_sample must be odd"
    if np.isscalar(frames):
        frames = np.arange(min_frame, int(frames))
    frames = np.asarray(frames, dtype=np.int64)
    if max_frames is None:
        max_frames = int(frames.max()) + 1 if frames.size else 0
    mid = num_frames_per_sample // 2
    i = np.arange(num_frames_per_sample)
    off = np.abs(frame_delta * (mid - i) ** 2)
    sign = np.sign(i - mid)
    idx = frames[:, None] + (sign * off)[None, :]
    if clamp:
        left = np.maximum(idx, min_frame)
        right = np.minimum(idx, max_frames - 1)
        idx = np.where(i[None, :] <= mid, left, right)
    return idx.astype(np.int32)
