"""Seeded synthetic inputs: ult_logger JSONL logs and 1080p BGR gameplay frames.

Everything here is integer / float64 arithmetic with fixed seeds, so the CPU oracle and the GPU
path see byte-identical inputs. Frames are produced by the same torch code on any device
(integer ops only), in bounded sub-chunks.
"""
from __future__ import annotations

import json
from typing import Iterable

import numpy as np
import torch

# One known-good motion hash ("wait", reference playaid/fighter_test.py:26,48) -- the value only
# feeds Fighter.action_string, which is outside the hot path.
_MOTION_KIND_WAIT = 19292652517


def synth_log_records(
    n_frames: int,
    n_fighters: int = 2,
    seed: int = 2024,
    stage_id: int = 0,
    fighter_names: Iterable[int] = (86, 39, 8, 72),
    pos_x_range=(-45.0, 45.0),
    pos_y_range=(0.0, 25.0),
) -> list[list[dict]]:
    """`n_frames` x `n_fighters` ult_logger records with exactly the keys `Fighter.set_from_json`
    reads (reference playaid/fighter.py:461-554). Positions and camera follow a
    bounded random walk so consecutive boxes overlap like a real match."""
    rng = np.random.default_rng(seed)
    fighter_names = list(fighter_names)

    def walk(lo, hi, n, step):
        x = np.empty(n)
        x[0] = rng.uniform(lo, hi)
        d = rng.normal(0.0, step, n)
        for i in range(1, n):
            v = x[i - 1] + d[i]
            if v < lo:
                v = 2 * lo - v
            if v > hi:
                v = 2 * hi - v
            x[i] = min(max(v, lo), hi)
        return x

    cam_x = walk(-2.0, 2.0, n_frames, 0.05)
    cam_y = walk(10.0, 20.0, n_frames, 0.15)
    cam_z = walk(120.0, 220.0, n_frames, 1.0)
    tgt_x = cam_x + rng.normal(0, 1e-3, n_frames)
    tgt_y = cam_y - rng.uniform(1.0, 4.0)
    pos_x = [walk(pos_x_range[0], pos_x_range[1], n_frames, 0.9) for _ in range(n_fighters)]
    pos_y = [walk(pos_y_range[0], pos_y_range[1], n_frames, 0.7) for _ in range(n_fighters)]
    damage = [np.cumsum(rng.random(n_frames) < 0.01) * 3.5 for _ in range(n_fighters)]

    frames = []
    for i in range(n_frames):
        recs = []
        for k in range(n_fighters):
            recs.append(
                {
                    "camera_fov": 30.0,
                    "camera_position": {"x": float(cam_x[i]), "y": float(cam_y[i]), "z": float(cam_z[i])},
                    "camera_target_position": {"x": float(tgt_x[i]), "y": float(tgt_y[i]), "z": 0.0},
                    "damage": float(damage[k][i]),
                    "facing": 1.0 if (i // 90 + k) % 2 else -1.0,
                    "fighter_id": k,
                    "fighter_name": fighter_names[k % len(fighter_names)],
                    "motion_kind": _MOTION_KIND_WAIT,
                    "num_frames_left": 25200 - i,
                    "pos_x": float(pos_x[k][i]),
                    "pos_y": float(pos_y[k][i]),
                    "shield_size": 50.0,
                    "status_kind": 0,
                    "stock_count": 3,
                    "attack_connected": False,
                    "stage_id": stage_id,
                    "hitstun_left": 0.0,
                    "can_act": True,
                    "animation_frame_num": float(i % 40),
                }
            )
        frames.append(recs)
    return frames


def write_log(path: str, records: list[list[dict]]) -> None:
    """One JSON object per line, one line per fighter per frame (reference timeline.py:241-243)."""
    with open(path, "w") as f:
        for frame in records:
            for rec in frame:
                f.write(json.dumps(rec) + "\n")


def synth_free_boxes(n_frames: int, n_fighters: int = 4, seed: int = 7) -> np.ndarray:
    """cfg3 'variable bbox sizes': normalised (cx, cy, w, h) float64 [n_frames, n_fighters, 4] drawn
    directly (centres in-frame), as a slow random walk between keyframes."""
    rng = np.random.default_rng(seed)
    n_key = n_frames // 30 + 2
    key = np.empty((n_key, n_fighters, 4))
    key[..., 0] = rng.uniform(0.05, 0.95, (n_key, n_fighters))
    key[..., 1] = rng.uniform(0.05, 0.95, (n_key, n_fighters))
    key[..., 2] = rng.uniform(0.02, 0.30, (n_key, n_fighters))
    key[..., 3] = rng.uniform(0.03, 0.50, (n_key, n_fighters))
    t = np.arange(n_frames) / 30.0
    i0 = np.floor(t).astype(int)
    a = (t - i0)[:, None, None]
    return key[i0] * (1 - a) + key[i0 + 1] * a


def _hash32(x: torch.Tensor) -> torch.Tensor:
    """Integer avalanche on int64 tensors holding 32-bit values (identical on CPU and CUDA)."""
    m = 0xFFFFFFFF
    x = x & m
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & m
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & m
    return x ^ (x >> 16)


@torch.no_grad()
def synth_frames(
    frame_ids,
    boxes_px,
    H: int = 1080,
    W: int = 1920,
    device="cpu",
    seed: int = 1234,
    sub: int = 8,
    out: torch.Tensor | None = None,
) -> torch.Tensor:
    """uint8 BGR frames [n, H, W, 3]: drifting gradients + hash noise + one textured ellipse per
    fighter centred on its box. `boxes_px` int [n, F, 4] = (cx, cy, w, h) in frame pixels."""
    frame_ids = torch.as_tensor(np.asarray(frame_ids), dtype=torch.int64, device=device)
    boxes = torch.as_tensor(np.asarray(boxes_px), dtype=torch.int64, device=device)
    n = int(frame_ids.numel())
    F = int(boxes.shape[1]) if boxes.ndim == 3 else 0
    if out is None:
        out = torch.empty((n, H, W, 3), dtype=torch.uint8, device=device)
    x = torch.arange(W, device=device, dtype=torch.int64)[None, None, :]
    y = torch.arange(H, device=device, dtype=torch.int64)[None, :, None]
    for s in range(0, n, sub):
        f = frame_ids[s : s + sub][:, None, None]
        b = (x * 255 // max(W - 1, 1) + 3 * f) & 255
        g = (y * 255 // max(H - 1, 1) + 5 * f) & 255
        r = ((x + y) // 2 + 7 * f) & 255
        noise = _hash32(x * 73856093 + y * 19349663 + f * 83492791 + seed) & 31
        b = b + noise
        g = g + ((noise * 5) & 31)
        r = r + ((noise * 11) & 31)
        for k in range(F):
            bx = boxes[s : s + sub, k]
            cx, cy = bx[:, 0][:, None, None], bx[:, 1][:, None, None]
            w, h = bx[:, 2][:, None, None].clamp(min=2), bx[:, 3][:, None, None].clamp(min=2)
            dx, dy = x - cx, y - cy
            inside = (dx * dx * h * h + dy * dy * w * w) * 4 <= w * w * h * h
            # "action" id changes every 24 frames; stripes / checker texture depends on it
            act = _hash32((f // 24) * 977 + k * 131 + seed) & 63
            s1, s2 = (act & 7) + 1, ((act >> 3) & 7) + 1
            tex = ((dx * s1 + dy * s2 + 4 * f) >> 3) & 1
            blob = ((dx * dx + dy * dy) * 255 // (w * w + h * h).clamp(min=1)) & 255
            sb = (40 + 60 * k + 96 * tex + (act * 3)) & 255
            sg = (200 - 50 * k + blob) & 255
            sr = (90 + 37 * k + 128 * (1 - tex) + act) & 255
            b = torch.where(inside, sb + (noise >> 1), b)
            g = torch.where(inside, sg + (noise >> 1), g)
            r = torch.where(inside, sr + (noise >> 1), r)
        o = out[s : s + sub]
        o[..., 0] = b.clamp(0, 255).to(torch.uint8)
        o[..., 1] = g.clamp(0, 255).to(torch.uint8)
        o[..., 2] = r.clamp(0, 255).to(torch.uint8)
    return out
