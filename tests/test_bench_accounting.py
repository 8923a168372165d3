"""bench.py's byte accounting for the end-to-end line (host logic, no GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _mask_bytes(boxes_px, H, W, pad):
    """brute force: bytes of the union of the two windows of every frame"""
    total = 0
    for f in range(boxes_px.shape[0]):
        m = np.zeros((H, W), bool)
        for cx, cy, cw, ch in boxes_px[f]:
            sd = max(int(cw), int(ch)); half = sd // 2
            y0, y1 = max(int(cy) - half - pad, 0), min(int(cy) + half + pad, H)
            x0, x1 = max(int(cx) - half - pad, 0), min(int(cx) + half + pad, W)
            m[y0:y1, x0:x1] = True
        total += 3 * int(m.sum())
    return total


def test_staged_window_bytes_is_the_union_for_two_fighters():
    import bench

    rng = np.random.default_rng(5)
    n = 40
    px = np.stack([rng.integers(0, bench.W, (n, 2)), rng.integers(0, bench.H, (n, 2)), rng.integers(20, 500, (n, 2)),
                   rng.integers(20, 500, (n, 2))], axis=-1)
    px[:8, 1] = px[:8, 0] + rng.integers(-60, 60, (8, 4))          # heavily overlapping pairs
    px[8:10, 1] = px[8:10, 0]                                        # identical windows
    px = np.abs(px)
    assert bench.staged_window_bytes(px) == _mask_bytes(px, bench.H, bench.W, 30)
    assert bench.staged_window_bytes(px) <= bench.window_bytes(px).sum()
