"""Fighter bounding boxes and the square-crop entry point (host side of the hot path).

Mirrors, for this path only, reference playaid/fighter.py:

* camera geometry `calculate_focal_length` :31-48, `calculate_intrinsic_matrix` :66-84,
  `calculate_lookat_matrix` :87-120, `project_point_to_pixel` :123-155 and the bbox part of
  `Fighter.set_from_json` :487-539 -- restated as one vectorised float64 routine
  (`boxes_from_records`) so a whole match is projected in a few numpy calls;
* `YoloCrop` :158-392 -- same constructor / `from_pixel_coordinates` / `from_string` /
  `yolo_crop` / `yolo_pixels` / `xyxy_pixels` / `square_crop` / `__str__` surface. `square_crop`
  keeps the reference contract `(True, uint8[out,out,3])` / `(False, None)` but runs the fused
  CUDA resample kernel through the C-ABI library; there is no CPU path.

The rest of `Fighter` (damage / hit-stun bookkeeping) is analytics state outside the hot path.
"""
from __future__ import annotations

import numpy as np

from .anim_ontology import stage_fov

VIRTUAL_W, VIRTUAL_H = 1280, 720  # the projection always targets a virtual 1280x720 image (:497)
# world-space corner offsets around the fighter position (:507-526)
_CORNERS = np.array([[-10.0, 20.0, 0.0], [10.0, 20.0, 0.0], [-10.0, -3.0, 0.0], [10.0, -3.0, 0.0]])


def calculate_focal_length(fov, image_width):
    return image_width / (2 * np.tan(np.deg2rad(fov) / 2))


def calculate_intrinsic_matrix(fov, image_width, image_height):
    f = calculate_focal_length(fov, image_width)
    return np.array([[f, 0, image_width / 2], [0, f, image_height / 2], [0, 0, 1]])


def _lookat_batch(cam: np.ndarray, tgt: np.ndarray) -> np.ndarray:
    """[n,3],[n,3] -> [n,4,4] camera poses: rows right / up / -forward, last column = position."""
    fwd = cam - tgt
    fwd = fwd / np.sqrt((fwd * fwd).sum(-1, keepdims=True))
    up = np.broadcast_to(np.array([0.0, 1.0, 0.0]), fwd.shape)
    right = np.cross(up, fwd)
    right = right / np.sqrt((right * right).sum(-1, keepdims=True))
    up2 = np.cross(fwd, right)
    M = np.zeros((cam.shape[0], 4, 4))
    M[:, 3, 3] = 1.0
    M[:, 0, :3] = right
    M[:, 1, :3] = up2
    M[:, 2, :3] = -fwd
    M[:, :3, 3] = cam
    return M


def project_points_to_pixels(points, fov, cam, tgt, image_height=VIRTUAL_H) -> np.ndarray:
    """points [n,p,3], fov [n], cam/tgt [n,3] -> int64 [n,p,2] pixel coordinates in the virtual
    1280x720 image (y flipped, round-half-even like `np.round`)."""
    points = np.asarray(points, dtype=np.float64)
    n, p, _ = points.shape
    Minv = np.linalg.inv(_lookat_batch(np.asarray(cam, np.float64), np.asarray(tgt, np.float64)))
    ph = np.concatenate([points, np.ones((n, p, 1))], axis=-1)
    pc = np.einsum("nij,npj->npi", Minv, ph)
    nrm = pc[..., :3] / pc[..., 2:3]
    f = VIRTUAL_W / (2 * np.tan(np.deg2rad(np.asarray(fov, np.float64)) / 2))
    K = np.zeros((n, 3, 3))
    K[:, 0, 0] = f
    K[:, 1, 1] = f
    K[:, 0, 2] = VIRTUAL_W / 2
    K[:, 1, 2] = VIRTUAL_H / 2
    K[:, 2, 2] = 1.0
    px = np.einsum("nij,npj->npi", K, nrm)
    px[..., 1] = image_height - px[..., 1]
    return np.round(px[..., :2]).astype(np.int64)


def boxes_from_records(records) -> np.ndarray:
    """Flat list of ult_logger dicts -> float64 [n,4] normalised (cx, cy, w, h), the value
    `Fighter.set_from_json` stores in `fighter.crop` (reference fighter.py:487-539). A record
    carrying an AI `"crop"` string overrides the projection (:503-504)."""
    n = len(records)
    out = np.empty((n, 4))
    if n == 0:
        return out
    pos = np.array([[r["pos_x"], r["pos_y"], 0.0] for r in records], dtype=np.float64)
    cam = np.array([list(r["camera_position"].values()) for r in records], dtype=np.float64)
    tgt = np.array([list(r["camera_target_position"].values()) for r in records], dtype=np.float64)
    fov = np.array([stage_fov(r["stage_id"]) for r in records], dtype=np.float64)
    px = project_points_to_pixels(pos[:, None, :] + _CORNERS[None], fov, cam, tgt)
    xs, ys = px[..., 0], px[..., 1]
    out[:, 0] = xs.sum(-1) / 4 / VIRTUAL_W
    out[:, 1] = ys.sum(-1) / 4 / VIRTUAL_H
    out[:, 2] = (xs.max(-1) - xs.min(-1)) / VIRTUAL_W
    out[:, 3] = (ys.max(-1) - ys.min(-1)) / VIRTUAL_H
    for i, r in enumerate(records):
        if "crop" in r:
            c = YoloCrop.from_string(r["crop"])
            out[i] = c.yolo_crop()
    return out


def log_record_array(records, frame_index=None) -> np.ndarray:
    """Flat list of ult_logger dicts -> float64 [n, 10] rows {pos_x, pos_y, camera xyz, target xyz, focal length, frame}:
    the input of `boxes_from_records_device` / `pa_boxes_from_log`. The focal length of the record's stage is taken here
    (calculate_focal_length, reference fighter.py:31-48) so that the device uses the very double numpy computes."""
    n = len(records)
    a = np.empty((n, 10), dtype=np.float64)
    if n == 0:
        return a
    a[:, 0] = [r["pos_x"] for r in records]
    a[:, 1] = [r["pos_y"] for r in records]
    a[:, 2:5] = [list(r["camera_position"].values()) for r in records]
    a[:, 5:8] = [list(r["camera_target_position"].values()) for r in records]
    fov = np.array([stage_fov(r["stage_id"]) for r in records], dtype=np.float64)
    a[:, 8] = VIRTUAL_W / (2 * np.tan(np.deg2rad(fov) / 2))
    a[:, 9] = np.arange(n) if frame_index is None else np.asarray(frame_index, dtype=np.float64)
    return a


def boxes_from_records_device(log_array, image_width: int, image_height: int, device=None):
    """Device form of `boxes_from_records` + `crop_records`: float64 [n,10] rows from `log_record_array`
    -> (boxes float64 CUDA [n,4], crop records int32 CUDA [n,8]); `pa_boxes_from_log`, one thread per record. Records that
    carry an AI "crop" override are not handled here (use `boxes_from_records`)."""
    import ctypes

    import torch

    from . import _lib

    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ctx = _lib.Context.get(dev)
    rec = torch.as_tensor(np.ascontiguousarray(log_array, dtype=np.float64)).to(dev)
    n = int(rec.shape[0])
    boxes = torch.empty((n, 4), dtype=torch.float64, device=dev)
    crops = torch.empty((n, _lib.BOX_STRIDE), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = ctx.lib.pa_boxes_from_log(ctx.handle, rec.data_ptr(), n, int(image_width), int(image_height), boxes.data_ptr(),
                                       crops.data_ptr(), _lib.current_stream_ptr(dev))
    _lib.check(rc, ctx.handle, "pa_boxes_from_log")
    return boxes, crops


def boxes_from_timeline(timeline) -> np.ndarray:
    """`load_ground_truth_from_path` output -> float64 [n_frames, n_fighters, 4]; fighters in
    `fighter_id` order like `update_fighters_from_timeline` (reference timeline.py:186-201)."""
    n_f = len(timeline[0]) if timeline else 0
    flat = [r for frame in timeline for r in sorted(frame, key=lambda x: x["fighter_id"])]
    return boxes_from_records(flat).reshape(len(timeline), n_f, 4)


def yolo_pixels_batch(boxes, image_width, image_height) -> np.ndarray:
    """float64 [...,4] normalised -> int32 [...,4] (cx, cy, cw, ch): `int()` truncation toward zero
    of the float64 products, as `YoloCrop.yolo_pixels` (reference fighter.py:305-314)."""
    b = np.asarray(boxes, dtype=np.float64)
    scale = np.array([image_width, image_height, image_width, image_height], dtype=np.float64)
    return np.trunc(b * scale).astype(np.int32)


class YoloCrop:
    """Normalised YOLO box; `square_crop` is the drop-in for reference fighter.py:323-381."""

    def __init__(self, center_x, center_y, crop_width, crop_height, confidence=0, class_id=-1):
        self.center_x = center_x
        self.center_y = center_y
        self.crop_width = crop_width
        self.crop_height = crop_height
        self.confidence = confidence
        self.class_id = class_id

    @classmethod
    def from_pixel_coordinates(cls, image_width, image_height, x1, y1, x2, y2, x3, y3, x4, y4):
        cx = (x1 + x2 + x3 + x4) / 4
        cy = (y1 + y2 + y3 + y4) / 4
        w = max(x1, x2, x3, x4) - min(x1, x2, x3, x4)
        h = max(y1, y2, y3, y4) - min(y1, y2, y3, y4)
        return cls(cx / image_width, cy / image_height, w / image_width, h / image_height)

    @classmethod
    def from_pixel_yolo(cls, image_width, image_height, center_x, center_y, width, height):
        return cls(center_x / image_width, center_y / image_height, width / image_width, height / image_height)

    @classmethod
    def from_string(cls, yolo_string):
        class_id, cx, cy, w, h, conf = yolo_string.split(" ")
        return cls(float(cx), float(cy), float(w), float(h), confidence=float(conf), class_id=int(class_id))

    def yolo_crop(self):
        return (self.center_x, self.center_y, self.crop_width, self.crop_height)

    def xyxy_norm(self):
        return (
            self.center_x - (self.crop_width / 2),
            self.center_y - (self.crop_height / 2),
            self.center_x + (self.crop_width / 2),
            self.center_y + (self.crop_height / 2),
        )

    def xyxy_pixels(self, image_width, image_height):
        x1, y1, x2, y2 = self.xyxy_norm()
        return (
            max(0, int(x1 * image_width)),
            max(0, int(y1 * image_height)),
            min(image_width, int(x2 * image_width)),
            min(image_height, int(y2 * image_height)),
        )

    def center_pixels(self, image_width, image_height):
        return (int(self.center_x * image_width), int(self.center_y * image_height))

    def yolo_pixels(self, image_width, image_height):
        return (
            int(self.center_x * image_width),
            int(self.center_y * image_height),
            int(self.crop_width * image_width),
            int(self.crop_height * image_height),
        )

    def square_crop(self, image, output_size=128, padding=0):
        """(True, uint8 [output_size, output_size, 3]) or (False, None); channel order preserved.

        Same contract as the reference, computed by the fused CUDA kernel (`pa_preprocess`).
        Raises ZeroDivisionError where the reference does (zero-sized box with a non-empty window).
        """
        from .preprocess import square_crop_single

        return square_crop_single(image, self.yolo_crop(), output_size, padding)

    def __str__(self):
        return (
            f"{self.class_id} {self.center_x} {self.center_y} {self.crop_width} "
            + f"{self.crop_height} {self.confidence}"
        )

    def __repr__(self):
        return str(self)
