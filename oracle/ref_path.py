"""TEST INFRASTRUCTURE ONLY -- CPU port of the reference hot path.

Restates, with the same third-party calls the reference makes (Pillow, OpenCV, torch /
torchvision on the CPU), every step between an ult_logger record + a decoded frame and a
per-frame action label. Each function cites the reference code it follows. It is the checker for
the CUDA path and the `cpu_baseline` / `--impl reference` leg of bench.py; the product package
never imports it.

Parity pin: tests/golden/*.npz hold outputs of the reference's own functions run under
oracle/ref_shims.py (oracle/gen_golden.py); tests/test_oracle_golden.py compares this port and
oracle/resample.c against them.
"""
from __future__ import annotations

import json
import math

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

# reference playaid/anim_ontology.py:497-570 ("fov" column) -- unknown stages fall back to 0
_STAGE_FOV = {0: 50, 3: 50, 44: 50, 51: 50, 86: 50, 89: 50, 95: 30, 107: 50, 118: 50, 242: 50, 257: 50, 268: 50,
              293: 50, 295: 50, 330: 50, 347: 50, 351: 50, 361: 50}


# ------------------------------------------------------------------ bbox geometry (fighter.py:31-155, 487-539)
def _lookat(cam, tgt):
    forward = np.array(cam, dtype=np.float64) - np.array(tgt, dtype=np.float64)
    forward /= np.linalg.norm(forward)
    up = np.array([0, 1, 0])
    right = np.cross(up, forward)
    right /= np.linalg.norm(right)
    up = np.cross(forward, right)
    m = np.eye(4)
    m[0, :3] = right
    m[1, :3] = up
    m[2, :3] = -forward
    m[:3, 3] = cam
    return m


def _project(point_world, K, pose, image_height=720):
    ph = np.append(point_world, 1)
    pc = np.linalg.inv(pose) @ ph
    pn = pc[:3] / pc[2]
    px = K @ pn
    px[1] = image_height - px[1]
    return np.round(px[:2]).astype(int)


def fighter_box(rec: dict):
    """Normalised (cx, cy, w, h) that `Fighter.set_from_json` stores in `fighter.crop`."""
    if "crop" in rec:
        _, cx, cy, w, h, _ = rec["crop"].split(" ")
        return (float(cx), float(cy), float(w), float(h))
    stage = rec["stage_id"] if rec["stage_id"] in _STAGE_FOV else 0
    fov = _STAGE_FOV[stage]
    f = 1280 / (2 * np.tan(np.deg2rad(fov) / 2))
    K = np.array([[f, 0, 1280 / 2], [0, f, 720 / 2], [0, 0, 1]])
    pose = _lookat(list(rec["camera_position"].values()), list(rec["camera_target_position"].values()))
    pos = [rec["pos_x"], rec["pos_y"], 0]
    pts = [_project(pos + np.array(o), K, pose) for o in ([-10, 20, 0], [10, 20, 0], [-10, -3, 0], [10, -3, 0])]
    xs = [int(p[0]) for p in pts]
    ys = [int(p[1]) for p in pts]
    cx = (xs[0] + xs[1] + xs[2] + xs[3]) / 4
    cy = (ys[0] + ys[1] + ys[2] + ys[3]) / 4
    w = max(xs) - min(xs)
    h = max(ys) - min(ys)
    return (cx / 1280, cy / 720, w / 1280, h / 720)


# ------------------------------------------------------------------ timeline (timeline.py:204-280)
def load_ground_truth(label_path, log_offset=0):
    gt, prev, index, skipped = [], -1, 0, 0
    with open(label_path) as f:
        for line in f:
            if skipped < 2 * log_offset:
                skipped += 1
                continue
            rec = json.loads(line)
            fn = index // 2
            if fn >= len(gt):
                gt.append([])
            diff = prev - rec["num_frames_left"]
            if prev > 0 and diff > 1:
                gt += [gt[-1]] * (diff - 1)
                index += (diff - 1) * 2
            gt[fn].append(rec)
            index += 1
            prev = rec["num_frames_left"]
    for i, fr in enumerate(gt):
        fr = sorted(fr, key=lambda r: r["fighter_id"])
        for j, r in enumerate(fr):
            r["fighter_id"] = j
        gt[i] = fr
    return gt


# ------------------------------------------------------------------ crop (fighter.py:305-381) with the library calls
def _imutils_resize_width(image, width):
    import cv2

    h, w = image.shape[:2]
    r = width / float(w)
    return cv2.resize(image, (width, int(h * r)), interpolation=cv2.INTER_AREA)


def square_crop_libs(image, box, output_size=128, padding=0):
    from PIL import Image, ImageOps

    H, W = image.shape[:2]
    cx, cy, cw, ch = int(box[0] * W), int(box[1] * H), int(box[2] * W), int(box[3] * H)
    sd = max(cw, ch)
    half = int(sd / 2)
    raw = image[max(cy - half - padding, 0) : min(cy + half + padding, H), max(cx - half - padding, 0) : min(cx + half + padding, W), :]
    if raw.shape[0] != sd or raw.shape[1] != sd:
        try:
            raw = np.array(ImageOps.pad(Image.fromarray(raw), (sd, sd), color="black"))
        except ValueError:
            return False, None
    if raw.shape[0] == 0 or raw.shape[1] == 0:
        return False, None
    crop = _imutils_resize_width(raw, output_size)
    if crop.shape[0] != output_size or crop.shape[1] != output_size:
        crop = np.array(ImageOps.pad(Image.fromarray(crop), (output_size, output_size), color="black"))
    if crop.shape != (output_size, output_size, 3):
        raise Exception(f"Bad output shape {crop.shape}")
    return True, crop


# ------------------------------------------------------------------ windows (dataset_utils.py:109-138)
def middle_out(middle_frame, n, delta, max_frames, min_frame=0, clamp=True):
    assert n % 2 == 1, "num_frames_per_sample must be odd"
    mid = math.floor(n / 2)
    out = []
    for i in range(n):
        off = abs(delta * ((mid - i) ** 2))
        if i < n / 2:
            v = middle_frame - off
            if clamp:
                v = max(min_frame, v)
        elif i == n / 2:
            v = middle_frame
        else:
            v = middle_frame + off
            if clamp:
                v = min(max_frames - 1, middle_frame + off)
        out.append(v)
    return out


# ------------------------------------------------------------------ model (models/cnn_action_detector.py:13-43,86-92)
class RefSpatialStreamCNN(nn.Module):
    def __init__(self, num_actions, sequence_length):
        super().__init__()
        from torchvision.models import resnet18

        self.cnn2d = resnet18(weights=None)
        self.cnn1d = nn.Sequential(nn.Conv1d(1000, 512, kernel_size=sequence_length, stride=1), nn.ReLU())
        self.classifier = nn.Sequential(nn.Linear(512, 128), nn.ReLU(), nn.Linear(128, num_actions))

    def features(self, crops):  # [n,3,H,W] -> [n,1000]
        return self.cnn2d(crops)

    def head_logits(self, feats):  # [B,S,1000] -> [B,A]
        x = feats.permute(0, 2, 1)
        x = self.cnn1d(x)
        x = x.view(x.size(0), -1)
        return self.classifier(x)

    def forward(self, x):
        B, S, C, H, W = x.size()
        f = self.cnn2d(x.view(B * S, C, H, W))
        return self.head_logits(f.view(B, S, -1))


class RefCNNActionDetector(nn.Module):
    def __init__(self, actions, sequence_length=4):
        super().__init__()
        self.actions = list(actions)
        self.num_actions = len(self.actions)
        self.sequence_length = sequence_length
        self.model = RefSpatialStreamCNN(self.num_actions, sequence_length)

    def forward(self, x):
        return F.log_softmax(self.model(x), dim=1)


def to_tensor(crops_rgb):
    """ai_runner.py:461-463: list of [128,128,3] u8 RGB -> [1,S,3,128,128] float / 255."""
    t = torch.tensor(np.array(crops_rgb))
    return t.permute(0, 3, 1, 2).unsqueeze(0).float() / 255.0


@torch.no_grad()
def classify_clip(frames_bgr, boxes, model: RefCNNActionDetector, padding=30, output_size=128, delta=3, min_frame=0,
                  as_shipped=False, batch=32, crops_out=None):
    """SURVEY 3.2 composition on the CPU. frames_bgr [N,H,W,3] u8, boxes [N,F,4] float64.

    as_shipped=True : one forward per window, each crop through ResNet-18 seven times, batch 1
                      (what ai_runner.py:493-520 does).
    as_shipped=False: identical results, but features are computed once per crop and reused.
    Returns label [N,F] int64 (-1 where the centre crop is invalid), logp [N,F,A], prob [N,F].
    """
    import cv2

    N, Fn = boxes.shape[:2]
    S = model.sequence_length
    A = model.num_actions
    rgb = np.zeros((N, Fn, output_size, output_size, 3), np.uint8)
    ok = np.zeros((N, Fn), bool)
    for i in range(N):
        for k in range(Fn):
            res, crop = square_crop_libs(frames_bgr[i], boxes[i, k], output_size, padding)
            ok[i, k] = res
            if res:
                rgb[i, k] = cv2.cvtColor(crop, cv2.COLOR_BGR2RGB)
    if crops_out is not None:
        crops_out["rgb"], crops_out["ok"] = rgb, ok
    label = np.full((N, Fn), -1, np.int64)
    logp = np.zeros((N, Fn, A), np.float32)
    prob = np.zeros((N, Fn), np.float32)
    model.eval()
    if as_shipped:
        for k in range(Fn):
            for i in range(min_frame, N):
                idx = middle_out(i, S, delta, N, min_frame)
                x = to_tensor([rgb[j, k] for j in idx])
                lp = model(x)
                p = int(torch.argmax(lp))
                label[i, k], logp[i, k], prob[i, k] = p, lp[0].numpy(), float(torch.exp(lp)[0][p])
        return label, logp, prob
    x = torch.from_numpy(rgb.reshape(N * Fn, output_size, output_size, 3)).permute(0, 3, 1, 2).float() / 255.0
    feats = torch.cat([model.model.features(x[s : s + batch]) for s in range(0, N * Fn, batch)]).view(N, Fn, -1)
    for k in range(Fn):
        idx = torch.tensor([middle_out(i, S, delta, N, min_frame) for i in range(min_frame, N)])
        lp = F.log_softmax(model.model.head_logits(feats[:, k][idx]), dim=1)
        p = torch.argmax(lp, dim=1)
        label[min_frame:, k] = p.numpy()
        logp[min_frame:, k] = lp.numpy()
        prob[min_frame:, k] = torch.exp(lp)[torch.arange(lp.shape[0]), p].numpy()
    return label, logp, prob
