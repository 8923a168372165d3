"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/ by running the REAL reference.

Run in the build container (needs /root/reference):  python -m oracle.gen_golden
Imports the reference under oracle/ref_shims.py and records, for seeded inputs that the tests can
rebuild without the reference:

  crops.npz     YoloCrop.square_crop (fighter.py:323-381): sha256 / sum / ok for ~260 boxes on three
                seeded 1080p frames (all resample regimes, clipped + letterboxed + 127-row cases),
                full crops for a handful
  bbox.npz      Fighter(data=record).crop.yolo_crop() (fighter.py:458-539) for 600 synthetic records
                + the fighter_test.py record (Appendix D1)
  windows.npz   action_sample_from_frame_middle_out (dataset_utils.py:109-138)
  timeline.json load_ground_truth_from_path (timeline.py:204-280) on tests/golden/sample_log.jsonl
  resformer.npz ResnetTransformerDetector forward (resnet_transformer_detector.py:99-143), seed 0, batch 2 and batch 1
  ai_timeline.json load_timeline_from_ai_output (timeline.py:52-105) on golden_ai_output(): digest + samples
  model.npz     CNNActionDetector(seed 0, default init).forward on a seeded input
                (models/cnn_action_detector.py:86-92) and the argmax / exp head (ai_runner.py:474-477)
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def golden_frames():
    """Three seeded 1080p test frames the tests can rebuild: noise, gradient (Appendix D3/D4), synthetic."""
    from workloads import synthetic

    noise = np.random.default_rng(0).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    y, x = np.mgrid[0:1080, 0:1920]
    grad = np.stack([x * 255 // 1919, y * 255 // 1079, (x + y) % 256], -1).astype(np.uint8)
    syn = synthetic.synth_frames([5], np.array([[[700, 500, 247, 283], [1200, 620, 180, 300]]]), device="cpu").numpy()[0]
    return [noise, grad, syn]


def golden_boxes():
    """(frame_id, box, padding) cases covering every regime of the chain."""
    rng = np.random.default_rng(42)
    cases = []
    d1 = (0.673046875, 0.5368055555555555, 0.12890625, 0.2625)
    for fid in (0, 1, 2):
        for pad in (30, 0):
            cases.append((fid, d1, pad))
    # 127-row quirk (D5), integer scales, same-size, upscales, odd sizes
    for sd in (196, 49, 98, 103, 107, 161, 187, 128, 256, 384, 512, 640, 127, 129, 64, 40, 33, 90, 255, 257, 300, 301, 700, 900):
        for pad in (0, 30):
            cases.append((1, (0.5, 0.5, sd / 1920 + 1e-9, 100 / 1080), pad))
            cases.append((0, (0.41, 0.52, 20 / 1920, sd / 1080 + 1e-9), pad))
    # clipped at every edge / corner, fully off-screen, degenerate
    for c in [(0.02, 0.5), (0.98, 0.5), (0.5, 0.03), (0.5, 0.97), (0.01, 0.02), (0.99, 0.98), (1.3, 0.5), (0.5, 1.4), (0.0, 0.0)]:
        for pad in (0, 30):
            cases.append((2, (c[0], c[1], 0.128, 0.2625), pad))
    cases.append((0, (0.5, 0.5, 0.0, 0.0), 0))
    # random in-frame boxes (cfg3 distribution)
    for _ in range(120):
        cases.append((int(rng.integers(0, 3)), (rng.uniform(0.05, 0.95), rng.uniform(0.05, 0.95), rng.uniform(0.02, 0.3),
                                               rng.uniform(0.03, 0.5)), int(rng.choice([0, 30]))))
    # large boxes
    for _ in range(8):
        cases.append((2, (rng.uniform(0.3, 0.7), rng.uniform(0.3, 0.7), rng.uniform(0.4, 0.9), rng.uniform(0.4, 0.9)), 30))
    return cases


def golden_ai_output(n_frames: int = 600, fighters=("Joker", "Pikachu")) -> dict:
    """Deterministic ai_output.yaml content (the schema AIRunner.write_output dumps, ai_runner.py:493-520)."""
    rng = np.random.default_rng(17)
    acts = ["Jab", "Wait", "Run", "Shield", "ForwardAir", "Damaged"]
    out = {}
    for name in fighters:
        per = {}
        for i in range(n_frames):
            cx, cy, w, h = rng.uniform(0.1, 0.9), rng.uniform(0.2, 0.8), rng.uniform(0.05, 0.3), rng.uniform(0.08, 0.4)
            per[i] = {"crop": f"0 {cx:.6f} {cy:.6f} {w:.6f} {h:.6f} 0", "action": acts[int(rng.integers(len(acts)))],
                      "predicted_action_confidence": float(np.round(rng.uniform(5, 100), 4))}
            if i % 7 == 0:
                per[i]["damage"] = float(np.round(rng.uniform(0, 150), 2))
        out[name] = per
    return out


def gen_ai_timeline():
    """load_timeline_from_ai_output (timeline.py:52-105) on golden_ai_output(): digest + sampled frames."""
    import hashlib
    import tempfile

    import yaml

    from oracle import ref_shims

    ref_shims.install()
    from playaid.timeline import load_timeline_from_ai_output

    path = os.path.join(tempfile.mkdtemp(), "ai_output.yaml")
    with open(path, "w") as f:
        yaml.dump(golden_ai_output(), f)
    tl = load_timeline_from_ai_output(path)
    blob = json.dumps(tl, sort_keys=True).encode()
    with open(os.path.join(GOLD, "ai_timeline.json"), "w") as f:
        json.dump({"n_frames": len(tl), "sha256": hashlib.sha256(blob).hexdigest(),
                   "samples": {str(i): tl[i] for i in (0, 1, 7, 299, 599)}}, f)
    print("ai timeline frames:", len(tl))


def gen_resformer():
    """ResnetTransformerDetector (models/resnet_transformer_detector.py:99-143), seed 0 default init, under the timm
    shim: log-probs for a batch of 2 and for its first window alone (the encoder attends across the batch)."""
    import torch

    from oracle import ref_shims

    ref_shims.install()
    from playaid.anim_ontology import MOVE_TO_CLASS_ID
    from playaid.models.resnet_transformer_detector import ResnetTransformerDetector

    torch.manual_seed(0)
    ref = ResnetTransformerDetector(actions=list(MOVE_TO_CLASS_ID.keys()), sequence_length=7).eval()
    x = torch.rand((2, 7, 3, 128, 128), generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        y2, y1 = ref(x), ref(x[:1])
    np.savez_compressed(os.path.join(GOLD, "resformer.npz"), logp_b2=y2.numpy(), logp_b1=y1.numpy(),
                        n_params=sum(p.numel() for p in ref.parameters()), keys=json.dumps(sorted(ref.state_dict().keys())),
                        freq_encoding=ref.model.freq_encoding.numpy(), torch_version=torch.__version__)
    print("resformer: logp range", float(y2.min()), float(y2.max()), "batch effect", float((y2[:1] - y1).abs().max()))


def main():
    from oracle import ref_shims

    ref_shims.install()
    import torch
    from playaid.dataset_utils import action_sample_from_frame_middle_out
    from playaid.fighter import Fighter, YoloCrop
    from playaid.timeline import load_ground_truth_from_path

    from workloads import synthetic

    os.makedirs(GOLD, exist_ok=True)
    import cv2, PIL, torchvision  # noqa: E401

    versions = dict(cv2=cv2.__version__, pillow=PIL.__version__, numpy=np.__version__, torch=torch.__version__,
                    torchvision=torchvision.__version__)

    # ---- crops
    frames = golden_frames()
    cases = golden_boxes()
    ok, sums, shas, full, full_idx, errs = [], [], [], [], [], []
    for i, (fid, box, pad) in enumerate(cases):
        try:
            res, crop = YoloCrop(*box).square_crop(frames[fid], 128, padding=pad)
            err = 0
        except ZeroDivisionError:
            res, crop, err = False, None, 1
        ok.append(bool(res)); errs.append(err)
        sums.append(int(crop.sum()) if res else -1)
        shas.append(hashlib.sha256(crop.tobytes()).hexdigest() if res else "")
        if res and (i < 8 or i % 29 == 0):
            full.append(crop); full_idx.append(i)
    np.savez_compressed(
        os.path.join(GOLD, "crops.npz"),
        frame_id=np.array([c[0] for c in cases]), box=np.array([c[1] for c in cases], dtype=np.float64),
        padding=np.array([c[2] for c in cases]), ok=np.array(ok), zero_div=np.array(errs), sum=np.array(sums),
        sha256=np.array(shas), full=np.array(full), full_idx=np.array(full_idx), versions=json.dumps(versions),
    )
    print("crops:", len(cases), "ok", sum(ok), "zero_div", sum(errs))

    # ---- bbox
    recs = [r for f in synthetic.synth_log_records(150, 2, seed=11) for r in f]
    recs += [r for f in synthetic.synth_log_records(150, 2, seed=12, stage_id=95, pos_x_range=(-80, 80), pos_y_range=(0, 60)) for r in f]
    d1 = {"camera_fov": 30.0, "camera_position": {"x": -0.00013416587898973376, "y": 14.01315975189209, "z": 167.240966796875},
          "camera_target_position": {"x": -0.0001499500940553844, "y": 11.852787017822266, "z": 0.0}, "damage": 0.0, "facing": -1.0,
          "fighter_id": 0, "motion_kind": 19292652517, "num_frames_left": 25200, "pos_x": 27.0, "pos_y": 0.1, "shield_size": 50.0,
          "status_kind": 0, "stock_count": 3, "attack_connected": False, "stage_id": 0, "fighter_name": 86, "hitstun_left": 0.0}
    recs.append(d1)
    boxes = np.array([Fighter(frame_num=0, data=r).crop.yolo_crop() for r in recs], dtype=np.float64)
    px = np.array([YoloCrop(*b).yolo_pixels(1920, 1080) for b in boxes])
    np.savez_compressed(os.path.join(GOLD, "bbox.npz"), records=json.dumps(recs), boxes=boxes, yolo_pixels=px)
    print("bbox:", boxes.shape, "D1", boxes[-1], px[-1])

    # ---- windows
    w_args = [(100, 7, 3, 1000, 1), (2, 7, 3, 10, 1), (0, 7, 3, 64, 0), (63, 7, 3, 64, 0), (30, 7, 3, 64, 0), (5, 5, 2, 40, 0),
              (10, 3, 1, 12, 1), (0, 7, 3, 5, 1)]
    w_out = [action_sample_from_frame_middle_out(a, b, c, d, min_frame=e) for (a, b, c, d, e) in w_args]
    with open(os.path.join(GOLD, "windows.json"), "w") as f:
        json.dump({"args": w_args, "out": w_out}, f)

    # ---- timeline: a log with an offset-able head and a dropped-frame gap
    log = synthetic.synth_log_records(12, 2, seed=3)
    for r in log[6]:
        r["fighter_id"] = 4 - r["fighter_id"] * 4  # ids 4, 0: exercises the re-id sort
    del log[8:10]  # frames 8, 9 missing -> gap of 3 in num_frames_left
    path = os.path.join(GOLD, "sample_log.jsonl")
    synthetic.write_log(path, log)
    tl = {}
    for off in (0, 2):
        gt = load_ground_truth_from_path(path, log_offset=off)
        tl[str(off)] = [[(r["num_frames_left"], r["fighter_id"], r["pos_x"]) for r in fr] for fr in gt]
    with open(os.path.join(GOLD, "timeline.json"), "w") as f:
        json.dump(tl, f)
    print("timeline frames:", {k: len(v) for k, v in tl.items()})

    # ---- model: reference class, seed 0, default init (Appendix D7)
    from playaid.anim_ontology import MOVE_TO_CLASS_ID
    from playaid.models.cnn_action_detector import CNNActionDetector

    torch.manual_seed(0)
    ref = CNNActionDetector(actions=list(MOVE_TO_CLASS_ID.keys()), sequence_length=7).eval()
    g = torch.Generator().manual_seed(7)
    x = torch.rand((3, 7, 3, 128, 128), generator=g)
    with torch.no_grad():
        lp = ref(x)
    pred = torch.argmax(lp, dim=1)
    conf = [float(torch.exp(lp)[i][int(pred[i])]) * 100.0 for i in range(lp.shape[0])]
    np.savez_compressed(os.path.join(GOLD, "model.npz"), logp=lp.numpy(), pred=pred.numpy(), conf=np.array(conf),
                        n_params=sum(p.numel() for p in ref.parameters()), actions=json.dumps(list(MOVE_TO_CLASS_ID.keys())),
                        keys=json.dumps(sorted(k for k in ref.state_dict().keys())), versions=json.dumps(versions))
    print("model: pred", pred.tolist(), "logp range", float(lp.min()), float(lp.max()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ai_timeline":   # regenerate only that fixture
        gen_ai_timeline()
    elif len(sys.argv) > 1 and sys.argv[1] == "resformer":
        gen_resformer()
    else:
        main()
        gen_ai_timeline()
        gen_resformer()
