"""TEST INFRASTRUCTURE ONLY -- golden labels for a slice of BASELINE cfg4 (7-minute match, fp32 vs 16-bit parity through
the reference's timeline + Stats consumers; SURVEY 8d parity gate).

Run in the build container (needs /root/reference):  python -m oracle.gen_cfg4_golden

The first SLICE frames of the cfg4 synthetic match (log seed 4242, frame seed 99, 2 fighters) are classified by the fp32
CPU oracle (oracle/ref_path.py, pinned by tests/golden/model.npz + crops.npz); the labels go through
`ActionDetector.ai_output`-shaped yaml into the REFERENCE's `load_timeline_from_ai_output` (its hard-coded first 600
frames) and, for the whole slice, into our kwargs-lifted loader, then through the reference's
`update_fighters_from_timeline` + `Stats.record_frame`; the digests of the resulting `Stats.stats` dicts are recorded.
tests/test_gpu_model.py::test_cfg4_slice_labels_and_stats asserts that the GPU label stream is IDENTICAL to these labels
(so its Stats are, `Stats` being a function of the labels and boxes); tests/test_oracle_golden.py re-derives the digests
from the committed labels whenever /root/reference is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
SLICE = 2048
REACH = 27
LOG_SEED, FRAME_SEED = 4242, 99
NAMES = ["Joker", "Pikachu"]   # the pair the reference loader is hard-wired to (timeline.py:57-62)


def slice_boxes(n):
    from playaid_core_b200.fighter import boxes_from_records
    from workloads import synthetic

    recs = synthetic.synth_log_records(n, 2, seed=LOG_SEED)
    return boxes_from_records([r for f in recs for r in f]).reshape(n, 2, 4)


def ai_output(label, prob, boxes):
    from playaid_core_b200.anim_ontology import ACTIONS
    from playaid_core_b200.fighter import YoloCrop

    out = {}
    for k, name in enumerate(NAMES):
        out[name] = {i: {"crop": str(YoloCrop(*[float(v) for v in boxes[i, k]])), "action": ACTIONS[int(label[i, k])],
                         "predicted_action_confidence": float(prob[i, k]) * 100.0} for i in range(label.shape[0])}
    return out


def canon(d):
    if isinstance(d, dict):
        return [[repr(k), canon(v)] for k, v in sorted(d.items(), key=lambda kv: repr(kv[0]))]   # keys mix str and int
    if isinstance(d, (list, tuple)):
        return [canon(v) for v in d]
    return d if isinstance(d, (int, float, str, bool, type(None))) else str(d)


def digest(d):
    return hashlib.sha256(json.dumps(canon(d)).encode()).hexdigest()


def reference_stats_digests(label, prob, boxes):
    """(sha of Stats.stats over the reference loader's 600 frames, sha over the whole slice). Needs /root/reference."""
    import yaml

    from oracle import ref_shims

    ref_shims.install()
    import playaid.constants as constants

    constants.AI_CACHE = tempfile.mkdtemp()     # Stats.__init__ makes directories under it (stats.py:65-67)
    from playaid.stats import Stats
    from playaid.timeline import load_timeline_from_ai_output as ref_loader, update_fighters_from_timeline

    from playaid_core_b200.timeline import load_timeline_from_ai_output as our_loader

    def run_stats(timeline):
        stats = Stats("/tmp/cfg4/match.mp4")
        fighters = []
        for i, frame in enumerate(timeline):
            fighters = update_fighters_from_timeline(i, frame, fighters)
            stats.record_frame(fighters)
        return stats.stats.to_dict()

    path = os.path.join(tempfile.mkdtemp(), "ai_output.yaml")
    with open(path, "w") as f:
        yaml.dump(ai_output(label, prob, boxes), f)
    t_ref = ref_loader(path)
    assert t_ref == our_loader(path), "our loader differs from the reference's on its own range"
    full = our_loader(path, max_frames=None, fighters=NAMES, fighter_to_player_id={"Pikachu": 0, "Joker": 1})
    assert len(full) == label.shape[0]
    return digest(run_stats(t_ref)), digest(run_stats(full))


def main():
    import torch

    from oracle import ref_path
    from playaid_core_b200.anim_ontology import ACTIONS
    from playaid_core_b200.fighter import yolo_pixels_batch
    from workloads import synthetic, weights

    torch.set_num_threads(os.cpu_count() or 8)
    n = SLICE + REACH
    boxes = slice_boxes(n)
    model = ref_path.RefCNNActionDetector(ACTIONS, 7).eval()
    model.load_state_dict(weights.calibrated_state_dict(0))
    px = yolo_pixels_batch(boxes, 1920, 1080)
    # crops chunk by chunk (a 2 075-frame 1080p clip is 12.9 GB), then one pass of the model over all crops
    rgb = np.zeros((n, 2, 128, 128, 3), np.uint8)
    import cv2

    for s in range(0, n, 64):
        e = min(n, s + 64)
        frames = synthetic.synth_frames(np.arange(s, e), px[s:e], device="cpu", seed=FRAME_SEED).numpy()
        for i in range(s, e):
            for k in range(2):
                ok, crop = ref_path.square_crop_libs(frames[i - s], boxes[i, k], 128, 30)
                assert ok
                rgb[i, k] = cv2.cvtColor(crop, cv2.COLOR_BGR2RGB)
    with torch.no_grad():
        x = torch.from_numpy(rgb.reshape(n * 2, 128, 128, 3)).permute(0, 3, 1, 2).float() / 255.0
        feats = torch.cat([model.model.features(x[s : s + 32]) for s in range(0, n * 2, 32)]).view(n, 2, -1)
        label = np.zeros((n, 2), np.int64); prob = np.zeros((n, 2), np.float32); logp = np.zeros((n, 2, len(ACTIONS)), np.float32)
        for k in range(2):
            idx = torch.tensor([ref_path.middle_out(i, 7, 3, n, 0) for i in range(n)])
            lp = torch.log_softmax(model.model.head_logits(feats[:, k][idx]), dim=1)
            p = torch.argmax(lp, dim=1)
            label[:, k] = p.numpy(); logp[:, k] = lp.numpy()
            prob[:, k] = torch.exp(lp)[torch.arange(n), p].numpy()
    label, prob, logp = label[:SLICE], prob[:SLICE], logp[:SLICE]
    srt = np.sort(logp, -1)
    margin = (srt[..., -1] - srt[..., -2]).astype(np.float32)
    sha600, sha_full = reference_stats_digests(label, prob, boxes[:SLICE])
    np.savez_compressed(os.path.join(GOLD, "cfg4_slice.npz"), label=label.astype(np.int8), prob=prob, margin=margin,
                        logp_max_abs=np.abs(logp).max(-1).astype(np.float32), logp_top=srt[..., -1].astype(np.float32),
                        stats_sha256_first600=sha600, stats_sha256_slice=sha_full, slice=SLICE, log_seed=LOG_SEED, frame_seed=FRAME_SEED)
    print(f"cfg4 slice: {SLICE} frames, {len(np.unique(label))} distinct labels, margin min {margin.min():.4f} median {np.median(margin):.3f}; "
          f"stats sha (first 600 / slice) {sha600[:16]} / {sha_full[:16]}")


if __name__ == "__main__":
    main()
