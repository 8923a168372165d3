"""CPU tests of the product's host-side mirrors (no GPU): bbox geometry, window tables, timeline
loader, crop records, checkpoint plumbing -- against the reference's golden outputs."""
import json
import os

import numpy as np
import pytest

from playaid_core_b200 import dataset_utils, fighter, timeline
from playaid_core_b200.anim_ontology import ACTIONS, MOVE_TO_CLASS_ID, stage_fov


def test_action_classes(golden_dir):
    g = np.load(os.path.join(golden_dir, "model.npz"))
    assert ACTIONS == json.loads(str(g["actions"]))  # list(MOVE_TO_CLASS_ID.keys()) of the reference
    assert len(ACTIONS) == 63 and MOVE_TO_CLASS_ID["Jab"] == 0 and MOVE_TO_CLASS_ID["Undefined"] == 61 and MOVE_TO_CLASS_ID["Grabbed"] == 62
    assert stage_fov(95) == 30 and stage_fov(0) == 50 and stage_fov(12345) == 50


def test_boxes_from_records_bit_equal(golden_dir):
    g = np.load(os.path.join(golden_dir, "bbox.npz"))
    recs = json.loads(str(g["records"]))
    got = fighter.boxes_from_records(recs)
    assert np.array_equal(got, g["boxes"])
    assert np.array_equal(fighter.yolo_pixels_batch(got, 1920, 1080), g["yolo_pixels"])
    c = fighter.YoloCrop(*got[-1])
    assert c.yolo_pixels(1920, 1080) == (1292, 579, 247, 283) and c.xyxy_pixels(1920, 1080) == (1168, 438, 1416, 721)  # D1
    recs[0]["crop"] = "2 0.25 0.5 0.1 0.2 0.9"  # AI override (fighter.py:503-504)
    assert tuple(fighter.boxes_from_records(recs[:1])[0]) == (0.25, 0.5, 0.1, 0.2)


def test_yolocrop_surface():
    c = fighter.YoloCrop.from_string("3 0.5 0.25 0.125 0.0625 0.75")
    assert (c.class_id, c.confidence) == (3, 0.75) and str(c) == "3 0.5 0.25 0.125 0.0625 0.75"
    p = fighter.YoloCrop.from_pixel_coordinates(1280, 720, 100, 50, 300, 50, 100, 250, 300, 250)
    assert p.yolo_crop() == (200 / 1280, 150 / 720, 200 / 1280, 200 / 720)


def test_window_tables(golden_dir):
    w = json.load(open(os.path.join(golden_dir, "windows.json")))
    for (a, b, c, d, e), out in zip(w["args"], w["out"]):
        assert dataset_utils.action_sample_from_frame_middle_out(a, b, c, d, min_frame=e) == out
        tab = dataset_utils.window_index_table(np.array([a]), b, c, max_frames=d, min_frame=e)
        assert tab.dtype == np.int32 and tab[0].tolist() == out
    with pytest.raises(AssertionError):
        dataset_utils.action_sample_from_frame_middle_out(3, 4, 1, 10)
    full = dataset_utils.window_index_table(64, 7, 3, max_frames=64)
    assert full.shape == (64, 7) and full.min() == 0 and full.max() == 63
    for i in (0, 5, 30, 63):
        assert full[i].tolist() == dataset_utils.action_sample_from_frame_middle_out(i, 7, 3, 64)


def test_timeline_loader(golden_dir):
    tl = json.load(open(os.path.join(golden_dir, "timeline.json")))
    path = os.path.join(golden_dir, "sample_log.jsonl")
    for off, want in tl.items():
        gt = timeline.load_ground_truth_from_path(path, log_offset=int(off))
        assert [[[r["num_frames_left"], r["fighter_id"], r["pos_x"]] for r in fr] for fr in gt] == want
    boxes = fighter.boxes_from_timeline(timeline.load_ground_truth_from_path(path))
    assert boxes.shape == (12, 2, 4) and np.isfinite(boxes).all()


def test_timeline_validation(tmp_path):
    p = tmp_path / "bad.jsonl"
    from workloads import synthetic

    recs = synthetic.synth_log_records(3, 2, seed=1)
    recs[1] = recs[1][:1]
    synthetic.write_log(str(p), recs)
    with pytest.raises(AssertionError):
        timeline.load_ground_truth_from_path(str(p))


def test_crop_records():
    from playaid_core_b200.preprocess import crop_records

    rec = crop_records([(0.673046875, 0.5368055555555555, 0.12890625, 0.2625), (0.9999, 0.0, 0.5, 0.5)], [3, 4], 1920, 1080)
    assert rec.dtype == np.int32 and rec.shape == (2, 8)
    assert rec[0].tolist() == [3, 1292, 579, 247, 283, 0, 0, 0] and rec[1].tolist()[:5] == [4, 1919, 0, 960, 540]


def test_synthetic_frames_deterministic():
    from workloads import synthetic

    b = np.array([[[700, 500, 247, 283], [1200, 620, 180, 300]]])
    a = synthetic.synth_frames([5], b, device="cpu")
    c = synthetic.synth_frames([5], b, device="cpu")
    assert a.shape == (1, 1080, 1920, 3) and bool((a == c).all())


def test_load_timeline_from_ai_output_matches_reference(golden_dir, tmp_path):
    """ai_output.yaml -> timeline records, pinned by the reference's own loader
    (oracle/gen_golden.py::gen_ai_timeline); kwargs lift the hard-coded 600 frames / fighter pair."""
    import hashlib
    import json
    import os

    import yaml

    from oracle.gen_golden import golden_ai_output
    from playaid_core_b200.timeline import load_timeline_from_ai_output

    gold = json.load(open(os.path.join(golden_dir, "ai_timeline.json")))
    path = str(tmp_path / "ai_output.yaml")
    with open(path, "w") as f:
        yaml.dump(golden_ai_output(), f)
    tl = load_timeline_from_ai_output(path)
    assert len(tl) == gold["n_frames"] == 600
    for i, frame in gold["samples"].items():
        assert tl[int(i)] == frame
    assert hashlib.sha256(json.dumps(tl, sort_keys=True).encode()).hexdigest() == gold["sha256"]
    # kwargs: another pair, every frame in the file
    small = {"Byleth": {i: {"crop": "0 0.5 0.5 0.1 0.2 0", "action": "Wait"} for i in range(5)},
             "Diddy Kong": {i: {"crop": "0 0.4 0.5 0.1 0.2 0", "action": "Run"} for i in range(5)}}
    with open(path, "w") as f:
        yaml.dump(small, f)
    tl = load_timeline_from_ai_output(path, max_frames=None, fighters=["Byleth", "Diddy Kong"])
    assert len(tl) == 5 and [r["fighter_id"] for r in tl[0]] == [0, 1]
    assert tl[2][1]["fighter_name"] == 39 and tl[2][1]["action"] == "Run" and tl[2][0]["stage_id"] == 86
    # the AI crop string overrides the projected box (fighter.py:503-504)
    from playaid_core_b200.fighter import boxes_from_timeline
    b = boxes_from_timeline(tl)
    assert b.shape == (5, 2, 4) and abs(b[0, 1, 0] - 0.4) < 1e-12 and abs(b[0, 0, 3] - 0.2) < 1e-12
