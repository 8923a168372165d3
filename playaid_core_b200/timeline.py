"""ult_logger JSONL -> per-frame fighter records (host side; feeds bboxes into the GPU path).

Mirrors reference playaid/timeline.py:204-280 (`load_ground_truth_from_path`): one JSON object
per line, one line per fighter per frame; `log_offset` frames skipped from the top; a jump of
more than one in `num_frames_left` repeats the frame being assembled; fighter ids rewritten to
0..n-1 in `fighter_id` order; `validate` asserts exactly two fighters per frame.

`fighters_per_frame` generalises the hard-coded 2 (reference timeline.py:243,274-279) for the
4-fighter configuration; with the default of 2 the behaviour is the reference's.
"""
from __future__ import annotations

import json

import yaml

from .anim_ontology import FIGHTER_NAME_TO_ENUM


def load_ground_truth_from_path(
    label_path: str, validate: bool = True, log_offset: int = 0, max_lines=0, fighters_per_frame: int = 2
):
    fpf = fighters_per_frame
    ground_truth = []
    prev_left = -1
    index = 0
    skipped = 0

    if log_offset < 0:
        # reference timeline.py:219-228: duplicate the first frame |log_offset| times
        with open(label_path, "r") as f:
            first = [json.loads(f.readline()) for _ in range(fpf)]
        ground_truth = [first] * abs(log_offset)
        index += fpf * abs(log_offset)
        log_offset = 0

    with open(label_path, "r") as f:
        for line in f:
            if max_lines and index > max_lines:
                break
            if skipped < fpf * log_offset:
                skipped += 1
                continue
            rec = json.loads(line)
            frame_number = index // fpf
            if frame_number >= len(ground_truth):
                ground_truth.append([])
            gap = prev_left - rec["num_frames_left"]
            if prev_left > 0 and gap > 1:
                # the logger dropped frames: alias the frame under assembly gap-1 more times
                ground_truth += [ground_truth[-1]] * (gap - 1)
                index += (gap - 1) * fpf
            ground_truth[frame_number].append(rec)
            index += 1
            prev_left = rec["num_frames_left"]

    for i, frame in enumerate(ground_truth):
        frame = sorted(frame, key=lambda r: r["fighter_id"])
        for j, rec in enumerate(frame):
            rec["fighter_id"] = j
        ground_truth[i] = frame

    if validate:
        for i, frame in enumerate(ground_truth):
            assert len(frame) == fpf, (
                f"there should be the ground truth for {fpf} players for every frame, found "
                + f"{len(frame)} for frame #{i}"
            )
    return ground_truth


# The record every AI-labelled frame starts from before its {crop, action, ...} entry is merged in
# (reference timeline.py:64-91): a fixed camera / fighter state, so that `Fighter.set_from_json` has every
# key it reads; the AI `crop` string then overrides the projected box (fighter.py:503-504).
_AI_BASE_RECORD = {
    "raw_animation_frame_num": 0, "attack_connected": False, "camera_fov": 30.0,
    "camera_position": {"x": 0.0002484553260728717, "y": 15.847139358520508, "z": 148.460693359375},
    "camera_target_position": {"x": 0.0002776149194687605, "y": 11.162917137145996, "z": 0.0},
    "can_act": True, "damage": 0.0, "facing": 1.0, "hitstun_left": 0.0, "motion_kind": 19292652517,
    "num_frames_left": 54000, "pos_x": -50.0, "pos_y": 0.21623137593269348, "shield_size": 50.0,
    "stage_id": 86, "status_kind": 0, "stock_count": 20,
}


def load_timeline_from_ai_output(file_path: str, max_frames: int | None = 600, fighters=("Joker", "Pikachu"),
                                 fighter_to_player_id: dict | None = None, fighter_name_to_enum: dict | None = None):
    """ai_output.yaml (written by `ActionDetector.write_output` / `AIRunner.write_output`) -> the per-frame record
    list `update_fighters_from_timeline` consumes, as reference playaid/timeline.py:52-105 builds it: for every
    frame and fighter a copy of the fixed base record with `fighter_id` / `fighter_name` filled in and the
    frame's AI entry (`crop`, `action`, `predicted_action_confidence`, ...) merged over it.

    With no keyword arguments the reference's hard-coded run is reproduced (600 frames; Joker = player 1,
    Pikachu = player 0). `max_frames=None` takes every frame the file holds for the first fighter;
    `fighter_to_player_id` defaults to the position in `fighters` for any other pair."""
    with open(file_path, "r") as f:
        ai_output = yaml.safe_load(f)
    fighters = list(fighters)
    if fighter_to_player_id is None:
        if fighters == ["Joker", "Pikachu"]:
            fighter_to_player_id = {"Pikachu": 0, "Joker": 1}
        else:
            fighter_to_player_id = {name: i for i, name in enumerate(fighters)}
    enum = dict(FIGHTER_NAME_TO_ENUM)
    enum.update(fighter_name_to_enum or {})
    if max_frames is None:
        max_frames = max(ai_output[fighters[0]].keys()) + 1 if ai_output[fighters[0]] else 0
    timeline = []
    for i in range(max_frames):
        frame = []
        for name in fighters:
            rec = {k: (dict(v) if isinstance(v, dict) else v) for k, v in _AI_BASE_RECORD.items()}
            rec["fighter_id"] = fighter_to_player_id[name]
            rec["fighter_name"] = enum[name]
            rec.update(ai_output[name][i])   # KeyError on a missing frame, like the reference
            frame.append(rec)
        timeline.append(frame)
    return timeline
