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
