"""`AIRunner`: the reference runner's action-recognition surface over the GPU path.

Reference playaid/ai_runner.py keeps one cropped jpg per (fighter, frame) on disk, and for every frame
number 1..max_frames-1 re-reads seven of them, resizes, runs `CNNActionDetector` at batch 1 and fills
`ai_output_data[fighter][frame_num - 1]` (:426-520); `write_output` dumps that as ai_output.yaml (:606-608),
`load_ai_output` reads it back (:592-604). This class keeps those method names, argument meaning, the
1-indexed frame numbers and the return shapes, but takes the decoded frames and the ult_logger boxes
directly: crops are cut on the GPU (`pa_preprocess`), features are computed once per (frame, fighter) and
`run_action_recognition` classifies the whole clip in batches.

Frame numbering (as in the reference): frame number k is `frames[k - 1]`; `max_frames` is the number of the
last frame; frames 1 .. max_frames-1 are classified and window indices are clamped to that range
(`action_sample_from_frame_middle_out(..., max_frames=self.max_frames, min_frame=1)`, :430-440).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import yaml

from . import _lib
from .action_detector import ActionDetector
from .dataset_utils import action_sample_from_frame_middle_out
from .fighter import YoloCrop
from .models.cnn_action_detector import CNNActionDetector
from .preprocess import crop_records, preprocess_crops


class AIRunner:
    def __init__(self, frames: torch.Tensor, boxes: np.ndarray, fighters: list[str], model: CNNActionDetector,
                 ai_output_file: str | None = None, dataset_args: dict | None = None, char_list: list[str] | None = None,
                 chunk: int = 256):
        """frames uint8 [N,H,W,3] BGR (CUDA or pinned host) = frame numbers 1..N; boxes float64 [N,F,4]
        normalised (cx, cy, w, h) per frame and fighter (`fighter.boxes_from_timeline`); `fighters` names the F
        columns. `dataset_args` may carry `num_frames_per_sample` (7) and a one-element `frame_delta` list."""
        assert frames.ndim == 4 and frames.shape[0] == boxes.shape[0] and boxes.shape[1] == len(fighters)
        self.frames, self.boxes, self.fighters = frames, np.asarray(boxes, dtype=np.float64), list(fighters)
        self.model = model
        self.ai_output_file = ai_output_file
        self.dataset_args = dict(dataset_args or {})
        self.char_list = list(char_list) if char_list is not None else list(fighters)
        self.max_frames = int(frames.shape[0])
        self.chunk = chunk
        deltas = self.dataset_args.get("frame_delta", [3])
        assert len(deltas) == 1, "a random frame_delta choice would make labels non-deterministic; pass one value"
        self.frame_delta = int(deltas[0])
        self.num_frames_per_sample = int(self.dataset_args.get("num_frames_per_sample", 7))
        self.detector = ActionDetector(model, num_frames_per_sample=self.num_frames_per_sample, frame_delta=self.frame_delta,
                                       min_frame=1)
        ok, data = self.load_ai_output()
        self.ai_output_data = data if ok else {}

    # ------------------------------------------------------------------ one window (ai_runner.py:426-491)
    def get_action_recognition_input_for_frame(self, frame: int, fighter: str):
        """-> (input_frames float32 [1,S,3,128,128] in [0,1] RGB on the model's device, [S] uint8 RGB HWC arrays)."""
        k = self.fighters.index(fighter)
        nums = action_sample_from_frame_middle_out(frame, num_frames_per_sample=self.num_frames_per_sample,
                                                   frame_delta=self.frame_delta, max_frames=self.max_frames, min_frame=1)
        idx = np.asarray(nums, dtype=np.int64) - 1
        H, W = int(self.frames.shape[1]), int(self.frames.shape[2])
        dev = self.model._device
        rec = torch.from_numpy(crop_records(self.boxes[idx, k], idx, W, H)).to(dev)
        det = self.detector
        x, status = preprocess_crops(self.frames, rec, det.output_size, det.padding, swap_rb=True, dtype=_lib.DTYPE_F32,
                                     layout=_lib.LAYOUT_NCHW)
        u8, _ = preprocess_crops(self.frames, rec, det.output_size, det.padding, swap_rb=True, dtype=_lib.DTYPE_U8,
                                 layout=_lib.LAYOUT_NHWC)
        bad = (status != _lib.CROP_OK).nonzero()
        assert bad.numel() == 0, f"Failed to get frame {int(nums[int(bad[0])])} for {fighter}"   # the reference asserts on a missing crop file
        return x.unsqueeze(0), [a for a in u8.cpu().numpy()]

    def action_recognition(self, frame_num: int, fighter: str):
        input_frames, frames = self.get_action_recognition_input_for_frame(frame_num, fighter)
        predictions = self.model(input_frames)
        predicted_action_id = int(torch.argmax(predictions))
        confidence = float(torch.exp(predictions)[0][predicted_action_id]) * 100.0
        crop = YoloCrop(*[float(v) for v in self.boxes[frame_num - 1, self.fighters.index(fighter)]])
        return (
            input_frames,
            self.char_list.index(fighter),
            torch.tensor(predicted_action_id),
            {"char": fighter, "predicted_action": self.model.actions[predicted_action_id], "confidence": confidence,
             "crop": crop, "frames": frames},
        )

    # ------------------------------------------------------------------ whole clip (ai_runner.py:493-520)
    def run_action_recognition(self, overwrite: bool = False):
        todo = [f for f in self.fighters if overwrite or not self.ai_output_data.get(f, {}).get(0, {}).get("action")]
        if not todo:
            return self.ai_output_data
        N = self.max_frames
        H, W = int(self.frames.shape[1]), int(self.frames.shape[2])
        # global frame numbers 1..N live at array rows 0..N-1; windows clamp to [1, N-1]; labels for 1..N-1
        st = self.detector.stream(self.boxes, H, W, frame_offset=1, total_frames=N, own=(1, N))
        for s in range(0, N, self.chunk):
            st.push(self.frames[s : s + self.chunk])
        label, prob = st.label.cpu().numpy(), st.prob.cpu().numpy()
        # the reference asserts on every crop it needs ("Failed to get square crop from frame j", ai_runner.py:418-419;
        # "Failed to get frame", :447): a window that touches a missing crop is an error here too, not a label
        self.detector.check_status(st.status)
        status = st.status.cpu().numpy()
        for fighter in todo:
            k = self.fighters.index(fighter)
            bad = np.nonzero(status[:, k] != _lib.CROP_OK)[0]
            assert bad.size == 0, f"Failed to get square crop from frame {int(bad[0]) + 1} for {fighter}"
        for fighter in todo:
            k = self.fighters.index(fighter)
            per = self.ai_output_data.setdefault(fighter, {})
            for frame_num in range(1, N):
                e = per.setdefault(frame_num - 1, {})   # "Yolo is 1 indexed, switch to 0 indexed" (:515)
                e["crop"] = str(YoloCrop(*[float(v) for v in self.boxes[frame_num - 1, k]]))
                e["action"] = self.model.actions[int(label[frame_num - 1, k])]
                e["predicted_action_confidence"] = float(prob[frame_num - 1, k]) * 100.0
        return self.ai_output_data

    # ------------------------------------------------------------------ yaml (ai_runner.py:592-608)
    def load_ai_output(self):
        if not self.ai_output_file or not os.path.exists(self.ai_output_file):
            return False, {}
        with open(self.ai_output_file, "r") as f:
            try:
                data = yaml.safe_load(f)
                return True, dict(data)
            except Exception:
                return False, {}

    def write_output(self):
        with open(self.ai_output_file, "w") as f:
            yaml.dump(self.ai_output_data, f)
