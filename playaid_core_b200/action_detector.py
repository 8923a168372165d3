"""`ActionDetector`: the ult_logger-driven crop -> classify path, end to end on one GPU.

The reference spreads this path over an offline script and a deprecated runner
(SURVEY.md 3.2): `process_pairing` (playaid/data_gen_scripts/gen_gt_action_detection.py:26-91)
crops every fighter with `square_crop(frame, 128, padding=30)`; `AIRunner`
(playaid/ai_runner.py:426-520) builds a 7-frame middle-out window per (frame, fighter), runs
`CNNActionDetector`, takes argmax / exp and fills
`ai_output_data[fighter][frame] = {crop, action, predicted_action_confidence}` which
`write_output` dumps as ai_output.yaml (:606-608) for `load_timeline_from_ai_output`
(playaid/timeline.py:52-105). This module composes the same steps on the GPU:

    boxes (host fp64, fighter.py geometry) -> pa_preprocess -> pa_features (once per crop)
    -> feature table -> pa_head over middle-out windows -> labels / log-probs / confidence

`MatchStream` feeds a long match in chunks (frames need not be resident all at once); windows
reach +-27 frames (delta 3), so labels trail the pushed frames by 27.
"""
from __future__ import annotations

import numpy as np
import torch
import yaml

from . import _lib
from .dataset_utils import window_index_table
from .fighter import YoloCrop, boxes_from_timeline
from .models.cnn_action_detector import CNNActionDetector
from .preprocess import crop_records, preprocess_crops, stage_windows


class MatchStream:
    """Streaming state for one match (or one rank's shard of it): crop records and window tables
    (device-resident, built once from the log), the feature table, and how far crops / labels have
    progressed.

    `boxes` holds the frames this stream will be fed, i.e. global frames
    [frame_offset, frame_offset + len(boxes)) of a video with `total_frames` frames; labels are produced
    for global frames [own[0], own[1]) (default: all of them). Window indices are clamped at the ends of
    the *video*, not of the shard, so a shard with a `reach`-frame halo reproduces the single-GPU labels.
    """

    def __init__(self, det: "ActionDetector", boxes: np.ndarray, H: int, W: int, frame_offset: int = 0,
                 total_frames: int | None = None, own: tuple[int, int] | None = None, rec: torch.Tensor | None = None):
        self.det, self.H, self.W = det, int(H), int(W)
        self.N, self.F = int(boxes.shape[0]), int(boxes.shape[1])
        self.offset = int(frame_offset)
        self.total = int(total_frames) if total_frames is not None else self.offset + self.N
        lo, hi = own if own is not None else (self.offset, self.offset + self.N)
        lo = max(lo, det.min_frame)
        self.own = (lo, max(hi, lo))
        self.boxes = boxes
        dev = det.model._device
        A, S = det.model.num_actions, det.num_frames_per_sample
        if rec is None:     # host geometry -> crop records; `rec` = records already on the device (pa_boxes_from_log)
            rec = torch.from_numpy(crop_records(boxes.reshape(-1, 4), np.repeat(np.arange(self.N), self.F), self.W, self.H)).to(dev)
        assert rec.dtype == torch.int32 and tuple(rec.shape) == (self.N * self.F, _lib.BOX_STRIDE) and rec.is_cuda
        self.rec = rec
        n_own = self.own[1] - self.own[0]
        self.feat = torch.zeros((self.N * self.F, 1000), dtype=torch.float32, device=dev)
        self.logp = torch.empty((n_own, self.F, A), dtype=torch.float32, device=dev)
        self.label = torch.full((n_own, self.F), -1, dtype=torch.int32, device=dev)
        self.prob = torch.zeros((n_own, self.F), dtype=torch.float32, device=dev)
        self.status = torch.zeros((self.N, self.F), dtype=torch.int32, device=dev)
        self._hs_recorded = False
        self.pushed = 0   # local frames whose features are in the table
        self.labeled = 0  # own frames whose windows have been classified
        # window indices of every own frame as LOCAL frame numbers -- dataset_utils.py:109-138
        wf = window_index_table(np.arange(self.own[0], self.own[1]), S, det.frame_delta, max_frames=self.total,
                                min_frame=det.min_frame) - self.offset
        assert n_own == 0 or (wf.min() >= 0 and wf.max() < self.N), "shard lacks the halo frames its windows reach"
        self.win_frames = wf
        self.need = wf.max(axis=1) if n_own else np.zeros((0,), np.int64)  # last local frame each window needs
        idx = wf[:, None, :].astype(np.int64) * self.F + np.arange(self.F)[None, :, None]
        self.win_rows = torch.from_numpy(np.ascontiguousarray(idx.astype(np.int32))).to(dev)  # [n_own, F, S]

    def push(self, frames: torch.Tensor, defer_labels: bool = False) -> tuple[int, int]:
        """frames uint8 [n,H,W,3], CUDA or pinned host: local frames [pushed, pushed+n). Returns the [a, b)
        range of own frames (indices into `label`) classified by this call; labels trail the pushed
        frames by the window reach until the last frame arrives. Pinned host frames are not copied
        whole: their crop windows are staged into HBM on the detector's copy stream (overlapping the
        kernels of the previous chunk), or read in place over PCIe when `det.host_mode == "inplace"`.
        `defer_labels=True` leaves the temporal head of this chunk running on the detector's head stream (small,
        latency-bound kernels that then overlap the next chunk's preprocess); call `wait_labels()` before reading
        `label` / `logp` / `prob` on the current stream. By default push() orders the current stream after it."""
        det = self.det
        n = int(frames.shape[0])
        f0 = self.pushed
        assert f0 + n <= self.N
        slot = None
        if not frames.is_cuda and det.host_mode == "stage" and frames.is_contiguous():
            frames, slot = det._stage(frames, self.rec[f0 * self.F : (f0 + n) * self.F], f0)
        rec = self.rec[f0 * self.F : (f0 + n) * self.F].clone()
        rec[:, 0] -= f0  # frame index relative to this chunk
        u8 = det.byte_crops
        crops, _ = preprocess_crops(
            frames, rec, det.output_size, det.padding, swap_rb=True, mean=det.mean, std=det.std,
            dtype=det.model.crop_dtype_u8 if u8 else det.model.crop_dtype, layout=_lib.LAYOUT_NHWC4P,
            out=det._crop_buffer(n * self.F), status=self.status[f0 : f0 + n].view(-1),
        )
        if slot is not None:
            det._stage_release(slot)
        det.model.features(crops, out=self.feat[f0 * self.F : (f0 + n) * self.F], u8=u8)
        self.pushed = f0 + n
        main = torch.cuda.current_stream(det.model._device)
        hs = det._head_stream_for(main)
        if not self._hs_recorded:   # allocated on the current stream, also used on the head stream
            for t in (self.feat, self.logp, self.label, self.prob, self.win_rows):
                t.record_stream(hs)
            self._hs_recorded = True
        det._feat_done.record(main)
        hs.wait_event(det._feat_done)
        with torch.cuda.stream(hs):
            rng = self._label_ready()
            det._labels_done.record(hs)
        if not defer_labels:
            main.wait_event(det._labels_done)
        return rng

    def wait_labels(self) -> None:
        """Order the current stream after the last head launched by push(defer_labels=True)."""
        torch.cuda.current_stream(self.det.model._device).wait_event(self.det._labels_done)

    def _label_ready(self) -> tuple[int, int]:
        det = self.det
        a = self.labeled
        b = int(np.searchsorted(self.need, self.pushed - 1, side="right"))
        if b <= a:
            return (a, a)
        wf = self.win_frames[a:b]
        lo, hi = int(wf.min()), int(wf.max()) + 1  # local feature frames the windows [a, b) touch
        idx = (self.win_rows[a:b] - lo * self.F).view(-1, wf.shape[1]).contiguous()
        # windows that touch a crop the reference would not have produced (off-screen: process_pairing skips it,
        # gen_gt_action_detection.py:54-56) are left unlabelled (-1) by the head kernel
        logp, label, prob = det.model.head(self.feat[lo * self.F : hi * self.F], idx,
                                           status=self.status.view(-1)[lo * self.F : hi * self.F])
        self.logp[a:b] = logp.view(b - a, self.F, -1)
        self.label[a:b] = label.view(b - a, self.F)
        self.prob[a:b] = prob.view(b - a, self.F)
        self.labeled = b
        return (a, b)


class ActionDetector:
    def __init__(
        self,
        model: CNNActionDetector,
        output_size: int = 128,
        padding: int = 30,
        num_frames_per_sample: int | None = None,
        frame_delta: int = 3,
        min_frame: int = 0,
        mean=(0.0, 0.0, 0.0),
        std=(1.0, 1.0, 1.0),
        byte_crops: bool | None = None,
    ):
        self.model = model
        self.output_size = output_size
        self.padding = padding
        self.num_frames_per_sample = num_frames_per_sample or model.sequence_length
        assert self.num_frames_per_sample == model.sequence_length, "window length must equal the Conv1d kernel"
        self.frame_delta = frame_delta
        self.min_frame = min_frame
        self.mean, self.std = tuple(float(v) for v in mean), tuple(float(v) for v in std)
        default_norm = self.mean == (0.0, 0.0, 0.0) and self.std == (1.0, 1.0, 1.0)
        if byte_crops and not default_norm:
            raise ValueError("byte-valued crops fold x = v / 255 into the stem: only valid for mean 0 / std 1")
        # None = automatic: byte-valued crops whenever the reference's own normalisation (no mean / std) is in force
        self._byte_crops = default_norm if byte_crops is None else bool(byte_crops)
        self._crops: torch.Tensor | None = None
        # pinned-host frames: "stage" = pa_stage_windows on a copy stream into one of two HBM frame buffers,
        # "inplace" = the preprocess kernel reads the pinned memory itself
        self.host_mode = "stage"
        self._stage_bufs: list[torch.Tensor] | None = None
        self._stage_i = 0
        self._head_stream: torch.cuda.Stream | None = None
        self._feat_done = self._labels_done = None

    @property
    def head_stream(self) -> torch.cuda.Stream:
        """The side stream the temporal head runs on. Work that consumes labels of `push(defer_labels=True)` can be
        queued there (it is ordered after the head); `torch.cuda.current_stream().wait_stream(det.head_stream)` joins."""
        return self._head_stream_for(torch.cuda.current_stream(self.model._device))

    def _head_stream_for(self, main: torch.cuda.Stream) -> torch.cuda.Stream:
        if self._head_stream is None:
            dev = self.model._device
            self._head_stream = torch.cuda.Stream(dev)
            self._feat_done, self._labels_done = torch.cuda.Event(), torch.cuda.Event()
            # everything queued so far (feature table zero-fill, uploads) precedes the first head
            first = torch.cuda.Event()
            first.record(main)
            self._head_stream.wait_event(first)
        return self._head_stream

    def _stage(self, host_frames: torch.Tensor, rec: torch.Tensor, frame_base: int):
        dev = self.model._device
        if self._stage_bufs is None or self._stage_bufs[0].shape[1:] != host_frames.shape[1:] or self._stage_bufs[0].shape[0] < host_frames.shape[0]:
            self._stage_bufs = [torch.empty(tuple(host_frames.shape), dtype=torch.uint8, device=dev) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(dev, priority=-1)
            self._stage_ready = [torch.cuda.Event() for _ in range(2)]
            self._stage_free = [torch.cuda.Event() for _ in range(2)]
            # the buffers may be recycled blocks of the caching allocator: order the copy stream after whatever the
            # allocating stream still has queued on them
            alloc_done = torch.cuda.Event()
            alloc_done.record(torch.cuda.current_stream(dev))
            self._copy_stream.wait_event(alloc_done)
        slot = self._stage_i & 1
        self._stage_i += 1
        buf = self._stage_bufs[slot][: host_frames.shape[0]]
        main = torch.cuda.current_stream(dev)
        self._copy_stream.wait_event(self._stage_free[slot])   # the kernels that read this buffer two chunks ago are done
        with torch.cuda.stream(self._copy_stream):
            stage_windows(host_frames, rec, buf, self.padding, frame_base)
            self._stage_ready[slot].record(self._copy_stream)
        main.wait_event(self._stage_ready[slot])
        return buf, slot

    def _stage_release(self, slot: int) -> None:
        self._stage_free[slot].record(torch.cuda.current_stream(self.model._device))

    @property
    def byte_crops(self) -> bool:
        """With the reference's normalisation (x = v / 255, no mean / std: ult_action_dataset.py:349-359) the crops keep
        the byte values and the stem applies 1/255 in fp32: the 16-bit input is exact in every precision."""
        return self._byte_crops

    def _crop_buffer(self, n: int) -> torch.Tensor:
        planes = 2 if (self.model.split and not self.byte_crops) else 1
        shape = (planes, n, self.output_size, self.output_size + 8, 4) if planes == 2 else (n, self.output_size, self.output_size + 8, 4)
        if self._crops is None or tuple(self._crops.shape) != shape or self._crops.dtype != self.model.act_dtype:
            self._crops = torch.empty(shape, dtype=self.model.act_dtype, device=self.model._device)
        return self._crops

    def stream(self, boxes: np.ndarray, H: int, W: int, **shard) -> MatchStream:
        """boxes float64 [N,F,4]: every (frame, fighter) box of the match (from the ult_logger log).
        `shard` = frame_offset / total_frames / own for one rank's slice (see parallel.frame_shard)."""
        return MatchStream(self, boxes, H, W, **shard)

    def stream_from_log(self, timeline, H: int, W: int, **shard) -> MatchStream:
        """`timeline` = `load_ground_truth_from_path(log)` records (per frame, per fighter). The box geometry runs on the
        device (`pa_boxes_from_log`): only nine doubles per record cross PCIe, the crop records never
        exist on the host. Fighters in `fighter_id` order like `update_fighters_from_timeline` (timeline.py:186-201)."""
        from .fighter import boxes_from_records_device, log_record_array

        n_f = len(timeline[0]) if timeline else 0
        flat = [r for frame in timeline for r in sorted(frame, key=lambda x: x["fighter_id"])]
        if any("crop" in r for r in flat):    # AI crop overrides (fighter.py:503-504) take the host path
            return self.stream(boxes_from_timeline(timeline), H, W, **shard)
        arr = log_record_array(flat, np.repeat(np.arange(len(timeline)), n_f))
        boxes_d, rec_d = boxes_from_records_device(arr, W, H, self.model._device)
        return MatchStream(self, boxes_d.cpu().numpy().reshape(len(timeline), n_f, 4), H, W, rec=rec_d, **shard)

    def classify_clip(self, frames: torch.Tensor, boxes: np.ndarray, chunk: int = 256) -> dict:
        """frames uint8 CUDA [N,H,W,3] (BGR, as decoded by cv2), boxes float64 [N,F,4] normalised.
        Returns device tensors: label [N,F] int32, logp [N,F,A], prob [N,F], status [N,F]."""
        N, H, W, _ = frames.shape
        st = self.stream(boxes[:N], H, W)
        for s in range(0, N, chunk):
            st.push(frames[s : s + chunk])
        label, logp, prob = st.label, st.logp, st.prob
        if self.min_frame > 0:  # AIRunner-style 1-indexed runs leave the frames below min_frame unlabelled
            pad = self.min_frame
            label = torch.cat([torch.full((pad, st.F), -1, dtype=label.dtype, device=label.device), label])
            logp = torch.cat([torch.zeros((pad,) + tuple(logp.shape[1:]), dtype=logp.dtype, device=logp.device), logp])
            prob = torch.cat([torch.zeros((pad, st.F), dtype=prob.dtype, device=prob.device), prob])
        self.check_status(st.status)
        return {"label": label, "logp": logp, "prob": prob, "status": st.status}

    @staticmethod
    def check_status(status: torch.Tensor) -> None:
        """Raise where the reference raises: a zero-row window makes ImageOps.contain divide by zero
        (fighter.py:356 lets ZeroDivisionError escape). Off-screen crops (status 0) are not an error: they are
        skipped like process_pairing skips them, their windows carry label -1."""
        st = status.detach().cpu().numpy()
        if (st == _lib.CROP_ZERO_DIV).any():
            i = int(np.argwhere(st.reshape(-1) == _lib.CROP_ZERO_DIV)[0, 0])
            raise ZeroDivisionError(f"division by zero in square_crop of crop {i} (zero-size box with padding)")
        if (st == _lib.CROP_TOO_LARGE).any():
            i = int(np.argwhere(st.reshape(-1) == _lib.CROP_TOO_LARGE)[0, 0])
            raise _lib.PlayaidLibraryError(f"crop {i}: window exceeds the preprocess kernel's staging limits")

    def classify_shard(self, frames_halo: torch.Tensor, boxes_halo: np.ndarray, halo_lo: int, own: tuple[int, int],
                       total_frames: int, chunk: int = 256) -> dict:
        """One rank's slice of a long video: `frames_halo` / `boxes_halo` cover global frames
        [halo_lo, halo_lo + n) (own range plus the window halo, parallel.frame_shard); labels are
        returned for global frames [own[0], own[1])."""
        n, H, W, _ = frames_halo.shape
        st = self.stream(boxes_halo[:n], H, W, frame_offset=halo_lo, total_frames=total_frames, own=own)
        for s in range(0, n, chunk):
            st.push(frames_halo[s : s + chunk])
        self.check_status(st.status)
        return {"label": st.label, "logp": st.logp, "prob": st.prob, "status": st.status}

    def classify_timeline(self, frames: torch.Tensor, timeline, chunk: int = 256) -> dict:
        """`timeline` = `load_ground_truth_from_path(log)` records; boxes via the fighter geometry."""
        return self.classify_clip(frames, boxes_from_timeline(timeline)[: frames.shape[0]], chunk)

    # ------------------------------------------------------------------ ai_output.yaml (ai_runner.py:493-520,606-608)
    def ai_output(self, result: dict, boxes: np.ndarray, fighter_names: list[str]) -> dict:
        """{fighter: {frame_idx: {crop, action, predicted_action_confidence}}} like AIRunner's
        `ai_output_data.to_dict()`; confidence = float(exp(logp)[label]) * 100.0 (ai_runner.py:476-477)."""
        label = result["label"].cpu().numpy()
        prob = result["prob"].cpu().numpy()
        out = {}
        for k, name in enumerate(fighter_names):
            per = {}
            for i in range(label.shape[0]):
                if label[i, k] < 0:
                    continue
                c = YoloCrop(*[float(v) for v in boxes[i, k]])
                per[i] = {
                    "crop": str(c),
                    "action": self.model.actions[int(label[i, k])],
                    "predicted_action_confidence": float(prob[i, k]) * 100.0,
                }
            out[name] = per
        return out

    @staticmethod
    def write_output(ai_output: dict, path: str) -> None:
        with open(path, "w") as f:
            yaml.dump(ai_output, f)
