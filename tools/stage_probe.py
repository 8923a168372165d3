"""Diagnostic: how fast does pa_stage_windows pull windows over PCIe alone, and does it overlap the compute path?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from playaid_core_b200 import _lib
from playaid_core_b200.action_detector import ActionDetector
from playaid_core_b200.anim_ontology import ACTIONS
from playaid_core_b200.fighter import boxes_from_records, yolo_pixels_batch
from playaid_core_b200.models.cnn_action_detector import CNNActionDetector
from playaid_core_b200.preprocess import crop_records, stage_windows
from workloads import synthetic, weights

H, W, B, F = 1080, 1920, 256, 2
dev = torch.device("cuda", 0)
recs = synthetic.synth_log_records(10800, F, seed=2024)[: B * 16]
boxes = boxes_from_records([r for f in recs for r in f]).reshape(B * 16, F, 4)
px = yolo_pixels_batch(boxes, W, H)
frames = synthetic.synth_frames(np.arange(B), px[:B], device=dev, seed=1)
host = [torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
for h in host: h.copy_(frames)
rec = torch.from_numpy(crop_records(boxes[:B].reshape(-1, 4), np.repeat(np.arange(B), F), W, H)).to(dev)
buf = torch.empty_like(frames)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2): stage_windows(host[0], rec, buf, 30, 0)
torch.cuda.synchronize()
e0.record()
for i in range(5): stage_windows(host[i % 2], rec, buf, 30, 0)
e1.record(); torch.cuda.synchronize()
def wbytes(p4, pad=30):
    cx, cy, cw, ch = [p4[..., i].astype(np.int64) for i in range(4)]
    half = np.maximum(cw, ch) // 2
    y0 = np.maximum(cy - half - pad, 0); y1 = np.minimum(cy + half + pad, H)
    x0 = np.maximum(cx - half - pad, 0); x1 = np.minimum(cx + half + pad, W)
    return int((3 * np.maximum(y1 - y0, 0) * np.maximum(x1 - x0, 0)).sum())
ms = e0.elapsed_time(e1) / 5
print("stage alone ms/chunk", ms, "window MB", wbytes(px[:B]) / 1e6, "GB/s", wbytes(px[:B]) / ms / 1e6)
print("mean window MB/chunk over chunks 0..7", wbytes(px[: 8 * B]) / 8e6)
e0.record()
for i in range(3): buf.copy_(host[i % 2], non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("whole-frame memcpy ms/chunk", ms, "GB/s", B * H * W * 3 / ms / 1e6)

ctx = _lib.Context.get(dev)
model = CNNActionDetector(ACTIONS, sequence_length=7, precision="f16", device=dev).eval()
model.load_state_dict(weights.default_state_dict(0))
det = ActionDetector(model)
for mode in ("device", "stage", "inplace"):
    det.host_mode = "inplace" if mode == "inplace" else "stage"
    st = det.stream(boxes, H, W)
    src = (lambda i: frames) if mode == "device" else (lambda i: host[i % 2])
    st.push(src(0)); st.push(src(1)); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for i in range(6): st.push(src(i))
    t1 = time.perf_counter()
    e1.record(); torch.cuda.synchronize()
    print(mode, "ms/chunk", e0.elapsed_time(e1) / 6, "host enqueue ms/chunk", (t1 - t0) * 1e3 / 6)
    ctx.profile_begin()
    for i in range(6): st.push(src(i))
    prof = ctx.profile_end()
    print("   spans:", {k: round(v[1] / 6, 3) for k, v in prof.items() if v[1] / 6 > 0.05})
