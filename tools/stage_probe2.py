"""Diagnostic: which part of the compute path slows pa_stage_windows down when it runs beside it?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from playaid_core_b200 import _lib
from playaid_core_b200.anim_ontology import ACTIONS
from playaid_core_b200.fighter import boxes_from_records, yolo_pixels_batch
from playaid_core_b200.models.cnn_action_detector import CNNActionDetector
from playaid_core_b200.preprocess import crop_records, stage_windows, preprocess_crops
from workloads import synthetic, weights

H, W, B, F = 1080, 1920, 256, 2
dev = torch.device("cuda", 0)
recs = synthetic.synth_log_records(B, F, seed=2024)
boxes = boxes_from_records([r for f in recs for r in f]).reshape(B, F, 4)
px = yolo_pixels_batch(boxes, W, H)
frames = synthetic.synth_frames(np.arange(B), px[:B], device=dev, seed=1)
host = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory(); host.copy_(frames)
rec = torch.from_numpy(crop_records(boxes.reshape(-1, 4), np.repeat(np.arange(B), F), W, H)).to(dev)
buf = torch.empty_like(frames)
model = CNNActionDetector(ACTIONS, sequence_length=7, precision="f16", device=dev).eval()
model.load_state_dict(weights.default_state_dict(0))
crops = torch.empty((B * F, 128, 136, 4), dtype=model.act_dtype, device=dev)
feat = torch.empty((B * F, 1000), dtype=torch.float32, device=dev)
cs = torch.cuda.Stream(dev, priority=-1)
main = torch.cuda.current_stream()

def pre():
    preprocess_crops(frames, rec, 128, 30, swap_rb=True, dtype=model.crop_dtype, layout=_lib.LAYOUT_NHWC4P, out=crops)
def conv():
    model.features(crops, out=feat)
pre(); conv(); stage_windows(host, rec, buf, 30, 0); torch.cuda.synchronize()

def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record()
    return e0, e1, n

for name, other, reps in (("nothing", None, 0), ("preprocess", pre, 12), ("features", conv, 12)):
    torch.cuda.synchronize()
    with torch.cuda.stream(cs):
        s = timed(lambda: stage_windows(host, rec, buf, 30, 0), 4)
    o = timed(other, reps) if other else None
    torch.cuda.synchronize()
    msg = f"stage beside {name}: {s[0].elapsed_time(s[1]) / s[2]:.3f} ms/launch"
    if o: msg += f"; {name} {o[0].elapsed_time(o[1]) / o[2]:.3f} ms/launch"
    print(msg)
for name, other in (("preprocess", pre), ("features", conv)):
    torch.cuda.synchronize()
    o = timed(other, 6); torch.cuda.synchronize()
    print(f"{name} alone {o[0].elapsed_time(o[1]) / o[2]:.3f} ms/launch")
