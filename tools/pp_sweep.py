"""Times pa_preprocess alone on one bench batch for the PA_PP_* settings in the environment."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from playaid_core_b200 import _lib
from playaid_core_b200.fighter import boxes_from_records, yolo_pixels_batch
from playaid_core_b200.preprocess import crop_records, preprocess_crops
from workloads import synthetic
N = 256
recs = synthetic.synth_log_records(N, 2, seed=2024)
boxes = boxes_from_records([r for f in recs for r in f]).reshape(N, 2, 4)
frames = synthetic.synth_frames(np.arange(N), yolo_pixels_batch(boxes, 1920, 1080), device="cuda")
rec_np = crop_records(boxes.reshape(-1, 4), np.repeat(np.arange(N), 2), 1920, 1080)
order = os.environ.get("PP_ORDER", "none")     # experiment: launch order of the crops (largest window first / last)
if order != "none":
    area = np.maximum(rec_np[:, 3], rec_np[:, 4]).astype(np.int64) + 60
    idx = np.argsort(-area if order == "desc" else area, kind="stable")
    rec_np = np.ascontiguousarray(rec_np[idx])
rec = torch.from_numpy(rec_np).cuda()
out = None
for _ in range(3):
    out, st = preprocess_crops(frames, rec, 128, 30, dtype=_lib.DTYPE_F16, layout=_lib.LAYOUT_NHWC4P, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    preprocess_crops(frames, rec, 128, 30, dtype=_lib.DTYPE_F16, layout=_lib.LAYOUT_NHWC4P, out=out)
e1.record(); torch.cuda.synchronize()
print(f"order={order} threads={os.environ.get('PA_PP_THREADS')} smem={os.environ.get('PA_PP_SMEM_KB')} xb={os.environ.get('PA_PP_XB')}: {e0.elapsed_time(e1)/10:.3f} ms  checksum {int(out.view(torch.int16).to(torch.int64).sum())}")
