"""Debug driver for the tensor-core preprocess path: a few crops of one synthetic frame, u8 output, against the C oracle.
python tools/pp_tc_debug.py [n_crops]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import resample
from playaid_core_b200 import _lib
from playaid_core_b200.preprocess import crop_records, preprocess_crops
from workloads import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
resample.build()
rng = np.random.default_rng(3)
boxes = np.stack([rng.uniform(0.3, 0.7, n), rng.uniform(0.35, 0.65, n), rng.uniform(0.08, 0.2, n), rng.uniform(0.15, 0.4, n)], 1)
boxes[0] = (0.673046875, 0.5368055555555555, 0.12890625, 0.2625)
px = np.array([[[700, 500, 247, 283], [1200, 620, 180, 300]]])
frame = synthetic.synth_frames([5], px, device="cuda")
rec = torch.from_numpy(crop_records(boxes, np.zeros(n, np.int64), 1920, 1080)).cuda()
out, st = preprocess_crops(frame, rec, 128, 30, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
torch.cuda.synchronize()
print("status", st.cpu().numpy()[:8])
fr = frame.cpu().numpy()[0]
bad = 0
for i in range(n):
    ok, want = resample.square_crop(fr, tuple(boxes[i]), 128, 30)
    got = out[i].cpu().numpy()
    d = (got != want)
    if d.any():
        bad += 1
        ys, xs, cs = np.nonzero(d)
        print(f"crop {i}: {int(d.sum())} bytes differ; rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}; first got {got[ys[0], xs[0]]} want {want[ys[0], xs[0]]}")
print("crops with differences:", bad, "of", n)
