"""How many crops of the random-box parity set end in which status (and which route)?  python tools/pp_status_count.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.gen_golden import golden_frames
from playaid_core_b200 import _lib
from playaid_core_b200.preprocess import crop_records, preprocess_crops
rng = np.random.default_rng(123)
frames = torch.from_numpy(np.stack(golden_frames())).cuda()
n = 600
boxes = np.stack([rng.uniform(-0.05, 1.05, n), rng.uniform(-0.05, 1.05, n), rng.uniform(0.005, 0.6, n), rng.uniform(0.005, 0.9, n)], 1)
fids = rng.integers(0, 3, n)
for pad in (0, 30, 7):
    rec = torch.from_numpy(crop_records(boxes, fids, 1920, 1080)).cuda()
    out, st = preprocess_crops(frames, rec, 128, pad, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
    st = st.cpu().numpy()
    print("pad", pad, {int(v): int((st == v).sum()) for v in np.unique(st)})
    big = np.nonzero(st == -7)[0]
    for i in big[:8]:
        print("   too large:", [int(boxes[i, 0] * 1920), int(boxes[i, 1] * 1080), int(boxes[i, 2] * 1920), int(boxes[i, 3] * 1080)])
