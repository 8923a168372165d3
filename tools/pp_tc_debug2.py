"""Golden crop cases one at a time through pa_preprocess (finds the case that breaks the tensor-core path).
python tools/pp_tc_debug2.py [start]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import resample
from oracle.gen_golden import golden_frames
from playaid_core_b200 import _lib
from playaid_core_b200.preprocess import crop_records, preprocess_crops

resample.build()
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "crops.npz"))
frames_np = np.stack(golden_frames())
frames = torch.from_numpy(frames_np).cuda()
start = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = len(g["ok"])
for i in range(start, n):
    box, fid, pad = g["box"][i], int(g["frame_id"][i]), int(g["padding"][i])
    px = (int(box[0] * 1920), int(box[1] * 1080), int(box[2] * 1920), int(box[3] * 1080))
    print(f"case {i}: frame {fid} pad {pad} px {px}", flush=True)
    rec = torch.from_numpy(crop_records(box[None], [fid], 1920, 1080)).cuda()
    out, st = preprocess_crops(frames, rec, 128, pad, swap_rb=False, dtype=_lib.DTYPE_U8, layout=_lib.LAYOUT_NHWC)
    torch.cuda.synchronize()
    s = int(st.cpu()[0])
    if s == 1:
        ok, want = resample.square_crop(frames_np[fid], tuple(box), 128, pad)
        got = out[0].cpu().numpy()
        if not ok or (got != want).any():
            d = got != want
            ys, xs, cs = np.nonzero(d)
            print(f"   DIFF {int(d.sum())} bytes rows {ys.min()}..{ys.max()} cols {xs.min()}..{xs.max()}", flush=True)
print("done")
