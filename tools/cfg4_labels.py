"""BASELINE cfg4 on the GPU box: a 7-minute 1080p synthetic match (25 200 frames, 50 400 crops) labelled twice, by
the 16-bit path (f16) and by the fp32-parity path (f16x2). Frames are generated chunk by chunk on the device and
never resident at once. Writes gpurun_out/cfg4_labels.npz (labels / confidences of both streams, boxes) and
gpurun_out/cfg4_agreement.json; tools/cfg4_stats.py then feeds both label streams through the reference's own
timeline + Stats consumers (SURVEY 8d parity gate).   python tools/cfg4_labels.py [n_frames]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from playaid_core_b200.action_detector import ActionDetector
from playaid_core_b200.anim_ontology import ACTIONS
from playaid_core_b200.fighter import boxes_from_records, yolo_pixels_batch
from playaid_core_b200.models.cnn_action_detector import CNNActionDetector
from workloads import synthetic, weights

N = int(sys.argv[1]) if len(sys.argv) > 1 else 25200
H, W, F, B = 1080, 1920, 2, 256
dev = torch.device("cuda", 0)
recs = synthetic.synth_log_records(N, F, seed=4242)
boxes = boxes_from_records([r for f in recs for r in f]).reshape(N, F, 4)
px = yolo_pixels_batch(boxes, W, H)
sd = weights.calibrated_state_dict(0)
dets, streams = {}, {}
for prec in ("f16", "f16x2"):
    dets[prec] = ActionDetector(CNNActionDetector(ACTIONS, sequence_length=7, precision=prec, device=dev).eval().load_state_dict(sd))
    streams[prec] = dets[prec].stream(boxes, H, W)
buf = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
t0 = time.perf_counter()
for s in range(0, N, B):
    n = min(B, N - s)
    frames = synthetic.synth_frames(np.arange(s, s + n), px[s : s + n], device=dev, seed=99, out=buf[:n])
    for prec in streams:
        streams[prec].push(frames[:n])
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out = {}
for prec, st in streams.items():
    out[prec] = (st.label.cpu().numpy().astype(np.int8), st.prob.cpu().numpy(), st.logp.cpu().numpy(), st.status.cpu().numpy())
la, lb = out["f16"][0], out["f16x2"][0]
lpa, lpb = out["f16"][2], out["f16x2"][2]
top2 = np.sort(lpb, axis=-1)
margin = top2[..., -1] - top2[..., -2]
dis = la != lb
rel = np.abs(lpa - lpb).max(-1) / np.abs(lpb).max(-1)
res = {"frames": N, "crops": int(N * F), "seconds_incl_frame_synthesis": dt, "label_agreement": float(1.0 - dis.mean()),
       "disagreements": int(dis.sum()), "max_margin_of_a_disagreement": float(margin[dis].max()) if dis.any() else 0.0,
       "median_margin": float(np.median(margin)), "max_rel_logprob_err_f16_vs_f16x2": float(rel.max()),
       "distinct_labels": int(len(np.unique(lb))), "crop_status_ok_fraction": float((out["f16x2"][3] == 1).mean())}
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/cfg4_labels.npz", label_f16=la, label_f16x2=lb, prob_f16=out["f16"][1].astype(np.float32),
                    prob_f16x2=out["f16x2"][1].astype(np.float32), boxes=boxes)
json.dump(res, open("gpurun_out/cfg4_agreement.json", "w"), indent=1)
print(res)
