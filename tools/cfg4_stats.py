"""BASELINE cfg4, consumer side (runs in the build container, where /root/reference is mounted): the two label streams
of tools/cfg4_labels.py (16-bit f16 path, fp32-parity f16x2 path) are written as ai_output.yaml by
`ActionDetector.ai_output / write_output`, read back by the REFERENCE's `load_timeline_from_ai_output`
(timeline.py:52-105, its hard-coded first 600 frames) and by ours (whole match), driven through the reference's
`update_fighters_from_timeline` + `Stats.record_frame` (timeline.py:186-201, stats.py:71-140), and the resulting
`Stats.stats` dicts are compared (SURVEY 8d parity gate).   python tools/cfg4_stats.py gpurun_out/cfg4_labels.npz out.json"""
import hashlib, json, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims
ref_shims.install()
import playaid.constants as constants
constants.AI_CACHE = tempfile.mkdtemp()     # Stats.__init__ makes directories under it (stats.py:65-67)
from playaid.fighter import Fighter
from playaid.stats import Stats
from playaid.timeline import load_timeline_from_ai_output as ref_loader, update_fighters_from_timeline
from playaid_core_b200.anim_ontology import ACTIONS
from playaid_core_b200.fighter import YoloCrop
from playaid_core_b200.timeline import load_timeline_from_ai_output as our_loader
import yaml

z = np.load(sys.argv[1])
boxes = z["boxes"]
names = ["Joker", "Pikachu"]   # the pair the reference loader is hard-wired to


def ai_output(label, prob):
    out = {}
    for k, name in enumerate(names):
        out[name] = {i: {"crop": str(YoloCrop(*[float(v) for v in boxes[i, k]])), "action": ACTIONS[int(label[i, k])],
                         "predicted_action_confidence": float(prob[i, k]) * 100.0} for i in range(label.shape[0])}
    return out


def run_stats(timeline):
    stats = Stats("/tmp/cfg4/match.mp4")
    fighters = []
    for i, frame in enumerate(timeline):
        fighters = update_fighters_from_timeline(i, frame, fighters)
        stats.record_frame(fighters)
    return stats.stats.to_dict()


def canon(d):
    if isinstance(d, dict):
        return [[repr(k), canon(v)] for k, v in sorted(d.items(), key=lambda kv: repr(kv[0]))]   # keys mix str and int
    if isinstance(d, (list, tuple)):
        return [canon(v) for v in d]
    return d if isinstance(d, (int, float, str, bool, type(None))) else str(d)


def digest(d):
    return hashlib.sha256(json.dumps(canon(d)).encode()).hexdigest()


res = {}
tmp = tempfile.mkdtemp()
stats = {}
for prec in ("f16", "f16x2"):
    path = os.path.join(tmp, f"ai_output_{prec}.yaml")
    with open(path, "w") as f:
        yaml.dump(ai_output(z[f"label_{prec}"], z[f"prob_{prec}"]), f)
    t_ref = ref_loader(path)                                             # reference loader: first 600 frames
    t_our = our_loader(path)                                             # ours with the reference defaults
    assert t_ref == t_our, "our loader differs from the reference's on its own range"
    full = our_loader(path, max_frames=None, fighters=names, fighter_to_player_id={"Pikachu": 0, "Joker": 1})
    stats[prec] = {"first600": run_stats(t_ref), "full": run_stats(full)}
    res[prec] = {"frames_full": len(full), "stats_sha256_first600": digest(stats[prec]["first600"]), "stats_sha256_full": digest(stats[prec]["full"]),
                 "actions_counted_full": {str(fid): int(sum(v.get("action_count", {}).values())) for fid, v in stats[prec]["full"].items()}}
# fp32 CPU oracle on the reference loader's range: frames 0..599 need frames up to 626 (window reach 27). The synthetic
# frames are integer-only torch ops, identical on CPU and GPU, so they are regenerated here (test infrastructure: this
# tool, like tests/, may run the oracle; the product never does).
if os.environ.get("CFG4_ORACLE", "1") == "1":
    import torch
    from oracle import ref_path
    from playaid_core_b200.fighter import yolo_pixels_batch
    from workloads import synthetic, weights

    n_or = 627
    frames = synthetic.synth_frames(np.arange(n_or), yolo_pixels_batch(boxes[:n_or], 1920, 1080), device="cpu", seed=99).numpy()
    model = ref_path.RefCNNActionDetector(ACTIONS, 7).eval()
    model.load_state_dict(weights.calibrated_state_dict(0))
    lab_or, logp_or, prob_or = ref_path.classify_clip(frames, boxes[:n_or], model)
    lab_or, prob_or = np.asarray(lab_or)[:600], np.asarray(prob_or)[:600]
    path = os.path.join(tmp, "ai_output_oracle.yaml")
    with open(path, "w") as f:
        yaml.dump(ai_output(lab_or, prob_or), f)
    stats["oracle"] = {"first600": run_stats(ref_loader(path))}
    res["oracle_fp32"] = {"stats_sha256_first600": digest(stats["oracle"]["first600"])}
    res["f16x2_labels_identical_to_fp32_oracle_first600"] = bool((z["label_f16x2"][:600] == lab_or).all())
    res["f16_label_disagreements_with_fp32_oracle_first600"] = int((z["label_f16"][:600] != lab_or).sum())
    # confidences are floats: compare the Stats dicts, which hold actions / counts / damage, not confidences
    res["stats_identical_f16x2_vs_fp32_oracle_first600"] = stats["f16x2"]["first600"] == stats["oracle"]["first600"]
la, lb = z["label_f16"], z["label_f16x2"]
res["labels_identical_first600"] = bool((la[:600] == lb[:600]).all())
res["label_disagreements_first600"] = int((la[:600] != lb[:600]).sum())
res["label_disagreements_full"] = int((la != lb).sum())
res["stats_identical_first600"] = stats["f16"]["first600"] == stats["f16x2"]["first600"]
res["stats_identical_full"] = stats["f16"]["full"] == stats["f16x2"]["full"]
# where the two Stats dicts differ, how far apart are the per-action counts?
diff = 0
for fid in stats["f16x2"]["full"]:
    a = stats["f16"]["full"][fid].get("action_count", {}); b = stats["f16x2"]["full"][fid].get("action_count", {})
    diff += sum(abs(a.get(k, 0) - b.get(k, 0)) for k in set(a) | set(b))
res["action_count_l1_distance_full"] = int(diff)
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res, indent=1))
