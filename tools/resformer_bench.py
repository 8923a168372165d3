"""Throughput of the second detector (ResnetTransformerDetector) on one B200, with the CPU oracle timed
beside it: B windows of 7 crops per call (crops resident in HBM as pa_preprocess writes them). Writes
gpurun_out/resformer_bench.json.   python tools/resformer_bench.py [B]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_resformer import RefResnetTransformerDetector   # checker / CPU baseline only
from playaid_core_b200 import _lib
from playaid_core_b200.anim_ontology import ACTIONS
from playaid_core_b200.models.resnet_transformer_detector import ResnetTransformerDetector

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = 7
torch.manual_seed(0)
oracle = RefResnetTransformerDetector(ACTIONS, sequence_length=S).eval()
res = {"windows_per_call": B, "crops_per_call": B * S, "gflop_per_crop_resnet50_128": 2.69}
x = torch.rand((B, S, 3, 128, 128), generator=torch.Generator().manual_seed(1))
for prec in ("f16", "f16x2"):
    m = ResnetTransformerDetector(ACTIONS, sequence_length=S, precision=prec).eval().load_state_dict(oracle.state_dict())
    x4 = torch.zeros((B * S, 128, 136, 4), dtype=torch.float32, device="cuda")
    x4[:, :, 4:132, :3] = x.cuda().reshape(B * S, 3, 128, 128).permute(0, 2, 3, 1)
    hi = x4.to(m.act_dtype)
    crops = torch.stack([hi, (x4 - hi.float()).to(m.act_dtype)]).contiguous() if m.split else hi.contiguous()
    for _ in range(3):
        y = m.forward_crops(crops, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for _ in range(K):
        y = m.forward_crops(crops, B)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    ctx = _lib.Context.get(torch.device("cuda", 0))
    ctx.profile_begin()
    for _ in range(4):
        m.forward_crops(crops, B)
    prof = ctx.profile_end()
    groups = {}
    for name, (n, tms) in prof.items():
        g = "encoder GEMMs" if ("transformer" in name or "classifier" in name or "resnet_ffn" in name) else \
            ("ResNet-50 convs" if name.startswith("conv") else name)
        groups[g] = groups.get(g, 0.0) + tms / 4
    res[prec] = {"ms_per_call": ms, "windows_per_s": B / ms * 1e3, "crops_per_s": B * S / ms * 1e3,
                 "resnet50_tflops": B * S * 2.69e9 / (groups.get("ResNet-50 convs", ms) / 1e3) / 1e12,
                 "ms_by_group": {k: round(v, 4) for k, v in sorted(groups.items(), key=lambda kv: -kv[1])},
                 "top_kernels_ms": {k: round(v[1] / 4, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:14]}}
    if prec == "f16x2":
        yg = y.cpu().numpy()
torch.set_num_threads(os.cpu_count() or 1)
nb = min(B, 8)
with torch.no_grad():
    oracle(x[:1])
    t0 = time.perf_counter()
    ref = oracle(x[:nb]).numpy()
    dt = time.perf_counter() - t0
res["cpu_oracle"] = {"windows_per_s": nb / dt, "cores": os.cpu_count(), "sample": f"{nb} windows ({nb * S} crops) in {dt:.2f} s"}
with torch.no_grad():
    full = oracle(x).numpy() if B <= 64 else None
if full is not None:
    res["f16x2_rel_err_vs_oracle"] = float(np.abs(yg - full).max() / np.abs(full).max())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/resformer_bench.json", "w"), indent=1)
print(json.dumps(res, indent=1))
