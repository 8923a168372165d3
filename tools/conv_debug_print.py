import json, sys
d = json.loads(sys.stdin.read()); k = d["kernels"]
keep = ("layer1.0", "layer2.1", "layer3.1", "layer4.1", "fc", "conv1_stem")
print("debug", sys.argv[1], "step %.3f" % d["ms_per_step"],
      " ".join("%s=%.3f" % (n.split(":")[-1], v["ms_per_step"]) for n, v in k.items() if any(s in n for s in keep)))
