#!/bin/bash
# multi-GPU lines on an N-GPU box (gpurun --gpus N -- 'bash tools/gpu_scale.sh N tag [cfg5]'): the weak-scaling bench line,
# then (optionally) cfg5 -- 64 matches dealt over the ranks, one label gather, rank-0 recheck.
N=${1:-2}; TAG=${2:-r02}; CFG5=${3:-}
mkdir -p gpurun_out
PORT=29541
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
echo "bench $N gpus exit $?"; tail -n 3 gpurun_out/bench_${TAG}_${N}gpu.err
if [ -n "$CFG5" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) bench.py --gpus $N --workload cfg5 \
      > gpurun_out/cfg5_${TAG}_${N}gpu.json 2> gpurun_out/cfg5_${TAG}_${N}gpu.err
  echo "cfg5 $N gpus exit $?"; tail -n 3 gpurun_out/cfg5_${TAG}_${N}gpu.err
fi
python - <<PY
import json
for n in ["bench_${TAG}_${N}gpu", "cfg5_${TAG}_${N}gpu"]:
    try:
        d = json.load(open("gpurun_out/" + n + ".json"))
        print(n, "value=%.1f ms/step=%s e2e=%s by_rank=%s" % (d["value"], d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("ms_per_step_by_rank")))
        if "cfg5" in n: print("   ", {k: d[k] for k in d if k in ("config", "mismatches", "checked")})
    except Exception as e:
        print(n, "ERR", e)
PY
