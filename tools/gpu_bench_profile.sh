#!/bin/bash
# bench + ncu evidence on the GPU box (one ncu invocation per call).
#   bash tools/gpu_bench_profile.sh <tag> bench                       benches (all precisions + reference arm) + ncu launch list
#   bash tools/gpu_bench_profile.sh <tag> full [regex] [skip] [count] one `ncu --set full` capture of the top kernels
TAG=${1:-r01}
MODE=${2:-bench}
KREGEX=${3:-"conv_gemm|conv_patch|conv1_kernel|preprocess_tc_kernel|preprocess_plan"}
SKIP=${4:-30}
COUNT=${5:-14}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fast16"
if [ "$MODE" = "full" ]; then
  $CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s $SKIP -c $COUNT \
      -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
  echo "ncu full exit $?"
  exit 0
fi
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -n 5 gpurun_out/bench_$TAG.err
python bench.py --steps 20 --warmup 3 --precision f16 --no-cpu-baseline > gpurun_out/bench_${TAG}_f16.json 2> gpurun_out/bench_${TAG}_f16.err
echo "bench f16 exit $?"
python bench.py --steps 10 --warmup 3 --weights default --no-cpu-baseline --no-fast16 > gpurun_out/bench_${TAG}_default_init.json 2> gpurun_out/bench_${TAG}_default_init.err
echo "bench default-init exit $?"
python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
echo "reference exit $?"
python - <<PY
import json
for n in ["bench_$TAG","bench_${TAG}_f16","bench_${TAG}_default_init","bench_${TAG}_reference"]:
    try:
        d=json.load(open("gpurun_out/"+n+".json"))
        print(n, "value=%.1f ms/step=%.3f e2e=%.1f launches=%s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d.get("gpu_launches")))
        if d.get("roofline"): print("   tensor frac %.3f  preprocess hbm frac %.4f"%(d["roofline"]["frac"], d["roofline_preprocess"]["frac"]))
        ks=d.get("kernels") or {}
        for k,v in sorted(ks.items(), key=lambda kv:-kv[1]["ms_per_step"])[:8]: print("   %-40s %.3f ms %.1f%%"%(k,v["ms_per_step"],100*v["share"]))
    except Exception as e: print(n, "ERR", e)
PY
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'preprocess|conv|maxpool|avgpool|head|split|stage|boxes' -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
