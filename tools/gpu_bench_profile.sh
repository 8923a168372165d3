#!/bin/bash
# bench + ncu evidence on the GPU box. Usage: bash tools/gpu_bench_profile.sh <round-tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench_$TAG.json; tail -n 5 gpurun_out/bench_$TAG.err
python bench.py --steps 10 --warmup 3 --precision bf16x2 --no-cpu-baseline > gpurun_out/bench_${TAG}_x2.json 2> gpurun_out/bench_${TAG}_x2.err
echo "bench x2 exit $?"; tail -c 600 gpurun_out/bench_${TAG}_x2.json
python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err
echo "reference exit $?"; cat gpurun_out/bench_${TAG}_reference.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'preprocess|conv|maxpool|avgpool|head|split' -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_gemm|conv1_kernel|preprocess' -s 30 -c 14 \
    -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | tail -20
