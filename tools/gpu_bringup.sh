#!/bin/bash
# Runs each GPU test file in its own process (a faulting kernel poisons the CUDA context) with a
# hard timeout, logging to gpurun_out/. Usage on the GPU box: bash tools/gpu_bringup.sh [files...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
FILES=${@:-"tests/test_gpu_preprocess.py tests/test_gpu_layers.py tests/test_gpu_model.py"}
for f in $FILES; do
  name=$(basename $f .py)
  echo "=== $f" | tee -a gpurun_out/summary.txt
  timeout 600 python -m pytest $f -q -m gpu --timeout 300 -s > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
done
