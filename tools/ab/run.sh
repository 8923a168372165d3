#!/bin/bash
# A/B of library builds on the GPU box: preprocess tests once, then the kernel spans of each build
timeout 600 python -m pytest tests/test_gpu_preprocess.py -x -q 2>&1 | tail -3
for v in "$@"; do
  PLAYAID_B200_LIB=$PWD/tools/ab/lib_$v.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fast16 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json;d=json.load(open('gpurun_out/ab_$v.json'));print('$v',round(d['value']),d['ms_per_step'],{k:round(x['ms_per_step'],4) for k,x in d['kernels'].items() if 'preprocess' in k})"
done
