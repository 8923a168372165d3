#!/bin/bash
# PA_CONV_DEBUG experiments (results are wrong, timings are the point): 1 = no epilogue, 2 = (almost) no TMA fill, 3 = both.
for d in ${@:-0 1 2 3}; do
  PA_CONV_DEBUG=$d python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/conv_debug_print.py $d
done
