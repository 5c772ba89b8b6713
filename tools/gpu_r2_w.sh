#!/bin/bash
set -x
mkdir -p gpurun_out
for w in 4 8; do
  ANNCUR_REFINE_WARPS=$w timeout 300 python tools/step_probe.py --n 1000000 --b 64 --steps 200
  ANNCUR_REFINE_WARPS=$w timeout 300 python tools/step_probe.py --n 1000000 --b 1 --steps 200
  ANNCUR_REFINE_WARPS=$w timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2w_ll_b64_w$w.csv python tools/step_probe.py --n 1000000 --b 64 --steps 3 --warmup 2 > /dev/null 2>&1
  grep refine_topk gpurun_out/r2w_ll_b64_w$w.csv | tail -2 | cut -d, -f5,15 | cut -c1-120
done
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -2
