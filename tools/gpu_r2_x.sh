#!/bin/bash
# A/B on one box: back-off between the polls of the long mbarrier waits (ANNCUR_WAIT_SLEEP_NS), N = 1M and C2, interleaved
set -x
mkdir -p gpurun_out
out=gpurun_out/r2x_wait_sleep.txt; : > $out
for rep in 1 2; do
for ns in 0 20 100 400; do
  echo "[ANNCUR_WAIT_SLEEP_NS=$ns] rep $rep" >> $out
  ANNCUR_WAIT_SLEEP_NS=$ns timeout 300 python tools/step_probe.py --n 1000000 --b 4096 --steps 150 >> $out 2>&1
  ANNCUR_WAIT_SLEEP_NS=$ns timeout 300 python tools/step_probe.py --n 100000 --b 4096 --steps 300 >> $out 2>&1
done
done
cat $out
