#!/bin/bash
set -x
mkdir -p gpurun_out
out=gpurun_out/r2h_sample_variants.txt; : > $out
for cfg in "" "ANNCUR_SAMPLE_STRIDED_VIEW=1" "ANNCUR_SMAX_WIDE=0" "ANNCUR_SAMPLE_STRIDE=32" "ANNCUR_SAMPLE_STRIDE=32 ANNCUR_SMAX_WIDE=0" "ANNCUR_SAMPLE_STRIDE=64"; do
  for shape in "--n 1000000 --b 4096 --steps 40" "--n 1000000 --b 64 --steps 100" "--n 100000 --b 4096 --steps 100" "--n 125000 --b 4096 --k 41 --steps 100" "--n 1000000 --b 64 --steps 200 --graph" "--n 1000000 --b 1 --steps 200"; do
    echo "[$cfg] $shape" >> $out
    env $cfg timeout 120 python tools/step_probe.py $shape --precision f32r >> $out 2>&1
  done
done
cat $out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_full_size.py tests/test_gpu_kernels.py -x -q > gpurun_out/r2h_pytest.log 2>&1; tail -4 gpurun_out/r2h_pytest.log
