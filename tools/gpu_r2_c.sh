#!/bin/bash
# round-2 GPU pass C (one GPU): full parity suite incl. the new rows, per-shard step sizes, the C5 sweep + rank tool
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest_gpu.log
tail -15 gpurun_out/r2c_pytest_gpu.log
for n in 125000 250000 500000; do
  timeout 120 python tools/step_probe.py --n $n --k 100 --precision f32r --steps 50 >> gpurun_out/r2c_probe_shards.txt 2>&1
done
cat gpurun_out/r2c_probe_shards.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c_ll_n125k.csv \
    python tools/step_probe.py --n 125000 --steps 2 --warmup 2 > gpurun_out/r2c_ll_n125k.out 2>&1
timeout 900 python -m anncur_b200.run_sweep_eval --synthetic 10000 1000000 --res_dir gpurun_out/r2c_sweep_c5 --methods cur \
    --k_i 50 100 200 500 1000 2000 --rank > gpurun_out/r2c_sweep_c5.log 2>&1; echo "sweep exit $?"
tail -8 gpurun_out/r2c_sweep_c5.log
nvidia-smi --query-gpu=memory.used --format=csv
