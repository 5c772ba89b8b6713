#!/bin/bash
# round-2 GPU pass A (one GPU): parity suite, default bench, k = 1000 probe, ncu launch list + one full capture of MAIN
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
nproc >> gpurun_out/r2_gpu.txt; free -g >> gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
tail -15 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2_bench_n1m.json 2> gpurun_out/r2_bench_n1m.err; echo "bench exit $?"
tail -3 gpurun_out/r2_bench_n1m.err
timeout 120 python tools/step_probe.py --n 100000 --k 1000 --precision f32r --steps 20 > gpurun_out/r2_probe_k1000.txt 2>&1
timeout 120 python tools/step_probe.py --n 100000 --k 1000 --precision f32x3 --steps 20 >> gpurun_out/r2_probe_k1000.txt 2>&1
timeout 120 python tools/step_probe.py --n 100000 --k 100 --precision f32r --steps 50 >> gpurun_out/r2_probe_k1000.txt 2>&1
timeout 120 python tools/step_probe.py --n 1000000 --k 100 --b 64 --precision f32r --steps 50 >> gpurun_out/r2_probe_k1000.txt 2>&1
cat gpurun_out/r2_probe_k1000.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ll_bench_n1m.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --extras none > gpurun_out/r2_ll_bench_n1m.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_score_topk -c 3 -f -o gpurun_out/r2_ncu_main_n1m \
    python tools/step_probe.py --n 1000000 --steps 1 --warmup 0 > gpurun_out/r2_ncu_main_n1m.out 2>&1
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n1m.json 2> gpurun_out/r2_bench_ref_n1m.err
ls -la gpurun_out | tail -12
