#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2t_pytest.log
timeout 300 python tools/adaptive_c3.py --solver incremental 2>&1 | tail -1
timeout 300 python tools/step_probe.py --n 100000 --b 4096 --k 250 --steps 50
timeout 300 python tools/step_probe.py --n 1000000 --b 4096 --k 250 --steps 20
