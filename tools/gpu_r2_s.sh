#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2s_pytest.log
timeout 300 python tools/adaptive_c3.py --solver incremental > gpurun_out/r2s_adaptive_c3.txt 2>&1; cat gpurun_out/r2s_adaptive_c3.txt
