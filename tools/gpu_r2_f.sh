#!/bin/bash
# round-2 GPU pass F (one GPU): driver-style default run incl. the c3 extra, smoke(), bound usage at large K, compute-sanitizer probe
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2f_bench_n1m.json 2> gpurun_out/r2f_bench_n1m.err; echo "bench exit $?"
tail -3 gpurun_out/r2f_bench_n1m.err
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_n1m.json')); print(d['value'], d['e2e']['value'], d['roofline']['step_frac'], d['roofline']['traffic']); print(json.dumps(d['extras'])[:1500])"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2f_smoke.log
timeout 600 python -m pytest tests/test_gpu_fused.py -q -s -k "large_k" > gpurun_out/r2f_bound_usage.log 2>&1; grep "bound used" gpurun_out/r2f_bound_usage.log
which compute-sanitizer; ANNCUR_WAIT_TIMEOUT_CYCLES=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fused.py -q -x -k "f32r_matches_oracle or f32x3_matches_oracle" > gpurun_out/r2f_sanitizer_memcheck.log 2>&1; echo "sanitizer exit $?"; tail -12 gpurun_out/r2f_sanitizer_memcheck.log
