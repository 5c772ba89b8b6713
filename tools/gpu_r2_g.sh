#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dropin.py tests/test_data_formats.py -x -q -m gpu > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2g_pytest.log
timeout 300 python - > gpurun_out/r2g_pinv_timing.txt 2>&1 <<'PY'
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.getcwd())
from anncur_b200 import engine
from oracle import cur_oracle as O
def t_ms(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (m, n) in [(2000, 500), (500, 2000), (5000, 2000), (2000, 2000), (500, 125)]:
    A = torch.from_numpy(O.synthetic_scores(m, n, rank=64, noise=0.05, seed=1)).cuda()
    if m == n: A = A + 3 * torch.eye(m, device="cuda")
    fast = t_ms(lambda: engine.pinv(A))
    jac = t_ms(lambda: engine.pinv(A, return_cond=True), reps=2)
    t0 = time.perf_counter(); np.linalg.pinv(A.cpu().numpy()); cpu = (time.perf_counter() - t0) * 1e3
    print(f"pinv {m}x{n}: cholesky route {fast:.3f} ms (incl. the host read-back of the status), jacobi route {jac:.2f} ms, numpy fp32 on the host {cpu:.1f} ms")
PY
cat gpurun_out/r2g_pinv_timing.txt
timeout 600 python bench.py --no-cpu --extras c2 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json')); print(d['value'], d['index_build_s'], d['extras']['c2']['index_build_s'])"
timeout 300 python tools/dense_probe.py > gpurun_out/r2g_dense_c2.txt 2>&1; cat gpurun_out/r2g_dense_c2.txt
timeout 600 python tools/dense_probe.py --n 1000000 --b 10000 > gpurun_out/r2g_dense_c5.txt 2>&1; cat gpurun_out/r2g_dense_c5.txt
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2g_pytest2.log 2>&1; tail -4 gpurun_out/r2g_pytest2.log
