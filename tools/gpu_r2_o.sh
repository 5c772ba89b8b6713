#!/bin/bash
# round-2 GPU pass O: small-batch launch list (N = 1M, B = 64), fp64 yardstick, ncu --set full of the DMMA Gram kernel
set -x
mkdir -p gpurun_out
timeout 300 python tools/step_probe.py --n 1000000 --b 64 --steps 100 > gpurun_out/r2o_b64.txt 2>&1; cat gpurun_out/r2o_b64.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2o_ll_b64.csv python tools/step_probe.py --n 1000000 --b 64 --steps 3 --warmup 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2o_ll_b64.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mn=hdr.index('Metric Name'); gi=hdr.index('Grid Size'); bi=hdr.index('Block Size')
sel=[r for r in rows[1:] if r[mn]=='gpu__time_duration.sum']
for r in sel[-16:]:
    print(f"{float(r[mv].replace(',',''))/1e3:8.1f} us {r[gi]:>14} {r[bi]:>12}  {r[ki].split('(')[0][:80]}")
PY
timeout 300 python - > gpurun_out/r2o_fp64_yardstick.txt 2>&1 <<'PY'
import torch
a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda"); b = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
for _ in range(3): a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): a @ b
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"cuBLAS fp64 GEMM 4096^3 (torch.matmul): {ms:.3f} ms = {2 * 4096 ** 3 / ms / 1e9:.1f} TFLOP/s")
PY
cat gpurun_out/r2o_fp64_yardstick.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:dgemm_nt_kernel<1>" -s 5 -c 1 -f -o gpurun_out/r2o_ncu_gram python tools/adaptive_c3.py --b 1024 --reps 1 > gpurun_out/r2o_ncu_gram.log 2>&1; tail -3 gpurun_out/r2o_ncu_gram.log
ls -la gpurun_out/r2o_ncu_gram.ncu-rep
