#!/bin/bash
# round-2 GPU pass R: per-row table (r2), ncu --set full of the DMMA Gram kernel, the warp Cholesky and the refine kernel at N = 1M
set -x
mkdir -p gpurun_out
timeout 900 python tools/bench_rows.py > gpurun_out/r2r_rows_c2.jsonl 2> gpurun_out/r2r_rows_c2.err; echo "rows exit $?"; tail -3 gpurun_out/r2r_rows_c2.err; cut -c1-220 gpurun_out/r2r_rows_c2.jsonl
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 600 $NCU -k "regex:dgemm_nt_kernel<\(bool\)1>" -s 5 -c 1 -o gpurun_out/r2r_ncu_gram python tools/adaptive_c3.py --b 1024 --reps 1 > gpurun_out/r2r_ncu_gram.log 2>&1; tail -2 gpurun_out/r2r_ncu_gram.log | cut -c1-200
timeout 600 $NCU -k "regex:chol_inv32_kernel" -s 12 -c 1 -o gpurun_out/r2r_ncu_chol32 python tools/adaptive_c3.py --b 1024 --reps 1 > gpurun_out/r2r_ncu_chol32.log 2>&1; tail -2 gpurun_out/r2r_ncu_chol32.log | cut -c1-200
timeout 600 $NCU -k "regex:backsolve_e_kernel" -s 8 -c 1 -o gpurun_out/r2r_ncu_backsolve python tools/adaptive_c3.py --b 1024 --reps 1 > gpurun_out/r2r_ncu_backsolve.log 2>&1; tail -2 gpurun_out/r2r_ncu_backsolve.log | cut -c1-200
timeout 600 $NCU -k "regex:refine_topk_kernel" -s 3 -c 1 -o gpurun_out/r2r_ncu_refine_n1m python tools/step_probe.py --n 1000000 --b 4096 --steps 2 --warmup 2 > gpurun_out/r2r_ncu_refine.log 2>&1; tail -2 gpurun_out/r2r_ncu_refine.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
