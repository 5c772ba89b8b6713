#!/bin/bash
# round-2 GPU pass Z: ncu --set full of the final warp Cholesky and of a large fp64-operand GEMM launch (Schur complement, round 3)
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 600 $NCU -k "regex:chol_inv32_kernel" -s 12 -c 1 -o gpurun_out/r2z_ncu_chol32 python tools/adaptive_c3.py --b 1024 --reps 1 > gpurun_out/r2z_ncu_chol32.log 2>&1
timeout 600 $NCU -k "regex:dgemm_nt_kernel<\(bool\)0>" -s 52 -c 8 -o gpurun_out/r2z_ncu_dgemm0 python tools/adaptive_c3.py --b 1024 --reps 1 > gpurun_out/r2z_ncu_dgemm0.log 2>&1
ls -la gpurun_out/r2z_*.ncu-rep
