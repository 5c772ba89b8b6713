#!/bin/bash
# round-2 GPU pass N (2+ GPUs): adaptive owned form over NCCL, c3 extra in the multi-GPU bench
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded_nccl.py -x -q > gpurun_out/r2n_pytest_${N}gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2n_pytest_${N}gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --extras c3 > gpurun_out/r2n_bench_${N}gpu.json 2> gpurun_out/r2n_bench_${N}gpu.err; echo "bench exit $?"
tail -5 gpurun_out/r2n_bench_${N}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2n_bench_${N}gpu.json"))
print(round(d["value"]), d["ms_per_step"], json.dumps(d["extras"], indent=1))
PY
