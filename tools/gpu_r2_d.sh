#!/bin/bash
# round-2 GPU pass D (N GPUs): rank-budgeted exchange + step breakdown
N=${1:-4}
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded_nccl.py tests/test_gpu_peer_exchange.py tests/test_gpu_kernels.py -x -q -k "nccl or peer or budgeted or orthonormalize or rank_large" > gpurun_out/r2d_pytest_${N}gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest_${N}gpu.log
tail -5 gpurun_out/r2d_pytest_${N}gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N > gpurun_out/r2d_bench_n1m_${N}gpu.json 2> gpurun_out/r2d_bench_n1m_${N}gpu.err; echo "bench auto exit $?"
tail -4 gpurun_out/r2d_bench_n1m_${N}gpu.err
timeout 400 $TR bench.py --gpus $N --local-k full --no-extra --extras none > gpurun_out/r2d_bench_n1m_${N}gpu_fullk.json 2> gpurun_out/r2d_bench_n1m_${N}gpu_fullk.err; echo "bench full exit $?"
timeout 400 $TR bench.py --gpus $N --exchange nccl --no-extra --extras none > gpurun_out/r2d_bench_n1m_${N}gpu_nccl.json 2> gpurun_out/r2d_bench_n1m_${N}gpu_nccl.err; echo "bench nccl exit $?"
python - <<PY
import json
for f in ("", "_fullk", "_nccl"):
    try:
        d = json.load(open(f"gpurun_out/r2d_bench_n1m_${N}gpu{f}.json"))
        print(f or "_auto", round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), round(d["roofline"]["launch_ms"], 4), d.get("rank_budgeted_exchange"), d.get("step_breakdown_ms"), d.get("sharded_equals_single_gpu", {}).get("indices_equal"), (d.get("extras") or {}).get("c4", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
