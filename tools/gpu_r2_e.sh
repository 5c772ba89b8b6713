#!/bin/bash
# round-2 GPU pass E (8 GPUs): the driver's scaling-bench command at N = 8 (+ the NCCL form for comparison)
N=${1:-8}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N > gpurun_out/r2e_bench_n1m_${N}gpu.json 2> gpurun_out/r2e_bench_n1m_${N}gpu.err; echo "bench auto exit $?"
tail -4 gpurun_out/r2e_bench_n1m_${N}gpu.err
timeout 400 $TR bench.py --gpus $N --local-k full --no-extra --extras none > gpurun_out/r2e_bench_n1m_${N}gpu_fullk.json 2> gpurun_out/r2e_bench_n1m_${N}gpu_fullk.err; echo "bench full exit $?"
timeout 400 $TR bench.py --gpus $N --exchange nccl --no-extra --extras none > gpurun_out/r2e_bench_n1m_${N}gpu_nccl.json 2> gpurun_out/r2e_bench_n1m_${N}gpu_nccl.err; echo "bench nccl exit $?"
python - <<PY
import json
for f in ("", "_fullk", "_nccl"):
    try:
        d = json.load(open(f"gpurun_out/r2e_bench_n1m_${N}gpu{f}.json"))
        print(f or "_auto", round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), round(d["roofline"]["launch_ms"], 4), d.get("rank_budgeted_exchange"), d.get("step_breakdown_ms"), d.get("sharded_equals_single_gpu", {}).get("indices_equal"), (d.get("extras") or {}).get("c4", {}).get("value"), d.get("pipelined_2_streams"))
    except Exception as e:
        print(f, "failed", e)
PY
