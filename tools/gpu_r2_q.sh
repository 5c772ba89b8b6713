#!/bin/bash
# round-2 GPU pass Q (N GPUs): the driver's scaling-bench command, default flags (extras c3 + c4 ride along)
N=${1:-8}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 900 $TR bench.py --gpus $N > gpurun_out/r2q_bench_n1m_${N}gpu.json 2> gpurun_out/r2q_bench_n1m_${N}gpu.err ); echo "bench exit $?"
tail -4 gpurun_out/r2q_bench_n1m_${N}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2q_bench_n1m_${N}gpu.json"))
print(round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d["roofline"]["frac"], d["roofline"]["peak_kind"], d["roofline"]["step_frac"], d.get("rank_budgeted_exchange"), d.get("step_breakdown_ms"), d.get("sharded_equals_single_gpu", {}).get("indices_equal"), d.get("pipelined_2_streams"), d["clocks"])
for k, v in (d.get("extras") or {}).items(): print(k, json.dumps(v)[:900])
PY
