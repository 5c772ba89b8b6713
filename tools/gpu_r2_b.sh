#!/bin/bash
# round-2 GPU pass B (N GPUs of one box, default 2): NCCL / peer-memory equality test, the item-sharded bench in its three exchange forms
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharded_nccl.py tests/test_gpu_peer_exchange.py -x -q > gpurun_out/r2_pytest_nccl_${N}gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_nccl_${N}gpu.log
tail -5 gpurun_out/r2_pytest_nccl_${N}gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N > gpurun_out/r2_bench_n1m_${N}gpu.json 2> gpurun_out/r2_bench_n1m_${N}gpu.err; echo "bench p2p exit $?"
tail -5 gpurun_out/r2_bench_n1m_${N}gpu.err
timeout 400 $TR bench.py --gpus $N --exchange nccl --steps 100 --no-extra > gpurun_out/r2_bench_n1m_${N}gpu_nccl.json 2> gpurun_out/r2_bench_n1m_${N}gpu_nccl.err; echo "bench nccl exit $?"
timeout 400 $TR bench.py --gpus $N --exchange allgather --steps 100 --no-extra > gpurun_out/r2_bench_n1m_${N}gpu_allgather.json 2> gpurun_out/r2_bench_n1m_${N}gpu_allgather.err; echo "bench allgather exit $?"
tail -3 gpurun_out/r2_bench_n1m_${N}gpu_nccl.err gpurun_out/r2_bench_n1m_${N}gpu_allgather.err
python - <<PY
import json
for f in ("", "_nccl", "_allgather"):
    try:
        d = json.load(open(f"gpurun_out/r2_bench_n1m_${N}gpu{f}.json"))
        print(f or "_p2p", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["launch_ms"], d.get("sharded_equals_single_gpu"), d.get("pipelined_2_streams"), (d.get("extras") or {}).get("c4", {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
