#!/bin/bash
# round-2 GPU pass V: what the driver runs at round end on one GPU (full GPU suite, smoke, default bench, reference arm)
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2v_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2v_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2v_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2v_bench_reference.json 2> gpurun_out/r2v_bench_reference.err; echo "ref exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2v_bench.json"))
print(round(d["value"]), d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["peak_kind"], d["roofline"]["step_frac"], d["cpu_baseline"]["value"], d["clocks"], d["gpu_launches"])
for k, v in d["extras"].items(): print(k, v.get("value"), v.get("ms_per_step"))
r = json.load(open("gpurun_out/r2v_bench_reference.json")); print({k: r[k] for k in ("impl", "value", "ms_per_step") if k in r})
PY
