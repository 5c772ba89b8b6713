#!/bin/bash
# round-2 GPU pass P: the driver's commands at N = 1 (smoke, default bench, reference arm), timed
set -x
mkdir -p gpurun_out
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2p_smoke.log 2>&1; tail -5 gpurun_out/r2p_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err ); echo "bench exit $?"; tail -3 gpurun_out/r2p_bench.err
( time timeout 600 python bench.py --impl reference > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err ); echo "ref exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2p_bench.json"))
print(round(d["value"]), d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["step_frac"], d["cpu_baseline"]["value"], d["clocks"])
for k, v in d["extras"].items(): print(k, json.dumps(v)[:700])
r = json.load(open("gpurun_out/r2p_bench_reference.json")); print({k: r[k] for k in ("impl", "value", "ms_per_step") if k in r})
PY
