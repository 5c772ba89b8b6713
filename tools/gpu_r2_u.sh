#!/bin/bash
# round-2 GPU pass U: launch list of the default bench step (final revision) + ncu --set full of its MAIN launch
set -x
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_ll_bench_n1m.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --extras none > gpurun_out/r2u_ll_bench_n1m.out 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2u_ll_bench_n1m.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mn=hdr.index('Metric Name')
sel=[r for r in rows[1:] if r[mn]=='gpu__time_duration.sum']
names=[r[ki] for r in sel]
packs=[i for i,n in enumerate(names) if 'pack_queries' in n]
print(len(sel), packs[-6:])
# the step after the last-but-one pack_queries launch of the device-resident loop: print launches between two packs
for a,b in zip(packs[:-1], packs[1:]):
    if b-a==7:
        last=(a,b)
a,b=last
tot=sum(float(r[mv].replace(',',''))/1e3 for r in sel[a:b])
for r in sel[a:b]:
    t=float(r[mv].replace(',',''))/1e3
    print(f"{t:10.1f} us {100*t/tot:6.1f} %  {r[ki].split('(')[0][:80]}")
print(f"{tot:10.1f} us total")
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_score_topk -c 3 -f -o gpurun_out/r2u_ncu_main_n1m \
    python tools/step_probe.py --n 1000000 --steps 1 --warmup 0 > gpurun_out/r2u_ncu_main_n1m.out 2>&1
ls -la gpurun_out/r2u_*
