#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2l_pytest.log
timeout 300 python tools/adaptive_c3.py --solver incremental > gpurun_out/r2l_adaptive_c3.txt 2>&1
cat gpurun_out/r2l_adaptive_c3.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2l_ll_adaptive.csv python tools/adaptive_c3.py --b 1024 --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2l_ll_adaptive.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mn=hdr.index('Metric Name')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    if r[mn]!='gpu__time_duration.sum': continue
    n=r[ki].split('(')[0][:60]; tot[n]+=float(r[mv].replace(',',''))/1e6; cnt[n]+=1
for n,t in tot.most_common(12): print(f"{t:9.3f} ms  x{cnt[n]:4d}  {n}")
PY
