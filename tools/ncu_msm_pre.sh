#!/bin/bash
# launch list (device time per kernel) of one precomputed-path MSM at 2^20: the last 13 launches of the probe
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/msm_pre_launches.csv python tools/msm_probe.py --lgn ${2:-20} --iters 1 --c 16 --pre ${1:-19} > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(l for l in open("gpurun_out/msm_pre_launches.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
t=[(r[ki].split("(")[0], r[gi], float(r[vi].replace(",",""))) for r in rows[1:]]
last=t[-15:]
for k,g,v in last: print("%-22s grid=%-18s %9.1f us" % (k,g,v/1e3))
print("sum %.1f us" % (sum(v for _,_,v in last)/1e3))
PY
