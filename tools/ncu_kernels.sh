#!/bin/bash
# per-kernel device times of one MSM (development aid): tools/ncu_kernels.sh <lgn>
LGN=${1:-20}
ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none \
    -k regex:'k_reduce|k_window|k_combine|k_fixup|k_accumulate|k_digits|k_scatter|k_phi' -s 24 -c 12 \
    --csv --log-file gpurun_out/kernels_$LGN.csv python tools/msm_probe.py --lgn $LGN --iters 2 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(l for l in open("gpurun_out/kernels_$LGN.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size"); bi=h.index("Block Size")
for r in rows[1:]:
    if r[mi].startswith("gpu__time"): print("%-16s grid=%-22s block=%-14s %10s ns" % (r[ki].split("(")[0], r[gi], r[bi], r[vi]))
PY
