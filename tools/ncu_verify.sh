#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_rp_expand|k_reduce_unit_plain|k_combine_plain|k_rp_accept' -c 8 \
    --csv --log-file gpurun_out/verify_kernels.csv python tools/verify_probe.py 2048 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(l for l in open("gpurun_out/verify_kernels.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
for r in rows[1:]:
    print("%-22s grid=%-18s %10s ns" % (r[ki].split("(")[0], r[gi], r[vi]))
PY
