#!/bin/bash
# per-kernel device times of the batch verifier (development aid); the third verify_packed call runs on tables
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_rp_|k_fb_lookup|k_fb_fold|k_reduce_unit_plain|k_combine_plain' -s 16 -c 28 \
    --csv --log-file gpurun_out/verify_kernels.csv python tools/verify_probe.py 8192 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(l for l in open("gpurun_out/verify_kernels.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
for r in rows[1:]:
    print("%-22s grid=%-18s %10s ns" % (r[ki].split("(")[0], r[gi], r[vi]))
PY
