#!/bin/bash
# per-kernel device times of the batch verifier (development aid); the last bp_rp_verify_batch call of verify_probe.py runs on
# the 16-bit tables.  Writes gpurun_out/verify_kernels.csv and prints the kernels of the last call (serialised, cold caches).
N=${1:-8192}
BP_PROBE_REPS=1 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none \
    -k regex:'k_rp_|k_sv_|k_reduce_scalars|k_sum_points' -c 4000 \
    --csv --log-file gpurun_out/verify_kernels.csv python tools/verify_probe.py $N > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(l for l in open("gpurun_out/verify_kernels.csv") if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
t=[(r[ki].split("(")[0], r[gi], float(r[vi].replace(",",""))) for r in rows[1:] if r[mi].startswith("gpu__time")]
# last call = everything after the last k_reduce_scalars whose grid belongs to chunk 0 ... simply: the last 4 chunks x 11 kernels
import os
last=t[-int(os.environ.get("BP_NCU_LAST","44")):]
tot={}
for k,g,v in last:
    tot[k]=tot.get(k,0)+v
    print("%-22s grid=%-18s %10.0f ns" % (k,g,v))
print("---- per kernel, summed over the call (us):")
for k,v in sorted(tot.items(), key=lambda kv:-kv[1]): print("%-22s %9.1f" % (k, v/1e3))
print("total %.1f us" % (sum(tot.values())/1e3))
PY
