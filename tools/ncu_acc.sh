#!/bin/bash
# ncu --set full of k_accumulate: plain path (2^20, c = 16) and precomputed path (window $1, default 19); reports land in gpurun_out/
C=${1:-19}
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 2 -c 1 -f -o gpurun_out/acc_plain python tools/msm_probe.py --lgn 20 --iters 1 > gpurun_out/ncu_acc.log 2>&1; tail -5 gpurun_out/ncu_acc.log
ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 6 -c 1 -f -o gpurun_out/acc_pre$C python tools/msm_probe.py --lgn 20 --iters 1 --pre $C > gpurun_out/ncu_acc.log 2>&1; tail -5 gpurun_out/ncu_acc.log
ls -la gpurun_out/*.ncu-rep
