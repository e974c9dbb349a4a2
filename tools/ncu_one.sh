#!/bin/bash
# full ncu capture of one kernel during tools/verify_probe.py: tools/ncu_one.sh <kernel-regex> <out-name> [nproofs]
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${4:-1} -c 1 -o gpurun_out/$2 python tools/verify_probe.py ${3:-2048} > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log
