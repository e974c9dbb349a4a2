#!/bin/bash
# ncu evidence for the bench line (run on the GPU box; outputs land in gpurun_out/, summaries are made by tools/ncu_extract.py HERE):
#   1. launch list (gpu__time_duration per kernel) of `python bench.py --steps 2 --warmup 1 --no-verify`  -> r2_launches_bench.csv
#   2. ncu --set full of the dominant MSM kernel, k_accumulate, on the precomputed path of that same command -> r2_accumulate.ncu-rep
#   3. ncu --set full of the dominant batch-verifier kernel, k_rp_lookup16 (tools/verify_probe.py)          -> r2_lookup16.ncu-rep
# (a number printed by a run under ncu is never a bench value; the verify leg is left out of 1. because proving 8192 proofs
#  under ncu's per-launch serialisation takes hours -- its kernel list is tools/ncu_verify.sh)
mkdir -p gpurun_out
export BP_BENCH_NO_SWEEP=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-verify --strong "" > gpurun_out/r2_launches_bench.log 2>&1
tail -c 300 gpurun_out/r2_launches_bench.log; echo
# k_accumulate launches of that command: 8 on the plain path (3 warm-up + 5), then the precomputed path: take its second
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 9 -c 1 -f -o gpurun_out/r2_accumulate \
    python bench.py --steps 2 --warmup 1 --no-verify --strong "" > gpurun_out/r2_accumulate.log 2>&1
tail -2 gpurun_out/r2_accumulate.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rp_lookup16 -s 8 -c 1 -f -o gpurun_out/r2_lookup16 \
    python tools/verify_probe.py 8192 > gpurun_out/r2_lookup16.log 2>&1
tail -2 gpurun_out/r2_lookup16.log
ls -la gpurun_out/r2_*.ncu-rep
