#!/bin/bash
# Final ncu evidence of round 2 (run on the GPU box; outputs in gpurun_out/, summarised HERE by tools/ncu_extract.py + tools/ncu_sort_extract.py):
#   1. launch list of the bench command with the final code                     -> r2_launches_bench.csv
#   3. ncu --set full of k_accumulate_slots, the dominant kernel of that command   -> r2_accumulate_slots.ncu-rep
#   2. ncu --set full of the two sort kernels of the precomputed path (k_digits_pre, k_scatter_pre; 2^20 terms, c = 18)
#      -> r2_digits_pre.ncu-rep, r2_scatter_pre.ncu-rep: are they bound by L2 atomic / scattered-store throughput?
mkdir -p gpurun_out
export BP_BENCH_NO_SWEEP=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-verify --strong "" > gpurun_out/r2_launches_bench.log 2>&1
tail -c 300 gpurun_out/r2_launches_bench.log; echo
[ "$1" = "list" ] && exit 0      # (the three full captures are ~27 MB each; gpurun brings back at most 64 MiB per call: take them in a second call)
for k in k_digits_pre k_scatter_pre; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_${k#k_} \
      env BP_PRE_SLOTS=0 python tools/msm_probe.py --lgn 20 --c 16 --iters 2 --pre 0 > gpurun_out/r2_${k#k_}.log 2>&1
  tail -1 gpurun_out/r2_${k#k_}.log
done
# 3. the dominant kernel after the slot sort: k_accumulate_slots (third launch of the bench's precomputed path)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate_slots -s 3 -c 1 -f -o gpurun_out/r2_accumulate_slots \
    python bench.py --steps 2 --warmup 1 --no-verify --strong "" > gpurun_out/r2_accumulate_slots.log 2>&1
tail -c 200 gpurun_out/r2_accumulate_slots.log; echo
ls -la gpurun_out/r2_*.ncu-rep
