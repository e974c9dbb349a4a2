bash tools/ncu_metrics.sh 2>&1 | tail -12
timeout 900 python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 400 gpurun_out/r2_bench_b.err; head -c 600 gpurun_out/r2_bench_b.json
