N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N > gpurun_out/r2_bench_${N}gpu_d.json 2> gpurun_out/r2_bench_${N}gpu_d.err; tail -c 150 gpurun_out/r2_bench_${N}gpu_d.err; head -c 150 gpurun_out/r2_bench_${N}gpu_d.json
