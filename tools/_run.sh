timeout 900 python -m pytest tests/test_gpu_protocols.py tests/test_gpu_fixedbase.py -x -q -m gpu 2>&1 | tail -3
export BP_PROBE_REPS=8 BP_PROBE_DISTINCT=64
for n in 1024 8192; do timeout 300 python tools/verify_probe.py $n 2>&1 | tail -1; done
LOCAL_WORLD_SIZE=4 timeout 300 python tools/verify_probe.py 1024 2>&1 | tail -1
