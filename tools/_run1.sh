python -m pytest tests/test_gpu_protocols.py -x -q -m gpu 2>&1 | tail -4
for i in 1 2; do BP_PROBE_REPS=9 BP_PROBE_DISTINCT=256 python tools/verify_probe.py 8192 2>&1 | grep "over"; done
BP_PROBE_REPS=9 BP_PROBE_DISTINCT=256 python tools/verify_probe.py 1024 2>&1 | grep "over"
