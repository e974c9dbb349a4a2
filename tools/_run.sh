timeout 900 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "host_operands_in_parts" 2>&1 | tail -5
