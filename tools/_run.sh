timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/proto_probe.py 2>&1 | tail -16
