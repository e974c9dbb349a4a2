import sys; sys.path.insert(0, ".")
from python_bulletproofs_b200 import _native as nat
import runpy
for pm in (0, 1 << 17):
    nat.load().bp_msm_set_pipeline_min(pm)
    print("pipeline_min", pm, flush=True)
    sys.argv = ["x", "--lgn", "18,20"]
    runpy.run_path("tools/msm_probe.py", run_name="__main__")
