"""Probe: msm_probe.py with the (disabled by default) window-pipelined single-MSM path off and on.  Run by hand on a GPU box."""
import runpy
import sys


def main():
    sys.path.insert(0, ".")
    from python_bulletproofs_b200 import _native as nat
    for pm in (0, 1 << 17):
        nat.load().bp_msm_set_pipeline_min(pm)
        print("pipeline_min", pm, flush=True)
        sys.argv = ["x", "--lgn", "18,20"]
        runpy.run_path("tools/msm_probe.py", run_name="__main__")


if __name__ == "__main__":
    main()
