import os, sys, time, ctypes, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0], "8192"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "verify_probe.py")).read().split("# raw C call timing")[0])
for c in (4, 5, 6, 7, 8):
    lib.bp_msm_set_window(c)
    best = 1e9
    for it in range(3):
        t = time.perf_counter(); acc = verify_packed(batch, g1, h1, gs, hs, u1); best = min(best, time.perf_counter() - t)
    assert acc == b"\x01" * total
    lib.bp_msm_set_profiling(1); verify_packed(batch, g1, h1, gs, hs, u1); lib.bp_msm_stage_ms(st); lib.bp_msm_set_profiling(0)
    print("c=%d  %.2f ms   stages(last chunk) %s" % (c, best * 1e3, ["%.2f" % x for x in st]), flush=True)
lib.bp_msm_set_window(0)
