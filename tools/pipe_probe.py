"""Integer-pipe microbenchmarks on the current GPU (roofline denominators)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from python_bulletproofs_b200 import _native as nat   # noqa: E402

nat.init(0)
lib = nat.load()
names = {0: "IMAD.WIDE.U32 Rd,Ra,Rb,RZ (32x32->64 product)", 1: "IMAD (32-bit mad.lo + accumulate)",
         2: "IMAD.WIDE.U32(.X) carry-chained rows (as in fp_mul)", 3: "fp_mul (field multiplications)",
         4: "32-bit add (IADD3 + IMAD.IADD, both pipes)", 5: "fp_sqr (field squarings)",
         10: "DFMA (fma.rz.f64) alone", 11: "DFMA interleaved 1:1 with IMAD.WIDE (counted: DFMA only)"}
for mode in (0, 1, 2, 4, 10, 11, 3, 5):
    for iters in ((1 << 14), (1 << 16)) if mode not in (3, 5) else ((1 << 9), (1 << 11)):
        ops, ms = ctypes.c_double(), ctypes.c_float()
        nat.check(lib.bp_pipe_probe(mode, iters, ctypes.byref(ops), ctypes.byref(ms)))
        per_clk_sm = ops.value / 148 / 1.965e9
        print("mode %d %-52s iters=%6d  %.2f ms  %.3f Tops/s  (%.1f /clk/SM at 1965 MHz)" % (mode, names[mode], iters, ms.value, ops.value / 1e12, per_clk_sm), flush=True)

# latency probes: one warp, dependent chain; ms / iters = latency of one operation
lat = {6: "fp_mul (dependent chain)", 7: "coop_dbl (4-lane doubling, 3 levels)", 8: "xyzz_dbl (one thread, 9 mults)", 9: "coop_add (4-lane addition, 4 levels)"}
for mode in (6, 7, 8, 9):
    iters = 4096
    ops, ms = ctypes.c_double(), ctypes.c_float()
    nat.check(lib.bp_pipe_probe(mode, iters, ctypes.byref(ops), ctypes.byref(ms)))
    print("mode %d %-44s %.3f us per op (%.0f cycles at 1965 MHz)" % (mode, lat[mode], ms.value * 1e3 / iters, ms.value * 1e-3 / iters * 1.965e9), flush=True)
