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
         4: "32-bit add (IADD3 + IMAD.IADD, both pipes)", 5: "fp_sqr (field squarings)"}
for mode in (0, 1, 2, 4, 3, 5):
    for iters in ((1 << 14), (1 << 16)) if mode not in (3, 5) else ((1 << 9), (1 << 11)):
        ops, ms = ctypes.c_double(), ctypes.c_float()
        nat.check(lib.bp_pipe_probe(mode, iters, ctypes.byref(ops), ctypes.byref(ms)))
        per_clk_sm = ops.value / 148 / 1.965e9
        print("mode %d %-52s iters=%6d  %.2f ms  %.3f Tops/s  (%.1f /clk/SM at 1965 MHz)" % (mode, names[mode], iters, ms.value, ops.value / 1e12, per_clk_sm), flush=True)
