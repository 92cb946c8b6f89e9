"""Issue rates of the multiply pipes on this device (inf_measure_imad_peak kinds 0-6)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib  # noqa: E402

ctx = ib.get_context(0)
for kind, name in ((0, "IMAD.LO, IMAD/s"), (1, "IMAD.WIDE, 2 IMAD-eq each"), (2, "IMAD.WIDE carry chains, 2 IMAD-eq each"),
                   (3, "IMAD.HI"), (4, "DFMA alone, instr/s"), (5, "DFMA + IMAD.WIDE 1:1, instr/s"),
                   (6, "DFMA + IMAD.WIDE 2:1, instr/s")):
    v, clk = C.c_double(), C.c_double()
    ctx.check(ctx.lib.inf_measure_imad_peak(ctx.handle, kind, C.byref(v), C.byref(clk)))
    print("kind %d  %-42s %7.3f T/s   clock %.0f MHz" % (kind, name, v.value / 1e12, clk.value), flush=True)
