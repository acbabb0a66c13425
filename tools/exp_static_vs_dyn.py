"""Mid-size Speed batches: dynamic chains vs the static (target, chunk) schedule with automatic chunk count."""
import sys; sys.path.insert(0,"tools"); sys.path.insert(0,".")
import exp_r2
for T in (5000, 8192, 16384, 30000, 65536):
    for R in (8, 32):
        exp_r2.batch("panda", T, R)
        exp_r2.batch("panda", T, R, static=True)
        for C in (2, 4, 8):
            if C <= R: exp_r2.batch("panda", T, R, static=True, chunks=C)
