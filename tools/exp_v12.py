import sys; sys.path.insert(0,"tools"); sys.path.insert(0,".")
import exp_r2
for v in (1, 2, 1, 2):
    exp_r2.batch("panda", 1 << 20, 32, variant=v)
    exp_r2.batch("ur5", 1 << 20, 32, variant=v)
    exp_r2.batch("panda", 1 << 18, 32, mode="quality", variant=v)
    exp_r2.batch("panda", 1 << 18, 32, variant=v)
