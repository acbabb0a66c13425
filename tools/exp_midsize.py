import sys; sys.path.insert(0,"tools"); sys.path.insert(0,".")
import optik_b200 as ob
ob.LIB_PATH=sys.argv[1]
import exp_r2
for T in (5000, 8192, 16384, 30000):
    for R in (2, 32):
        exp_r2.batch("panda", T, R)
exp_r2.batch("panda", 1<<16, 32)
exp_r2.batch("ur3e", 1<<14, 100)
