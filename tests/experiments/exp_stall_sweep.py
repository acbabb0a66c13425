import sys, numpy as np, itertools
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from exp2 import run
names=sys.argv[1].split(",")
for rel,cnt in [(0,1000),(1e-3,2),(1e-3,3),(1e-2,2),(1e-2,3),(5e-2,2),(5e-2,3),(1e-1,3),(1e-1,4),(2e-1,4)]:
    res=[run(n,lambda0=0.1,lambda_dec=0.3,lambda_inc=10.0,stall_rel=rel,stall_count=cnt) for n in names]
    print("rel %.0e cnt %d | "%(rel,cnt)+" | ".join("%s s %.3f e %.1f e/s %.1f"%((n,)+r) for n,r in zip(names,res)))
