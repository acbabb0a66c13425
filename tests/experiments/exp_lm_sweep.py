import sys, numpy as np, time, itertools
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O
robots={"panda":("panda_link0","panda_link8"),"ur5":("base_link","ee_link"),"ur3e":("ur_base_link","ur_ee_link"),"snake20":("seg0","tip")}
def run(name, R=400, ntgt=12, **lm):
    b,e=robots[name]; ch=O.Chain.from_urdf(open(f"optik_b200/data/{name}.urdf").read(),b,e)
    rng=np.random.default_rng(7)
    S=[];E=[]
    for t in range(ntgt):
        qs=rng.uniform(ch.lb,ch.ub); _,tgt=ch.fk(qs)
        x0=0.5*(ch.lb+ch.ub)
        P=O.twin_params(**lm)
        q,f,st,ev=O.twin_attempts(ch,tgt,x0,1,R+1,P)
        S.append((st==1).mean()); E.append(ev.mean())
    s=np.mean(S); e=np.mean(E)
    return s,e,e/max(s,1e-9)
if __name__=="__main__":
    names=sys.argv[1].split(",")
    for l0,dec,inc in itertools.product([1e-3,1e-2,1e-1,1.0],[0.1,0.3],[4.0,10.0]):
        res=[run(n,lambda0=l0,lambda_dec=dec,lambda_inc=inc) for n in names]
        print("l0 %.0e dec %.1f inc %4.1f | "%(l0,dec,inc)+" | ".join("%s s %.3f e %.1f e/s %.1f"%((n,)+r) for n,r in zip(names,res)))
