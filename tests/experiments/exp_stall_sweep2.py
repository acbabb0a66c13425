import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O
robots={"panda":("panda_link0","panda_link8"),"ur5":("base_link","ee_link"),"ur3e":("ur_base_link","ur_ee_link")}
chains={k:O.Chain.from_urdf(open(f"/root/repo/optik_b200/data/{k}.urdf").read(),b,e) for k,(b,e) in robots.items()}
def run(name, **lm):
    ch=chains[name]; rng=np.random.default_rng(7); S=[];E=[];MX=[]
    for t in range(10):
        _,tgt=ch.fk(rng.uniform(ch.lb,ch.ub)); x0=0.5*(ch.lb+ch.ub)
        q,f,st,ev=O.twin_attempts(ch,tgt,x0,1,601,O.twin_params(layout=1,**lm))
        S.append((st==1).mean()); E.append(ev.mean()); MX.append(np.percentile(ev,99.5))
    s=np.mean(S); e=np.mean(E)
    return s,e,e/s,np.mean(MX)
for rel,cnt,dec,inc,l0 in [(1e-2,3,0.3,10,0.1),(5e-2,2,0.3,10,0.1),(1e-1,2,0.3,10,0.1),(2e-1,2,0.3,10,0.1),(3e-1,2,0.3,10,0.1),(1e-1,1,0.3,10,0.1),(2e-1,3,0.3,10,0.1),(1e-1,2,0.2,10,0.1),(1e-1,2,0.3,5,0.1),(1e-1,2,0.3,10,0.03)]:
    res=[run(n,stall_rel=rel,stall_count=cnt,lambda_dec=dec,lambda_inc=inc,lambda0=l0) for n in robots]
    print(f"rel {rel:.0e} cnt {cnt} dec {dec} inc {inc} l0 {l0} | "+" | ".join("%s s %.3f e %.1f e/s %.1f p99.5 %.0f"%((n,)+r) for n,r in zip(robots,res)))
