import sys, numpy as np, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O
robots={"panda":("panda_link0","panda_link8"),"ur5":("base_link","ee_link"),"ur3e":("ur_base_link","ur_ee_link"),"snake20":("seg0","tip")}
def run(name, R=2000, ntgt=3, **lm):
    b,e=robots[name]; ch=O.Chain.from_urdf(open(f"optik_b200/data/{name}.urdf").read(),b,e)
    rng=np.random.default_rng(42)
    out=[]
    for t in range(ntgt):
        qs=rng.uniform(ch.lb,ch.ub); _,tgt=ch.fk(qs)
        x0=0.5*(ch.lb+ch.ub)
        P=O.twin_params(**lm)
        t0=time.time(); q,f,st,ev=O.twin_attempts(ch,tgt,x0,0,R,P); dt=time.time()-t0
        ok=st==1
        # verify with ref-style oracle
        bad=0
        for i in np.where(ok)[0][:50]:
            fo=ch.objective(q[i],tgt); 
            if not (fo<1e-6 and np.all(q[i]>=ch.lb) and np.all(q[i]<=ch.ub)): bad+=1
        out.append((ok.mean(), ev[ok].mean() if ok.any() else 0, ev[~ok].mean() if (~ok).any() else 0, ev.mean(), dt/R*1e6, bad, np.bincount(st,minlength=8)))
    for o in out: print(name, "succ %.3f ev_ok %.1f ev_fail %.1f ev_all %.1f us/att %.1f bad %d st %s"%o)
    s=np.mean([o[0] for o in out]); e=np.mean([o[3] for o in out])
    print(name,"=> succ %.3f evals/att %.1f  evals per success %.1f"%(s,e,e/max(s,1e-9)))
if __name__=="__main__":
    for name in sys.argv[1].split(","):
        run(name)
