import sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O
ch=O.Chain.from_urdf(open("/root/repo/optik_b200/data/panda.urdf").read(),"panda_link0","panda_link8")
P=O.twin_params()
def attempt(tgt,q0,variant,maxev=64):
    q=np.clip(q0,ch.lb,ch.ub); ev=O.twin_eval(ch,q,tgt,P); evals=1
    f=ev["f"]; lam=0.1; slow=0; nu=2.0; rej=0
    if f<1e-6: return True,evals,rej
    while evals<maxev:
        J=ev["Jr"]; r=ev["r"]; g=J@r
        pinned=((q<=ch.lb)&(g>0))|((q>=ch.ub)&(g<0)); m=(~pinned).astype(float)
        Jm=J*m[:,None]
        A=Jm.T@Jm+lam*np.eye(6)
        y=np.linalg.solve(A,r); dq=-Jm@y
        qt=np.clip(q+dq,ch.lb,ch.ub)
        evt=O.twin_eval(ch,qt,tgt,P); evals+=1
        ft=evt["f"]
        if ft<1e-6: return True,evals,rej
        if ft<f:
            df=f-ft
            if variant=="nielsen":
                d=qt-q; pred=f-np.sum((r+Jm.T@d)**2)
                rho=min(max(df/max(pred,1e-300),-10.0),10.0)
                lam=max(lam*max(1/3.,1-(2*rho-1)**3),1e-9); nu=2.0
            elif variant=="gain":
                d=qt-q; pred=f-np.sum((r+Jm.T@d)**2); rho=min(max(df/max(pred,1e-300),-10.0),10.0)
                lam = lam*0.1 if rho>0.75 else (lam*0.3 if rho>0.25 else lam*2)
                lam=max(lam,1e-9)
            else:
                lam=max(lam*0.3,1e-9)
            slow = slow+1 if df<0.1*f else 0
            q,ev,f=qt,evt,ft
            if df<1e-9 or slow>=2: return False,evals,rej
        else:
            rej+=1
            if variant=="nielsen": lam*=nu; nu*=2
            else: lam*=10
            if lam>1e6: return False,evals,rej
    return False,evals,rej
for variant in ["cur","nielsen","gain"]:
    S=[];E=[];R=[]
    rng=np.random.default_rng(3)
    for t in range(5):
        _,tgt=ch.fk(rng.uniform(ch.lb,ch.ub))
        for i in range(150):
            ok,e,rj=attempt(tgt,ch.restart_seed(1+i+1000*t),variant); S.append(ok);E.append(e);R.append(rj)
    S=np.array(S);E=np.array(E)
    print(variant,"succ %.3f evals %.2f e/s %.1f rejects/att %.2f"%(S.mean(),E.mean(),E.mean()/S.mean(),np.mean(R)))
