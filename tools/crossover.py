import sys; sys.path.insert(0,"."); sys.path.insert(0,"tests")
import numpy as np, torch, cases
from chebpol_b200 import _lib
shapes=[(20,20),(30,30),(64,64),(10,10,10),(16,16,16),(8,)*4,(10,)*4,(12,)*4,(20,)*3,(8,)*5,(32,32,32),(24,24,24),(10,)*5,(16,)*4,(6,)*6,(20,)*4,(20,20,20,4),(5,)*7,(12,)*5,(7,)*6]
for dims in shapes:
    S=0; p_=1
    for n_ in reversed(dims): p_*=n_; S+=p_
    cf=cases.cheb_stress(dims); n=int(min(8e6, max(2e5, 4e11/S)))
    x=torch.rand((n,len(dims)),device="cuda",dtype=torch.float64)*2-1; out=torch.empty(n,device="cuda",dtype=torch.float64)
    res={}
    for v in (100000,1):
        p=_lib.Plan.cheb(cf); p.set(_lib.OPT_VARIANT, v)
        for _ in range(2): p.eval_torch(x,out)
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); [p.eval_torch(x,out) for _ in range(3)]; e1.record(); torch.cuda.synchronize()
        res[p.kernel_used]=3*n/(e0.elapsed_time(e1)*1e-3); p.close()
    print(dims,"S=%d"%S, {k:"%.3e"%v for k,v in res.items()}, "mma/slab=%.2f"%(res.get("cheb_mma",0)/res.get("cheb_slab",1)))
