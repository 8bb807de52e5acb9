import os, sys, torch
sys.path.insert(0, '/root/repo')
from inverse_flow_b200 import _native, functional as IF
from inverse_flow_b200.stack import reference_init_weight
from tools.microbench import time_op
for (B,C,H,W,k) in [(64,48,32,32,3),(64,12,64,64,3)]:
    x = torch.randn(B,C,H,W,device='cuda'); w = reference_init_weight(C,k).cuda(); prep = IF.Prepared(w,1); out = torch.empty_like(x)
    for cfg in ["12,12","8,12","6,12","4,12","3,12","12,6","8,6","6,6","4,6","12,3","8,3","6,18","4,24","4,18","6,9","12,9","3,24","2,24"]:
        os.environ["IFK_STREAM_CFG"]=cfg
        d=_native.describe_solve(_native.problem(B,C,H,W,k,k,C,1))
        if not d.startswith("stream<cc=%s,nv=%s>"%tuple(cfg.split(','))): continue
        t=time_op(lambda: IF.inverse(x,w,out=out,prepared=prep), 3, reps=5)
        print((B,C,H,W,k), "%8.1f us"%t, d)
