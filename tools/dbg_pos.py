import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from _inputs import pos_inputs
from oracle import fa_oracle
from dualsuperreslearningforsemseg_b200.models.losses import FALoss
def rn(a,b): return np.linalg.norm(a-b)/np.linalg.norm(b)
for (s1,s2,k,red) in [((1,32,16,8),(1,32,16,8),1,"mean"), ((2,64,32,32),(2,64,32,32),1,"mean"), ((1,256,16,32),(1,256,16,32),1,"mean")]:
    x1,x2=pos_inputs(s1,s2,54321)
    ol,o1,o2=fa_oracle.fa_position(x1,x2,k,red)
    a=torch.from_numpy(x1).cuda(); b=torch.from_numpy(x2).cuda()
    l=FALoss(subsample_factor=k,reduction=red,affinity='position')(a,b); torch.cuda.synchronize()
    print(s1,'fwd-only loss',float(l),ol,abs(float(l)-ol)/ol, flush=True)
    a.requires_grad_(True); b.requires_grad_(True)
    l=FALoss(subsample_factor=k,reduction=red,affinity='position')(a,b); l.backward(); torch.cuda.synchronize()
    print(s1,'grad loss',float(l),ol,abs(float(l)-ol)/ol,'g1',rn(a.grad.cpu().numpy(),o1),'g2',rn(b.grad.cpu().numpy(),o2), flush=True)
