import sys, torch
sys.path.insert(0,'/root/repo')
from planedepth_b200 import functional as F
from planedepth_b200.boundary import decoder_tail
B,N,H,W=12,49,192,640
for mix in (False, True):
    lr=torch.randn(B,N,H,W,device='cuda',requires_grad=True); sr=torch.randn(B,N,H,W,device='cuda',requires_grad=True) if mix else None
    base=(300.0*(2.0/300.0)**(torch.arange(N,device='cuda')/(N-1.0))).reshape(1,N,1,1).repeat(B,1,1,1).requires_grad_(True)
    mask=torch.ones(B,N,H,W,device='cuda'); gl=torch.randn(B,N,H,W,device='cuda'); gd=torch.randn(B,1,H,W,device='cuda')
    def step():
        out=decoder_tail(lr,sr,mask,base.expand(B,N,H,W),mix)
        torch.autograd.grad([out['logits'],out['disp']],[lr,base],[gl,gd])
    for _ in range(3): step()
    torch.cuda.synchronize(); F.KERNEL_TIMELINE=[]
    for _ in range(5): step()
    torch.cuda.synchronize()
    per={}
    for n,s,e in F.KERNEL_TIMELINE: per.setdefault(n,[]).append(s.elapsed_time(e))
    F.KERNEL_TIMELINE=None
    print(mix,{k:sum(v)/len(v) for k,v in per.items()})
