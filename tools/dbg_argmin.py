import sys, torch
sys.path.insert(0,'/root/repo')
from pgtformer_b200 import ops
g=torch.Generator().manual_seed(1)
cb=torch.randn(1025,512,generator=g).cuda()
import os
for regime in ('random','near'):
    T=49152
    if regime=='random': z=torch.randn(T,512,generator=g).cuda()
    else:
        pick=torch.randint(0,1024,(T,),generator=g).cuda()
        z=(cb[pick]+0.05*torch.randn(T,512,generator=g).cuda()).contiguous()
    pack=ops.codebook_pack(cb,1024)
    idx=torch.empty(T,dtype=torch.int64,device='cuda')
    q=torch.empty(T,512,device='cuda')
    print(regime, file=sys.stderr)
    for i in range(2):
        ops.l2_argmin_tc(z,cb,pack,1024,idx,None)
        torch.cuda.synchronize()
