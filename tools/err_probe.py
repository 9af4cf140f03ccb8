import sys, math; sys.path.insert(0,'/root/repo')
import torch, torch.nn.functional as F
from cofii2p_b200 import ops
def rel(a,b): return float((a.double().cpu()-b.double()).abs().max()/b.abs().max())
for (m,n,k) in [(1024,128,64),(1024,128,512),(300,1024,3072),(1024,512,7680)]:
    g = torch.Generator().manual_seed(k)
    a = torch.randn((m,k),generator=g); w = torch.randn((n,k),generator=g)/math.sqrt(k)
    ref = F.linear(a.double(), w.double())
    out = []
    for e in ('fp32','tf32','tf32x3'):
        ops.set_engine(e); out.append((e, rel(ops.gemm(a.cuda(), w.cuda()), ref)))
    out.append(('torch_cuda_fp32', rel(F.linear(a.cuda(), w.cuda()), ref)))
    print(m,n,k,out)
