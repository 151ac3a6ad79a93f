import sys, time; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from morna_b200.search import MornaSearch
from morna_b200 import _lib
g=torch.Generator(device='cuda'); g.manual_seed(1234)
S=torch.randn((50000,3000),generator=g,device='cuda')
s=MornaSearch(vectors=S,stats=(50000,50000,3000))
rows=torch.randperm(50000)[:4096].cuda()
q=S[rows].double()
s.enable_tensor_path()
for i in range(3):
    ids,d=s.batched_search_device(q,100)
torch.cuda.synchronize()
print('stats',s.last_stats, 'self-first', bool((ids[:,0].long()==rows).all()))
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5): s.batched_search_device(q,100)
e1.record(); torch.cuda.synchronize()
print('batched ms/step', e0.elapsed_time(e1)/5)
e_ids,e_d=s.exact_search_device(q[:256],100)
print('equal to exact (256 q):', bool(torch.equal(e_ids,ids[:256])), bool(torch.equal(e_d,d[:256])))
