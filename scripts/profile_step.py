"""One headline-shape batched search step, repeated a few times, for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch
from morna_b200 import _lib
lib = _lib.load()
if len(sys.argv) > 2:
    lib.morna_debug_set_tuning(0, int(sys.argv[1])); lib.morna_debug_set_tuning(1, int(sys.argv[2]))
g = torch.Generator(device='cuda'); g.manual_seed(1234)
S = torch.randn((50000, 3000), generator=g, device='cuda')
s = MornaSearch(vectors=S, stats=(50000, 50000, 3000))
rows = torch.randperm(50000)[:4096].cuda()
q = S[rows].double()
s.enable_tensor_path()
for i in range(4):
    ids, d = s.batched_search_device(q, 100)
torch.cuda.synchronize()
assert bool((ids[:, 0].long() == rows).all())
print("ok", s.last_stats)
