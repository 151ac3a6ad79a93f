import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
from morna_b200.search import MornaSearch
g = torch.Generator(device='cuda'); g.manual_seed(1234)
S = torch.randn((50000, 3000), generator=g, device='cuda')
s = MornaSearch(vectors=S, stats=(50000, 50000, 3000))
rows = torch.randperm(50000)[:4096].cuda()
host_q = S[rows].cpu().pin_memory()
s.enable_tensor_path()
for rep in range(4):
    t0 = time.perf_counter(); ts = []
    for ids, d in s.search_batches((host_q for _ in range(10)), 100, depth=2):
        ts.append(time.perf_counter() - t0)
    print("run %d: total %.1f ms; per-yield ms: %s" % (rep, 1e3 * (time.perf_counter() - t0), " ".join("%.1f" % (1e3 * t) for t in ts)), flush=True)
