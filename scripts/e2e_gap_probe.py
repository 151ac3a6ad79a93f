"""Streaming API: HBM-resident queries vs pinned host queries (float32 / float64), depth 2 and 3."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch
N, Q, K = 50000, 4096, 100
g = torch.Generator(device='cuda'); g.manual_seed(1234)
S = torch.randn((N, 3000), generator=g, device='cuda')
s = MornaSearch(vectors=S, stats=(N, N, 3000))
rows = torch.randperm(N)[:Q].cuda()
q_dev64 = S[rows].double()
q_host32 = S[rows].cpu().pin_memory()
q_host64 = S[rows].double().cpu().pin_memory()
s.enable_tensor_path()
for depth in (2, 3):
    for name, src in (("resident f64", q_dev64), ("pinned host f32", q_host32), ("pinned host f64", q_host64)):
        for _ in s.search_batches((src for _ in range(6)), K, depth=depth): pass
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            for ids, d in s.search_batches((src for _ in range(40)), K, depth=depth): pass
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / 40 * 1e3)
        print("depth=%d %-16s %.3f ms per batch (wall clock, best of 3 runs of 40)" % (depth, name, best), flush=True)
