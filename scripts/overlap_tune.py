"""Does capping the re-rank's resident CTAs per SM let it share the SMs with the next batch's GEMM?
Streams HBM-resident batches through the 2-deep pipeline and times them."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch
from morna_b200 import _lib
lib = _lib.load()
N, Q, K = 50000, 4096, 100
g = torch.Generator(device='cuda'); g.manual_seed(1234)
S = torch.randn((N, 3000), generator=g, device='cuda')
s = MornaSearch(vectors=S, stats=(N, N, 3000))
rows = torch.randperm(N)[:Q].cuda()
q = S[rows].double()
s.enable_tensor_path()
for depth in (2, 3):
    for cap in (0, 5, 4, 3, 2, 1):
        lib.morna_debug_set_tuning(13, cap)
        for _ in s.search_batches((q for _ in range(6)), K, depth=depth):
            pass
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 0
        for ids, d in s.search_batches((q for _ in range(40)), K, depth=depth):
            n += 1
        for sl in s._pipes[(K, depth)].slots:
            torch.cuda.current_stream().wait_stream(sl.stream)
        e1.record(); torch.cuda.synchronize()
        ok = int(ids[0, 0]) == int(rows[0])
        print("depth=%d re-rank CTAs/SM cap=%d: %.3f ms per batch ok=%s" % (depth, cap, e0.elapsed_time(e1) / n, ok), flush=True)
lib.morna_debug_set_tuning(13, 0)
