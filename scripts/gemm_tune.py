"""GEMM variant / stage sweep at the headline shape (filter-pass time from the phase events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch, make_phase_events, PHASE_NAMES
from morna_b200 import _lib
lib = _lib.load()
N, Q = 50000, 4096
g = torch.Generator(device='cuda'); g.manual_seed(1234)
S = torch.randn((N, 3000), generator=g, device='cuda')
s = MornaSearch(vectors=S, stats=(N, N, 3000))
q = S[torch.randperm(N)[:Q].cuda()].double()
s.enable_tensor_path()
for pair, stages in ((1, 4), (1, 6), (0, 4)):
    lib.morna_debug_set_tuning(0, pair); lib.morna_debug_set_tuning(1, stages)
    events, arr = make_phase_events()
    acc = [0.0] * 6
    for _ in range(3):
        s.batched_search_device(q, 100, phase_events=arr)
    torch.cuda.synchronize()
    for _ in range(8):
        s.batched_search_device(q, 100, phase_events=arr); torch.cuda.synchronize()
        for j in range(6): acc[j] += events[j].elapsed_time(events[j + 1]) / 8
    tf = 2.0 * Q * (N - 8192) * 3000 / (acc[3] * 1e-3) / 1e12
    print("pair=%d stages=%d: pilot %.3f filter %.3f ms (%.0f TF/s)" % (pair, stages, acc[1], acc[3], tf), flush=True)
lib.morna_debug_set_tuning(0, 1); lib.morna_debug_set_tuning(1, 4)
