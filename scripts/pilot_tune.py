"""Pilot size / first refinement point sweep at the headline shape, checked against the exact scan."""
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
rows = torch.randperm(N)[:Q].cuda()
q = S[rows].double()
s.enable_tensor_path()
e_ids, e_d = s.exact_search_device(q[:128], 100)
for pilot, first in ((8192, 0), (4096, 0), (2048, 0), (4096, 16384), (2048, 8192), (2048, 16384), (1024, 8192), (1024, 4096)):
    lib.morna_debug_set_tuning(10, pilot); lib.morna_debug_set_tuning(11, first)
    events, arr = make_phase_events()
    acc = [0.0] * 6
    for _ in range(3):
        ids, d = s.batched_search_device(q, 100, phase_events=arr)
    torch.cuda.synchronize()
    for _ in range(5):
        ids, d = s.batched_search_device(q, 100, phase_events=arr); torch.cuda.synchronize()
        for j in range(6): acc[j] += events[j].elapsed_time(events[j + 1]) / 5
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): s.batched_search_device(q, 100)
    e1.record(); torch.cuda.synchronize()
    ok = bool(torch.equal(e_ids, ids[:128])) and bool(torch.equal(e_d, d[:128]))
    print("pilot=%d first=%d: %s | sum %.3f, loop %.3f ms/batch, stats %s equal=%s" % (
        pilot, first, " ".join("%s %.3f" % (n, v) for n, v in zip(PHASE_NAMES, acc)), sum(acc), e0.elapsed_time(e1) / 20, s.last_stats, ok), flush=True)
