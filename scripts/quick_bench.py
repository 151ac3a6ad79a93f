"""Timings of the batched path at the headline shape: phases (single stream) and pipelined totals."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch, make_phase_events, PHASE_NAMES
from morna_b200 import _lib
lib = _lib.load()
g = torch.Generator(device='cuda'); g.manual_seed(1234)
S = torch.randn((50000, 3000), generator=g, device='cuda')
s = MornaSearch(vectors=S, stats=(50000, 50000, 3000))
rows = torch.randperm(50000)[:4096].cuda()
q = S[rows].double()
s.enable_tensor_path()
e_ids, e_d = s.exact_search_device(q[:256], 100)

def timed(reps=10, **kw):
    for i in range(3):
        ids, d = s.batched_search_device(q, 100, **kw)
    torch.cuda.synchronize()
    ok = bool(torch.equal(e_ids, ids[:256])) and bool(torch.equal(e_d, d[:256]))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        s.batched_search_device(q, 100, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, ok

events, arr = make_phase_events()
acc = [0.0] * 6
for i in range(5):
    s.batched_search_device(q, 100, phase_events=arr, pipelined=False)
    torch.cuda.synchronize()
    if i >= 2:
        for j in range(6):
            acc[j] += events[j].elapsed_time(events[j + 1]) / 3
print("phases (single stream): " + " ".join("%s %.3f" % (n, v) for n, v in zip(PHASE_NAMES, acc)), "| stats", s.last_stats)
print("single stream: %.3f ms equal=%s" % timed(pipelined=False))
for chunk in (512, 1024, 2048):
    lib.morna_debug_set_tuning(2, chunk)
    print("pipelined chunk=%d: %.3f ms equal=%s" % ((chunk,) + timed(pipelined=True)))
print("pipelined, no overflow check chunk=1024:", end=" ")
lib.morna_debug_set_tuning(2, 1024)
print("%.3f ms equal=%s" % timed(pipelined=True, check_overflow=False))
