"""Phase timings of the batched path at the headline shape, for each GEMM variant."""
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
variants = [(1, 6), (1, 4), (0, 4)] if len(sys.argv) < 2 else [tuple(int(x) for x in a.split(',')) for a in sys.argv[1:]]
for pair, stages in variants:
    lib.morna_debug_set_tuning(0, pair); lib.morna_debug_set_tuning(1, stages)
    events, arr = make_phase_events()
    for i in range(3):
        ids, d = s.batched_search_device(q, 100, phase_events=arr)
    torch.cuda.synchronize()
    ok = bool(torch.equal(e_ids, ids[:256])) and bool(torch.equal(e_d, d[:256]))
    acc = [0.0] * 6
    reps = 10
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for i in range(reps):
        e0.record()
        s.batched_search_device(q, 100, phase_events=arr)
        e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1) / reps
        for j in range(6):
            acc[j] += events[j].elapsed_time(events[j + 1]) / reps
    print("pair=%d stages=%d equal=%s total %.3f ms | " % (pair, stages, ok, tot) +
          " ".join("%s %.3f" % (n, v) for n, v in zip(PHASE_NAMES, acc)), "| stats", s.last_stats)
