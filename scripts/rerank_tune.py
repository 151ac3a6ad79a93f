"""Re-rank tuning at the headline shape: rows per warp pass x phase size, checked against the exact scan.
Also times the other shapes the re-rank meets (few queries; many rows) so a knob that helps one
does not hurt the others."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch, make_phase_events, PHASE_NAMES
from morna_b200 import _lib
lib = _lib.load()


def phases(s, q, k, reps=5):
    events, arr = make_phase_events()
    acc = [0.0] * 6
    for _ in range(2):
        ids, d = s.batched_search_device(q, k, phase_events=arr)
    torch.cuda.synchronize()
    for _ in range(reps):
        ids, d = s.batched_search_device(q, k, phase_events=arr)
        torch.cuda.synchronize()
        for j in range(6):
            acc[j] += events[j].elapsed_time(events[j + 1]) / reps
    return acc, ids, d


for (N, Q) in ((50000, 4096), (125000, 4096), (50000, 512)):
    g = torch.Generator(device='cuda'); g.manual_seed(1234)
    S = torch.randn((N, 3000), generator=g, device='cuda')
    s = MornaSearch(vectors=S, stats=(N, N, 3000))
    rows = torch.randperm(N)[:Q].cuda()
    q = S[rows].double()
    s.enable_tensor_path()
    e_ids, e_d = s.exact_search_device(q[:128], 100)
    for rows_per_warp in (2, 4, 8):
        for phase_mb in (0, 32, 48, 64, 96):
            lib.morna_debug_set_tuning(5, rows_per_warp)
            lib.morna_debug_set_tuning(6, phase_mb)
            acc, ids, d = phases(s, q, 100)
            ok = bool(torch.equal(e_ids, ids[:128])) and bool(torch.equal(e_d, d[:128]))
            print("N=%d Q=%d rows/warp=%d phase_mb=%3d: rerank %.3f ms, step %.3f ms, equal=%s"
                  % (N, Q, rows_per_warp, phase_mb, acc[5], sum(acc), ok), flush=True)
    del s, S
lib.morna_debug_set_tuning(5, 4)
lib.morna_debug_set_tuning(6, 64)
