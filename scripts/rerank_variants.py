"""Re-rank kernel variants alone at the headline shape (50,000 x 3000, 4096 queries, k = 100): device ms per launch
(re-rank + order), each checked against the exact scan."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch
from morna_b200 import _lib, synth
lib = _lib.load()
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.gauss(N, D, "cuda")
s = MornaSearch(vectors=S, stats=(N, N, D))
q, rows = synth.queries(S, Q)
s.enable_tensor_path()
pick = torch.arange(0, Q, 61, device="cuda")
ref_ids, ref_d = s.exact_search_device(q[pick], K, allow_single=False)
need = lib.morna_knn_batched_workspace_bytes(N, Q, D, K)
ws = _lib.workspace(need, "cuda"); ov = torch.zeros(Q, dtype=torch.uint8, device="cuda"); st = torch.zeros(4, dtype=torch.int32, device="cuda")
oi = torch.empty((Q, K), dtype=torch.int32, device="cuda"); od = torch.empty((Q, K), dtype=torch.float64, device="cuda")
_lib.check(lib.morna_knn_batched_score(_lib.dev_ptr(s.hs), s.ld_h, _lib.dev_ptr(s.rho_max), N, D, 0, _lib.dev_ptr(q), Q, D, K,
           _lib.dev_ptr(ov), _lib.dev_ptr(st), _lib.dev_ptr(ws), ws.numel(), None, None, None, _lib.stream_ptr()), "score")
def rerank():
    _lib.check(lib.morna_knn_batched_rerank(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(oi), _lib.dev_ptr(od), _lib.dev_ptr(ov), _lib.dev_ptr(ws), ws.numel(), 0, _lib.stream_ptr()), "rerank")
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize(); time.sleep(0.2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for kern, pipe, rows_per, cap in ((1, 0, 8, 0), (1, 1, 8, 0), (1, 1, 4, 0), (1, 1, 16, 0), (1, 0, 4, 0), (1, 1, 4, 4), (1, 1, 4, 3), (0, 0, 8, 0)):
    lib.morna_debug_set_tuning(14, kern); lib.morna_debug_set_tuning(18, pipe); lib.morna_debug_set_tuning(5, rows_per)
    lib.morna_debug_set_tuning(13, cap); lib.morna_debug_set_tuning(6, 0)
    oi.zero_(); od.zero_()
    ms = timed(rerank)
    ok = torch.equal(oi[pick], ref_ids) and torch.equal(od[pick], ref_d)
    print("kernel=%s pipelined=%d rows/pass=%d ctas/sm cap=%d: %.3f ms ok=%s" % ("cta" if kern else "warp", pipe, rows_per, cap, ms, ok), flush=True)
