"""One scoring call and a few re-rank calls at the headline shape (for ncu captures of the re-rank kernels).
usage: rerank_only.py [kernel 0|1] [rows] [phase_mb]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch
from morna_b200 import _lib, synth
lib = _lib.load()
kern = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rows_per = int(sys.argv[2]) if len(sys.argv) > 2 else 8
phase_mb = int(sys.argv[3]) if len(sys.argv) > 3 else -1
if len(sys.argv) > 4:
    lib.morna_debug_set_tuning(34, int(sys.argv[4])); lib.morna_debug_set_tuning(20, int(sys.argv[5]) if len(sys.argv) > 5 else 1)
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.gauss(N, D, "cuda")
s = MornaSearch(vectors=S, stats=(N, N, D))
q, rows = synth.queries(S, Q)
s.enable_tensor_path()
lib.morna_debug_set_tuning(14, kern); lib.morna_debug_set_tuning(5, rows_per); lib.morna_debug_set_tuning(6, phase_mb)
need = lib.morna_knn_batched_workspace_bytes(N, Q, D, K)
ws = _lib.workspace(need, "cuda"); ov = torch.zeros(Q, dtype=torch.uint8, device="cuda"); st = torch.zeros(4, dtype=torch.int32, device="cuda")
oi = torch.empty((Q, K), dtype=torch.int32, device="cuda"); od = torch.empty((Q, K), dtype=torch.float64, device="cuda")
_lib.check(lib.morna_knn_batched_score(_lib.dev_ptr(s.hs), s.ld_h, _lib.dev_ptr(s.rho_max), N, D, 0, _lib.dev_ptr(q), Q, D, K,
           _lib.dev_ptr(ov), _lib.dev_ptr(st), _lib.dev_ptr(ws), ws.numel(), None, None, None, _lib.stream_ptr()), "score")
for _ in range(3):
    _lib.check(lib.morna_knn_batched_rerank(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(oi), _lib.dev_ptr(od), _lib.dev_ptr(ov), _lib.dev_ptr(ws), ws.numel(), 0, _lib.stream_ptr()), "rerank")
torch.cuda.synchronize()
assert int(oi[0, 0]) == int(rows[0])
print("ok")
