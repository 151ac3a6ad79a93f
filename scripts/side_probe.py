"""What the helper warps do to one scoring call: score alone, score carrying the previous batch's re-rank as a side
job, and the resume call that drains what the helpers left (its time ~ the share of items left)."""
import sys, os, time, ctypes
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
need = lib.morna_knn_batched_workspace_bytes(N, Q, D, K)
class Slot:
    def __init__(self):
        self.ws = _lib.workspace(need, 'cuda'); self.ov = torch.zeros(Q, dtype=torch.uint8, device='cuda')
        self.st = torch.zeros(4, dtype=torch.int32, device='cuda')
        self.ids = torch.empty((Q, K), dtype=torch.int32, device='cuda'); self.d = torch.empty((Q, K), dtype=torch.float64, device='cuda')
        self.job = _lib.RerankJob(s.vectors.data_ptr(), s.pp.data_ptr(), N, D, s.ld, 0, q.data_ptr(), Q, D, K, self.ov.data_ptr(), self.ws.data_ptr(), self.ws.numel())
a, b = Slot(), Slot()
def score(sl, side=None):
    _lib.check(lib.morna_knn_batched_score(_lib.dev_ptr(s.hs), s.ld_h, _lib.dev_ptr(s.rho_max), N, D, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(sl.ov), _lib.dev_ptr(sl.st), _lib.dev_ptr(sl.ws), sl.ws.numel(), None,
               ctypes.byref(side.job) if side is not None else None, None, _lib.stream_ptr()), "score")
def rerank(sl, resume):
    _lib.check(lib.morna_knn_batched_rerank(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(sl.ids), _lib.dev_ptr(sl.d), _lib.dev_ptr(sl.ov), _lib.dev_ptr(sl.ws), sl.ws.numel(), resume, _lib.stream_ptr()), "rerank")
ref_ids, ref_d = s.exact_search_device(q[:128], K, allow_single=False)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for phase_mb in (-1, 0):
    lib.morna_debug_set_tuning(6, phase_mb)
    for trial in range(3):
        score(a); torch.cuda.synchronize(); time.sleep(0.3)
        ev[0].record(); score(b); ev[1].record(); torch.cuda.synchronize()
        t_plain = ev[0].elapsed_time(ev[1])
        score(a); torch.cuda.synchronize(); time.sleep(0.3)        # fresh queue in a
        ev[0].record(); score(b, side=a); ev[1].record(); rerank(a, 1); ev[2].record(); rerank(b, 0); ev[3].record(); torch.cuda.synchronize()
        ok = torch.equal(a.ids[:128], ref_ids) and torch.equal(a.d[:128], ref_d) and torch.equal(b.ids[:128], ref_ids)
        print("phase_mb=%d: score alone %.3f ms | score + side job %.3f ms, resume (drain + order) %.3f ms, full re-rank + order %.3f ms ok=%s"
              % (phase_mb, t_plain, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ok), flush=True)
