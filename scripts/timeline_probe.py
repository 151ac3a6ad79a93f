"""Stream-level timeline of consecutive batches: score(i+1) on stream A while rerank(i) runs on stream B,
CUDA events at every phase boundary of both, all measured against one origin event.  Burst regime
(20 batches after a pause), re-rank CTAs per SM capped so that it fits beside the GEMM CTA."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from morna_b200.search import MornaSearch
from morna_b200 import _lib
lib = _lib.load()
g = torch.Generator(device='cuda'); g.manual_seed(1234)
N, D, Q, K = 50000, 3000, 4096, 100
S = torch.randn((N, D), generator=g, device='cuda')
s = MornaSearch(vectors=S, stats=(N, N, D))
rows = torch.randperm(N)[:Q].cuda()
q = S[rows].double()
s.enable_tensor_path()
need = lib.morna_knn_batched_workspace_bytes(N, Q, D, K)
class Slot:
    def __init__(self):
        self.ws = _lib.workspace(need, 'cuda'); self.ov = torch.zeros(Q, dtype=torch.uint8, device='cuda')
        self.st = torch.zeros(4, dtype=torch.int32, device='cuda')
        self.ids = torch.empty((Q, K), dtype=torch.int32, device='cuda'); self.d = torch.empty((Q, K), dtype=torch.float64, device='cuda')
        self.ev = torch.cuda.Event()
slots = [Slot(), Slot()]
def mk_events(n):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
    for e in ev: e.record()
    return ev
def score(sl, stream, arr=None):
    _lib.check(lib.morna_knn_batched_score(_lib.dev_ptr(s.hs), s.ld_h, _lib.dev_ptr(s.rho_max), N, D, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(sl.ov), _lib.dev_ptr(sl.st), _lib.dev_ptr(sl.ws), sl.ws.numel(), arr, None, None, None, _lib.stream_ptr(stream)), "score")
def rerank(sl, stream):
    _lib.check(lib.morna_knn_batched_rerank(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(sl.ids), _lib.dev_ptr(sl.d), _lib.dev_ptr(sl.ov), _lib.dev_ptr(sl.ws), sl.ws.numel(), 0, _lib.stream_ptr(stream)), "rerank")
cur = torch.cuda.current_stream()
for sl in slots:
    score(sl, cur); rerank(sl, cur)
torch.cuda.synchronize()
NB = 20
for cap, prio in ((0, False), (2, False), (2, True), (1, True), (3, True)):
    lib.morna_debug_set_tuning(13, cap); lib.morna_debug_set_tuning(15, cap)
    A = torch.cuda.Stream(priority=-1 if prio else 0); B = torch.cuda.Stream(priority=0)
    time.sleep(1.0)
    origin = torch.cuda.Event(enable_timing=True)
    sc_ev = [mk_events(7) for _ in range(NB)]
    rr_ev = [mk_events(2) for _ in range(NB)]
    torch.cuda.synchronize()
    origin.record(cur)
    A.wait_stream(cur); B.wait_stream(cur)
    for i in range(NB):
        sl = slots[i & 1]
        arr = (ctypes.c_void_p * 7)(*[e.cuda_event for e in sc_ev[i]])
        with torch.cuda.stream(A):
            A.wait_event(sl.ev)
            score(sl, A, arr); sc_ev[i][6].record(A)
            e = torch.cuda.Event(); e.record(A)
        with torch.cuda.stream(B):
            B.wait_event(e); rr_ev[i][0].record(B); rerank(sl, B); rr_ev[i][1].record(B); sl.ev.record(B)
    torch.cuda.synchronize()
    total = origin.elapsed_time(rr_ev[NB - 1][1])
    print("cap=%d priority=%s: %.3f ms per batch over %d batches" % (cap, prio, total / NB, NB))
    for i in (8, 9, 10):
        t = [origin.elapsed_time(e) for e in sc_ev[i][:6]] + [origin.elapsed_time(sc_ev[i][6])]
        r0, r1 = origin.elapsed_time(rr_ev[i][0]), origin.elapsed_time(rr_ev[i][1])
        print("  batch %2d score: start %.3f prep %.3f pilot %.3f kth %.3f filter %.3f kthf %.3f end %.3f (durations %s) | rerank %.3f -> %.3f (%.3f)"
              % (i, t[0], t[1], t[2], t[3], t[4], t[5], t[6], " ".join("%.3f" % (t[j + 1] - t[j]) for j in range(6)), r0, r1, r1 - r0))
lib.morna_debug_set_tuning(13, 0); lib.morna_debug_set_tuning(15, 0)
