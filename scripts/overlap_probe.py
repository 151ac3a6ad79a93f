"""Power/clock/time of the scoring half, the re-rank half, both in sequence and both overlapped on two
streams (consecutive batches), at the headline shape.  Decides whether cross-batch pipelining pays."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pynvml as nv
from morna_b200.search import MornaSearch
from morna_b200 import _lib
lib = _lib.load()
nv.nvmlInit(); H = nv.nvmlDeviceGetHandleByIndex(0)

class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True); self.p = []; self.c = []; self.r = set(); self.stop_evt = threading.Event()
    def run(self):
        while not self.stop_evt.is_set():
            self.p.append(nv.nvmlDeviceGetPowerUsage(H) / 1000.0); self.c.append(nv.nvmlDeviceGetClockInfo(H, nv.NVML_CLOCK_SM))
            m = nv.nvmlDeviceGetCurrentClocksEventReasons(H)
            if m & 0x4: self.r.add("sw_power_cap")
            if m & 0x8: self.r.add("hw_slowdown")
            if m & 0x20: self.r.add("sw_thermal")
            if m & 0x40: self.r.add("hw_thermal")
            time.sleep(0.02)
    def done(self):
        self.stop_evt.set(); self.join()
        return "power avg %.0f W max %.0f W, sm clock median %d MHz, reasons %s" % (np.mean(self.p[2:]), np.max(self.p), np.median(self.c[2:]), sorted(self.r))

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
def score(sl, stream):
    _lib.check(lib.morna_knn_batched_score(_lib.dev_ptr(s.hs), s.ld_h, _lib.dev_ptr(s.rho_max), N, D, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(sl.ov), _lib.dev_ptr(sl.st), _lib.dev_ptr(sl.ws), sl.ws.numel(), None, None, None, _lib.stream_ptr(stream)), "score")
def rerank(sl, stream):
    _lib.check(lib.morna_knn_batched_rerank(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(sl.ids), _lib.dev_ptr(sl.d), _lib.dev_ptr(sl.ov), _lib.dev_ptr(sl.ws), sl.ws.numel(), 0, _lib.stream_ptr(stream)), "rerank")
A, B = torch.cuda.Stream(), torch.cuda.Stream()
cur = torch.cuda.current_stream()
for sl in slots:
    score(sl, cur); rerank(sl, cur)
torch.cuda.synchronize()
ref_ids, ref_d = s.exact_search_device(q[:256], K)
assert torch.equal(slots[0].ids[:256], ref_ids) and torch.equal(slots[1].d[:256], ref_d)

def run(name, body, seconds=2.0):
    torch.cuda.synchronize(); smp = Sampler(); smp.start()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(20): body(n); n += 1
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%-34s %.3f ms/batch | %s" % (name, 1e3 * dt / n, smp.done()))
    time.sleep(1.0)

run("score only", lambda i: score(slots[0], cur))
run("rerank only", lambda i: rerank(slots[0], cur))
run("score + rerank, one stream", lambda i: (score(slots[0], cur), rerank(slots[0], cur)))
def overlapped(i):
    sl = slots[i & 1]
    with torch.cuda.stream(A):
        A.wait_event(sl.ev)              # its previous re-rank has consumed the workspace
        score(sl, A); e = torch.cuda.Event(); e.record(A)
    with torch.cuda.stream(B):
        B.wait_event(e); rerank(sl, B); sl.ev.record(B)
run("overlapped across batches", overlapped)
torch.cuda.synchronize()
assert torch.equal(slots[0].ids[:256], ref_ids) and torch.equal(slots[1].ids[:256], ref_ids)
print("overlapped results still equal the exact scan")
