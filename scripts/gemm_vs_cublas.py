"""The filter pass's contraction (4096 x 41808 x 3008, fp16 -> fp32 accumulate) through cuBLAS (torch.matmul, fp16
output) under the same timing as bench.py's phase events, next to knn_gemm2_kernel's filter pass."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch, make_phase_events
from morna_b200 import _lib
N, Q, D = 50000, 4096, 3000
g = torch.Generator(device='cuda'); g.manual_seed(1234)
S = torch.randn((N, D), generator=g, device='cuda')
s = MornaSearch(vectors=S, stats=(N, N, D))
q = S[torch.randperm(N)[:Q].cuda()].double()
s.enable_tensor_path()
A = torch.randn((Q, 3008), device='cuda', dtype=torch.float16)
B = s.hs[8192:]                                   # the filter pass's rows, fp16 [41808 x 3008]
flops = 2.0 * Q * B.shape[0] * 3008

def timed(fn, reps, sync_each):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    if sync_each:
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
        return tot / reps
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = torch.empty((Q, B.shape[0]), device='cuda', dtype=torch.float16)
events, arr = make_phase_events()

def ours():
    acc = 0.0
    for _ in range(3): s.batched_search_device(q, 100, phase_events=arr)
    torch.cuda.synchronize()
    for _ in range(10):
        s.batched_search_device(q, 100, phase_events=arr); torch.cuda.synchronize()
        acc += events[3].elapsed_time(events[4]) / 10
    print("knn_gemm2_kernel filter pass inside a search step (scores + threshold filter epilogue): %.3f ms = %.0f TF/s"
          % (acc, 2.0 * Q * (N - 8192) * 3000 / acc / 1e9), flush=True)

def cublas(name, sync_each, reps):
    ms = timed(lambda: torch.matmul(A, B.t(), out=out), reps, sync_each)
    print("cuBLAS fp16 %d x %d x 3008, %s: %.3f ms = %.0f TF/s" % (Q, B.shape[0], name, ms, flops / ms / 1e9), flush=True)

# interleaved so that neither side always runs on the chip the other one heated
ours(); cublas("one launch at a time", True, 20); time.sleep(1.0)
ours(); cublas("back to back x50", False, 50); time.sleep(1.0)
ours(); cublas("one launch at a time", True, 20)
