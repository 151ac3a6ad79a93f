"""search_batches with and without CUDA-graph replay of a slot's step: device ms per batch and host ms per batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch, BatchPipeline
from morna_b200 import _lib, synth
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.gauss(N, D, "cuda")
s = MornaSearch(vectors=S, stats=(N, N, D))
q, rows = synth.queries(S, Q)
hq = q.float().cpu().pin_memory()
s.enable_tensor_path()
ref = None
for graphs in (False, True, False, True):
    BatchPipeline.use_graphs = graphs
    s._pipes = {}
    for src, name in ((q, "resident"), (hq, "host")):
        for _ in s.search_batches((src for _ in range(6)), K): pass
        torch.cuda.synchronize(); time.sleep(0.3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); host = 0.0
        for ids, d in s.search_batches((src for _ in range(20)), K):
            last = ids.copy()
        pipe = s._pipes[(K, 2)]
        for sl in pipe.slots:
            torch.cuda.current_stream().wait_stream(sl.compute); torch.cuda.current_stream().wait_stream(sl.stream)
        e1.record(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 20 * 1e3
        if ref is None: ref = last
        print("graphs=%s %s queries: %.3f ms per batch (device), %.3f ms wall, same results %s" % (graphs, name, e0.elapsed_time(e1) / 20, wall, bool((last == ref).all())), flush=True)
