"""Per-phase CUDA-event times of one headline batch (50,000 x 3000, 4096 queries, k = 100), serial on one stream,
plus the pipelined step; results checked against the exact scan on a sample of the queries."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch, make_phase_events, PHASE_NAMES
from morna_b200 import _lib, synth
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.gauss(N, D, "cuda")
s = MornaSearch(vectors=S, stats=(N, N, D))
q, rows = synth.queries(S, Q)
s.enable_tensor_path()
pick = torch.arange(0, Q, 61, device="cuda")
ref_ids, ref_d = s.exact_search_device(q[pick], K, allow_single=False)
events, arr = make_phase_events()
for _ in range(3):
    ids, d = s.batched_search_device(q, K, phase_events=arr)
torch.cuda.synchronize()
acc = [0.0] * 6
for _ in range(10):
    ids, d = s.batched_search_device(q, K, phase_events=arr); torch.cuda.synchronize()
    for i in range(6):
        acc[i] += events[i].elapsed_time(events[i + 1]) / 10
ok = torch.equal(ids[pick], ref_ids) and torch.equal(d[pick], ref_d)
print("serial phases (ms): " + ", ".join("%s %.3f" % (n_, v) for n_, v in zip(PHASE_NAMES, acc)), "sum %.3f" % sum(acc), "ok=%s" % ok, flush=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
batches = [q] * 20
for rep in range(4):
    ms, last = bench.pipeline_ms(torch, s, batches, K)
    time.sleep(0.3)
    print("pipeline: %.3f ms per batch" % ms, flush=True)
