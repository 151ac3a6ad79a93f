"""Re-rank alone at the headline shape (50,000 x 3000, 4096 queries, k = 100): L2 policy of the candidate-row loads
(morna_debug_set_tuning key 32: 0 evict_normal, 1 evict_first, 2 evict_last, 3 half evict_last / half evict_first, 4 no hint),
interleaved repeats, each checked against the exact scan; then the pipelined step per policy."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
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
names = {0: "evict_normal", 1: "evict_first", 2: "evict_last", 3: "half last/half first", 4: "no hint"}
values = [int(v) for v in sys.argv[1:]] or [4, 0, 1, 2, 3]
res = {v: [] for v in values}
for rep in range(4):
    for v in values:
        lib.morna_debug_set_tuning(32, v)
        oi.zero_(); od.zero_()
        ms = timed(rerank)
        ok = torch.equal(oi[pick], ref_ids) and torch.equal(od[pick], ref_d)
        res[v].append("%.3f%s" % (ms, "" if ok else "!!"))
for v in values:
    print("rows %-22s re-rank + order: %s ms" % (names[v], " ".join(res[v])), flush=True)
for rep in range(2):
    for v in values:
        lib.morna_debug_set_tuning(32, v)
        s2 = MornaSearch(vectors=S, stats=(N, N, D)); s2.enable_tensor_path()
        ms = [bench.pipeline_ms(torch, s2, [q] * 20, K)[0] for _ in range(4)]
        print("rows %-22s pipelined step: %s ms" % (names[v], " ".join("%.3f" % m for m in ms[1:])), flush=True)
        time.sleep(0.3)
lib.morna_debug_set_tuning(32, 0)
