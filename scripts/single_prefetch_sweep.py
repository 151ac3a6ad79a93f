"""Single query, 21,504 x 3000, k = 100, replayed back to back from a CUDA graph: sweep of the bytes per first-pass row that
the kernel pulls into L2 before it stages the query (morna_debug_set_tuning key 27).  Interleaved repeats, in-kernel stamps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch
from morna_b200 import _lib
lib = _lib.load()
N, D, K = 21504, 3000, 100
g = torch.Generator(device='cuda'); g.manual_seed(1234)
S = torch.randn((N, D), generator=g, device='cuda')
s = MornaSearch(vectors=S, stats=(N, N, D))
q = S[N // 3].double()[None, :].contiguous()
oi = torch.empty((1, K), dtype=torch.int32, device='cuda'); od = torch.empty((1, K), dtype=torch.float64, device='cuda')
sws = _lib.workspace(lib.morna_knn_single_workspace_bytes(N), 'cuda'); fb = torch.zeros(1, dtype=torch.int32, device='cuda')
_lib.check(lib.morna_knn_single_workspace_init(_lib.dev_ptr(sws), sws.numel(), _lib.stream_ptr()), 'init')
def single():
    _lib.check(lib.morna_knn_single(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, 0, _lib.dev_ptr(q), K,
                                    _lib.dev_ptr(oi), _lib.dev_ptr(od), _lib.dev_ptr(fb), _lib.dev_ptr(sws), sws.numel(),
                                    _lib.stream_ptr()), "single")
ref_i, ref_d = s.exact_search_device(q, K, allow_single=False)
values = [v for v in sys.argv[1:]] or ["0", "1024", "2048", "4096", "12032"]    # "bytes[:rows[:rows_per_pass]]"
def apply(v):
    parts = [int(x) for x in v.split(":")] + [0, 0]
    lib.morna_debug_set_tuning(27, parts[0]); lib.morna_debug_set_tuning(28, parts[1]); lib.morna_debug_set_tuning(3, parts[2])
graphs = {}
for v in values:
    apply(v)
    single(); torch.cuda.synchronize()
    assert torch.equal(oi, ref_i) and torch.equal(od, ref_d) and int(fb.item()) == 0
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        single()
    graphs[v] = gr
best = {v: [] for v in values}
for rep in range(5):
    for v in values:
        gr = graphs[v]
        for _ in range(10): gr.replay()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): gr.replay()
        e1.record(); torch.cuda.synchronize()
        best[v].append(e0.elapsed_time(e1) / 200 * 1e3)
for v in values:
    apply(v)
    single(); torch.cuda.synchronize()
    st = sws[16:80].view(torch.int64).cpu().tolist()
    print("prefetch %12s: %s us per query back to back | stamps: scan %.1f, select %.1f, order %.1f us" % (
        v, " ".join("%.1f" % t for t in best[v]), (st[1] - st[0]) / 1e3, (st[2] - st[1]) / 1e3, (st[3] - st[2]) / 1e3), flush=True)
apply('6144')
