"""Single-query exact search (BASELINE configs[1]: 21,504 x 3000, k=100): per-kernel device times
with an L2 flush between repetitions, against the HBM roofline 4*N*D bytes per query."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch
from morna_b200 import _lib
lib = _lib.load()
peak = 6550.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
K, D = 100, 3000
for N in [int(a) for a in sys.argv[1:]] or (21504, 50000, 125000):
    g = torch.Generator(device='cuda'); g.manual_seed(1234)
    S = torch.randn((N, D), generator=g, device='cuda')
    s = MornaSearch(vectors=S, stats=(N, N, D))
    q = S[N // 3].double()[None, :].contiguous()
    dist = torch.empty((1, N), dtype=torch.float64, device='cuda')
    ws = _lib.workspace(lib.morna_select_topk_workspace_bytes(N, 1, K), 'cuda')
    oi = torch.empty((1, K), dtype=torch.int32, device='cuda'); od = torch.empty((1, K), dtype=torch.float64, device='cuda')
    def scan():
        _lib.check(lib.morna_angular_distances(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, _lib.dev_ptr(q), 1, D,
                                               _lib.dev_ptr(dist), N, _lib.stream_ptr()), "scan")
    def select():
        _lib.check(lib.morna_select_topk(_lib.dev_ptr(dist), None, N, N, 0, 1, K, _lib.dev_ptr(oi), _lib.dev_ptr(od),
                                         _lib.dev_ptr(ws), ws.numel(), _lib.stream_ptr()), "select")
    def both():
        return s.exact_search_device(q, K, allow_single=False)
    sws = _lib.workspace(lib.morna_knn_single_workspace_bytes(N), 'cuda'); fb = torch.zeros(1, dtype=torch.int32, device='cuda')
    _lib.check(lib.morna_knn_single_workspace_init(_lib.dev_ptr(sws), sws.numel(), _lib.stream_ptr()), 'init')
    def single():
        _lib.check(lib.morna_knn_single(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, 0, _lib.dev_ptr(q), K,
                                        _lib.dev_ptr(oi), _lib.dev_ptr(od), _lib.dev_ptr(fb), _lib.dev_ptr(sws), sws.numel(),
                                        _lib.stream_ptr()), "single")
    res = {}
    def single_ldg():
        lib.morna_debug_set_tuning(3, 4); single(); lib.morna_debug_set_tuning(3, 0)
    graphs = {}
    def graphed(name, fn):
        """Replay the call from a CUDA graph so host launch overhead stays out of the device timeline."""
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g.replay
    single_g = graphed("single", single)
    for name, fn in (("single_graph", single_g), ("scan", scan), ("select", select), ("scan+select", both), ("single", single), ("single_ldg", single_ldg)):
        for _ in range(3):
            fn()
        tot = 0.0
        reps = 20
        for _ in range(reps):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1) / reps
        res[name] = tot * 1e3
    # back to back, no flush (the 4*N*D-byte matrix is larger than L2 for N >= 21504) and read-only flush
    for label, pre in (("back-to-back", None), ("read-flush", lambda: flush.view(torch.int32).sum())):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        tot, reps = 0.0, 50
        if pre is None:
            e0.record()
            for _ in range(reps): single_g()
            e1.record(); torch.cuda.synchronize(); tot = e0.elapsed_time(e1) / reps
        else:
            for _ in range(reps):
                pre(); e0.record(); single_g(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1) / reps
        res["g_" + label] = tot * 1e3
        print("N=%d: graph replay, %s: %.1f us -> %.0f%% of the HBM floor" % (N, label, tot * 1e3, 100 * (4.0 * N * D / (peak * 1e9) * 1e6) / (tot * 1e3)))
    for _ in range(3):
        single(); torch.cuda.synchronize()
        st = sws[16:80].view(torch.int64).cpu().tolist()
        print("N=%d: in-kernel stamps: scan %.1f us, k-th key %.1f us (pivot %.1f, filter %.1f, k-th of survivors %.1f, emit %.1f), sort+write %.1f us" % (
              N, (st[1] - st[0]) / 1e3, (st[2] - st[1]) / 1e3, (st[4] - st[1]) / 1e3, (st[5] - st[4]) / 1e3, (st[6] - st[5]) / 1e3,
              (st[2] - st[6]) / 1e3, (st[3] - st[2]) / 1e3))
    ids, d = both()
    assert int(ids[0, 0]) == N // 3 and float(d[0, 0]) == 0.0
    single(); torch.cuda.synchronize()
    assert int(fb.item()) == 0 and torch.equal(oi, ids) and torch.equal(od, d)
    print("N=%d: fused fp64 scan+select replayed from a CUDA graph %.1f us -> %.0f q/s, %.0f%% of the HBM floor" % (
          N, res["single_graph"], 1e6 / res["single_graph"], 100 * (4.0 * N * D / (peak * 1e9) * 1e6) / res["single_graph"]))
    print("N=%d: fused fp64 scan+select (stream launch) %.1f us -> %.0f q/s, %.0f%% of the HBM floor; forced 4 rows/pass %.1f us" % (
          N, res["single"], 1e6 / res["single"], 100 * (4.0 * N * D / (peak * 1e9) * 1e6) / res["single"], res["single_ldg"]))
    floor = 4.0 * N * D / (peak * 1e9) * 1e6
    print("N=%d: scan %.1f us (%.0f GB/s, %.0f%% of %.0f), select %.1f us, scan+select %.1f us -> %.0f q/s, %.0f%% of the HBM floor %.1f us"
          % (N, res["scan"], 4.0 * N * D / res["scan"] / 1e3, 100 * floor / res["scan"], peak, res["select"], res["scan+select"],
             1e6 / res["scan+select"], 100 * floor / res["scan+select"], floor))
    del S, s
