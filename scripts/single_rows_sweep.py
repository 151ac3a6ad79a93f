"""Single-query kernel at N=21504 (and others): rows per warp pass sweep, in-kernel stamps + graph replay time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch
from morna_b200 import _lib
lib = _lib.load()
K, D = 100, 3000
for N in [int(a) for a in sys.argv[1:]] or (21504,):
    g = torch.Generator(device='cuda'); g.manual_seed(1234)
    S = torch.randn((N, D), generator=g, device='cuda')
    s = MornaSearch(vectors=S, stats=(N, N, D))
    q = S[N // 3].double().contiguous()
    oi = torch.empty((1, K), dtype=torch.int32, device='cuda'); od = torch.empty((1, K), dtype=torch.float64, device='cuda')
    sws = _lib.workspace(lib.morna_knn_single_workspace_bytes(N), 'cuda'); fb = torch.zeros(1, dtype=torch.int32, device='cuda')
    _lib.check(lib.morna_knn_single_workspace_init(_lib.dev_ptr(sws), sws.numel(), _lib.stream_ptr()), 'init')
    def single():
        _lib.check(lib.morna_knn_single(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, 0, _lib.dev_ptr(q), K,
                                        _lib.dev_ptr(oi), _lib.dev_ptr(od), _lib.dev_ptr(fb), _lib.dev_ptr(sws), sws.numel(),
                                        _lib.stream_ptr()), "single")
    ref = None
    for R in (0, 1, 2, 3, 4, 5):
        lib.morna_debug_set_tuning(3, R)
        for _ in range(3): single()
        torch.cuda.synchronize()
        if ref is None: ref = (oi.clone(), od.clone())
        assert torch.equal(oi, ref[0]) and torch.equal(od, ref[1]) and int(fb.item()) == 0
        acc = [0.0] * 3
        for _ in range(10):
            single(); torch.cuda.synchronize()
            st = sws[16:48].view(torch.int64).cpu().tolist()
            acc[0] += (st[1] - st[0]) / 1e4; acc[1] += (st[2] - st[1]) / 1e4; acc[2] += (st[3] - st[2]) / 1e4
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            single()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): gr.replay()
        e1.record(); torch.cuda.synchronize()
        print("N=%d rows/pass=%d: scan %.1f us, select %.1f us, order+write %.1f us | graph replay back-to-back %.1f us"
              % (N, R, acc[0], acc[1], acc[2], e0.elapsed_time(e1) * 1e3 / 50), flush=True)
    lib.morna_debug_set_tuning(3, 0)
