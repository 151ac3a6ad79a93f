"""A/B of the re-rank with each query's candidates walked in ascending row order (key 34) against the emission order:
the re-rank call alone, alternating the two settings, medians over many repetitions; results compared bit for bit."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                   # noqa: E402
from morna_b200 import _lib, synth             # noqa: E402
from morna_b200.search import MornaSearch      # noqa: E402

lib = _lib.load()
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.gauss(N, D, "cuda", 1234)
s = MornaSearch(vectors=S, stats=(N, N, D))
q, rows = synth.queries(S, Q, seed=99)
s.enable_tensor_path()
need = lib.morna_knn_batched_workspace_bytes(N, Q, D, K)
ws = _lib.workspace(need, "cuda"); ov = torch.zeros(Q, dtype=torch.uint8, device="cuda"); st = torch.zeros(4, dtype=torch.int32, device="cuda")
oi = torch.empty((Q, K), dtype=torch.int32, device="cuda"); od = torch.empty((Q, K), dtype=torch.float64, device="cuda")
_lib.check(lib.morna_knn_batched_score(_lib.dev_ptr(s.hs), s.ld_h, _lib.dev_ptr(s.rho_max), N, D, 0, _lib.dev_ptr(q), Q, D, K,
           _lib.dev_ptr(ov), _lib.dev_ptr(st), _lib.dev_ptr(ws), ws.numel(), None, None, None, _lib.stream_ptr()), "score")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def rr():
    _lib.check(lib.morna_knn_batched_rerank(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(oi), _lib.dev_ptr(od), _lib.dev_ptr(ov), _lib.dev_ptr(ws), ws.numel(), 0, _lib.stream_ptr()), "rerank")


ref, times = None, {0: [], 1: []}
for rep in range(12):
    for mode in (0, 1):
        lib.morna_debug_set_tuning(34, mode)
        rr()
        e0.record()
        for _ in range(5):
            rr()
        e1.record(); torch.cuda.synchronize()
        times[mode].append(e0.elapsed_time(e1) / 5)
        if ref is None:
            ref = (oi.clone(), od.clone())
        assert torch.equal(oi, ref[0]) and torch.equal(od, ref[1])
for mode in (0, 1):
    t = np.array(times[mode])
    print("sort_rows=%d: re-rank call median %.3f ms (min %.3f, max %.3f) over %d x 5 calls" % (mode, np.median(t), t.min(), t.max(), len(t)))
steps = {0: [], 1: []}
for rep in range(4):
    for mode in (0, 1):
        lib.morna_debug_set_tuning(34, mode)
        s._pipes = {}                                        # the slots capture their CUDA graph with the current setting
        bench.pipeline_ms(torch, s, [q] * 6, K)
        ms, _ = bench.pipeline_ms(torch, s, [q] * 20, K)
        steps[mode].append(ms)
for mode in (0, 1):
    print("sort_rows=%d: pipelined step %s ms" % (mode, " ".join("%.3f" % x for x in steps[mode])))
lib.morna_debug_set_tuning(34, 1)
