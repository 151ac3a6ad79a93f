"""Sparse, tie-heavy indexes (bench.py datasets `sparse_*`): the CSR exact path's step time at the headline shape, for timing
and for ncu captures of sparse_distances_kernel / select_radix_kernel.  usage: sparse_probe.py [sparse|sparse_fixture] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morna_b200 import synth                      # noqa: E402
from morna_b200.search import MornaSearch         # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "sparse_fixture"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if len(sys.argv) > 3:
    from morna_b200 import _lib
    _lib.load().morna_debug_set_tuning(35, int(sys.argv[3]))
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.matrix(kind, N, D, "cuda")
srch = MornaSearch(vectors=S, stats=(N, N, D))
q, rows = synth.queries(S, Q, seed=99)
assert srch.csr is not None
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ids, d = srch.exact_search_device(q, K)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    ids, d = srch.exact_search_device(q, K)
e1.record(); torch.cuda.synchronize()
print("%s: %.3f ms per 4096-query step (CSR exact path), checksum %d %.6f" % (kind, e0.elapsed_time(e1) / reps, int(ids.sum()), float(d.sum())))
