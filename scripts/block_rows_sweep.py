"""1,000,000 x 3000 rows on one GPU, 4096 out-of-index queries, exact top-100 (bench.py `rows_sharded` at N = 1):
rows scored between threshold refinements (key 9) x first refinement point (key 11), step time and the whole-step
fraction of the measured tensor peak; results compared with the default setting bit for bit."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from morna_b200 import _lib, synth, dist as mdist     # noqa: E402
from morna_b200.search import MornaSearch             # noqa: E402

lib = _lib.load()
N, D, Q, K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 3000, 4096, 100
peak = 1673.8
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
except Exception:
    pass
dev = torch.device("cuda:0")
S = synth.gauss(N, D, dev, 1234)
srch = MornaSearch(vectors=S, stats=(N, N, D), device=dev)
del S
srch.enable_tensor_path()
q = synth.gauss(Q, D, dev, 99).to(torch.float64)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
for block, first, growth in ((131072, 0, 1), (131072, 0, 2), (131072, 0, 4), (65536, 0, 2), (131072, 0, 1), (131072, 0, 2)):
    lib.morna_debug_set_tuning(9, block); lib.morna_debug_set_tuning(11, first); lib.morna_debug_set_tuning(36, growth)
    for _ in range(2):
        ids, d = mdist.sharded_batched_search(srch, q, K, check_overflow=False)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        ids, d = mdist.sharded_batched_search(srch, q, K, check_overflow=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if ref is None:
        ref = (ids.clone(), d.clone())
    same = torch.equal(ids, ref[0]) and torch.equal(d, ref[1])
    print("block %8d first %6d growth %d: %.2f ms per step, step fraction %.3f, same=%s"
          % (block, first, growth, ms, 2.0 * Q * N * D / (ms * 1e-3) / 1e12 / peak, same), flush=True)
lib.morna_debug_set_tuning(9, 131072); lib.morna_debug_set_tuning(11, 0); lib.morna_debug_set_tuning(36, 1)
