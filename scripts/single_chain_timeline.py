"""In-kernel %globaltimer stamps of two consecutive queries of morna_knn_single_stream (21,504 x 3000): when the second
query's kernel starts relative to the first one's scan end (last CTA at the ticket) and kernel end, per hand-over mode.
The stamps live in the control block of each workspace half: [0] CTA 0 starts, [1] last CTA takes the ticket (scan over),
[2] candidates known, [3] answer written."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morna_b200 import _lib, synth          # noqa: E402
from morna_b200.search import MornaSearch   # noqa: E402

lib = _lib.load()
dev = torch.device("cuda:0")
n, dim, k = 21504, 3000, 100
S = synth.gauss(n, dim, dev, 4321)
srch = MornaSearch(vectors=S, stats=(n, n, dim), device=dev)
Q = S[[7, 607]].to(torch.float64).contiguous()
half = lib.morna_knn_single_workspace_bytes(n)
for mode in (0, 1, 2):
    lib.morna_debug_set_tuning(29, mode)
    rows = []
    for rep in range(6):
        srch.single_search_stream(Q, k)
        torch.cuda.synchronize()
        ws = srch._single_stream_workspace(n)
        a = ws[:256].cpu().numpy().view(np.uint64)[2:10].astype(np.int64)          # (ticket + pad = 16 bytes, then 8 stamps)
        b = ws[half:half + 256].cpu().numpy().view(np.uint64)[2:10].astype(np.int64)
        if rep >= 2:
            t0 = a[0]
            rows.append((a[1] - t0, a[3] - t0, b[0] - t0, b[1] - t0, b[3] - t0))
    m = np.median(np.array(rows), axis=0) / 1e3
    print("hand-over %d: query A scan over at %.1f us, A done at %.1f us | query B starts at %.1f us, scan over at %.1f, done at %.1f us"
          % (mode, m[0], m[1], m[2], m[3], m[4]), flush=True)
lib.morna_debug_set_tuning(29, 2)
