"""Do consecutive batches on their own streams overlap when every kernel asks for the same shared-memory carve-out and
the re-rank is launched one CTA per query (so its CTAs retire one by one)?  20 resident batches through search_batches."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch
from morna_b200 import _lib, synth
lib = _lib.load()
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.gauss(N, D, "cuda")
s = MornaSearch(vectors=S, stats=(N, N, D))
q, rows = synth.queries(S, Q)
s.enable_tensor_path()
def run(depth, nb=20):
    last = None
    for ids_, d_ in s.search_batches((q for _ in range(nb)), K, depth=depth):
        last = ids_
    return last
for carve in (0, 1):
    for oneshot in (0, 1):
        for cap in (0, 4):
            for depth in (2, 3):
                lib.morna_debug_set_tuning(19, carve); lib.morna_debug_set_tuning(20, oneshot); lib.morna_debug_set_tuning(13, cap)
                run(depth, 4); torch.cuda.synchronize(); time.sleep(0.5)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); last = run(depth)
                pipe = s._pipes[(K, depth)]
                for sl in pipe.slots:
                    torch.cuda.current_stream().wait_stream(sl.compute)
                e1.record(); torch.cuda.synchronize()
                print("carve-out hint %d, one CTA per query %d, re-rank CTAs/SM cap %d, depth %d: %.3f ms per batch ok=%s"
                      % (carve, oneshot, cap, depth, e0.elapsed_time(e1) / 20, int(last[0, 0]) == int(rows[0])), flush=True)
lib.morna_debug_set_tuning(19, 0); lib.morna_debug_set_tuning(20, 0); lib.morna_debug_set_tuning(13, 0)
