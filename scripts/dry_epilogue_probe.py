"""Filter GEMM of the headline batch with and without the survivor appends in its epilogue (key 23: dry epilogue, results
invalid) and with fewer survivors (a tighter pilot: key 10 up) -- what the epilogue costs the contraction."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch, make_phase_events, PHASE_NAMES
from morna_b200 import _lib, synth
lib = _lib.load()
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.gauss(N, D, "cuda")
q, rows = synth.queries(S, Q)
s = MornaSearch(vectors=S, stats=(N, N, D))
s.enable_tensor_path()
events, arr = make_phase_events()
for rep in range(2):
    for dry in (0, 1):
        lib.morna_debug_set_tuning(23, dry)
        for _ in range(3):
            s.batched_search_device(q, K, phase_events=arr, check_overflow=False)
        torch.cuda.synchronize()
        acc = [0.0] * 6
        for _ in range(10):
            s.batched_search_device(q, K, phase_events=arr, check_overflow=False); torch.cuda.synchronize()
            for i in range(6):
                acc[i] += events[i].elapsed_time(events[i + 1]) / 10
        flops = 2.0 * Q * (N - 8192) * D
        print("dry epilogue=%d: %s | filter GEMM %.1f TF/s" % (dry, ", ".join("%s %.3f" % (n_[:6], v) for n_, v in zip(PHASE_NAMES, acc)), flops / (acc[3] * 1e-3) / 1e12), flush=True)
lib.morna_debug_set_tuning(23, 0)
