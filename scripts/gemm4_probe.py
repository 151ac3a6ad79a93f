"""The cluster-of-two-pairs GEMM (multicast sample tile, key 0 = 2) against the pair GEMM: results, phase times, MMA wait share."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch, make_phase_events
from morna_b200 import _lib, synth
lib = _lib.load()
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.gauss(N, D, "cuda")
s = MornaSearch(vectors=S, stats=(N, N, D))
q, rows = synth.queries(S, Q)
s.enable_tensor_path()
events, arr = make_phase_events()
buf = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
ref = None
for rnd in range(2):
    for variant, promo, stages in ((1, 3, 4), (1, 2, 4), (1, 0, 4), (1, 3, 6), (2, 3, 4)):
        lib.morna_debug_set_tuning(0, variant); lib.morna_debug_set_tuning(24, promo); lib.morna_debug_set_tuning(1, stages)
        for _ in range(3): s.batched_search_device(q, K, phase_events=arr)
        torch.cuda.synchronize(); time.sleep(0.3)
        acc = [0.0] * 6
        for _ in range(8):
            ids, d = s.batched_search_device(q, K, phase_events=arr); torch.cuda.synchronize()
            for i in range(6): acc[i] += events[i].elapsed_time(events[i + 1]) / 8
        if ref is None: ref = (ids.clone(), d.clone())
        buf.zero_()
        lib.morna_debug_gemm_counters(_lib.dev_ptr(buf)); s.batched_search_device(q, K); torch.cuda.synchronize(); lib.morna_debug_gemm_counters(None)
        c = buf.view(148, 4).cpu()[::2].float(); c = c[c[:, 2] > 0]
        print("GEMM variant %d, L2 promotion %d, %d stages: pilot GEMM %.3f ms, filter GEMM %.3f ms | %d MMA threads, %.0f k cycles, %.1f %% waiting for tiles | same results %s"
              % (variant, promo, stages, acc[1], acc[3], c.shape[0], c[:, 2].mean() / 1e3, 100 * (c[:, 0] / c[:, 2]).mean(), torch.equal(ids, ref[0]) and torch.equal(d, ref[1])), flush=True)
lib.morna_debug_set_tuning(0, 1); lib.morna_debug_set_tuning(24, 3); lib.morna_debug_set_tuning(1, 4)
