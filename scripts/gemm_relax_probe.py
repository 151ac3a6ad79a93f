"""Filter-pass GEMM time and MMA-thread wait share with the epilogue warps sleeping between barrier polls."""
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
for rnd in range(2):
    for ns in (0, 1):
        lib.morna_debug_set_tuning(23, ns)
        for _ in range(3): s.batched_search_device(q, K, phase_events=arr)
        torch.cuda.synchronize(); time.sleep(0.3)
        acc = [0.0] * 6
        for _ in range(8):
            ids, d = s.batched_search_device(q, K, phase_events=arr); torch.cuda.synchronize()
            for i in range(6): acc[i] += events[i].elapsed_time(events[i + 1]) / 8
        lib.morna_debug_gemm_counters(_lib.dev_ptr(buf)); s.batched_search_device(q, K); torch.cuda.synchronize(); lib.morna_debug_gemm_counters(None)
        c = buf.view(148, 4).cpu()[::2].float()
        print("dry epilogue=%d: pilot GEMM %.3f ms, filter GEMM %.3f ms | MMA thread: %.0f k cycles, %.1f %% waiting for tiles | epilogue busy %.0f %% | ok=%s"
              % (ns, acc[1], acc[3], c[:, 2].mean() / 1e3, 100 * (c[:, 0] / c[:, 2]).mean(), 100 * float(buf.view(148, 4).cpu()[:, 3].float().mean() / c[:, 2].mean()), int(ids[0, 0]) == int(rows[0])), flush=True)
lib.morna_debug_set_tuning(23, 0)
