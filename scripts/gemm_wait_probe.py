"""Where the MMA-issuing thread of the pair GEMM waits: cycles on the operand (TMA) barriers, on a free TMEM accumulator
(epilogue), and in total, per CTA pair, for one filter pass at the headline shape."""
import sys, os
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
for _ in range(3): s.batched_search_device(q, K)
buf = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
lib.morna_debug_gemm_counters(_lib.dev_ptr(buf))
s.batched_search_device(q, K); torch.cuda.synchronize()      # the last GEMM launch (filter pass) leaves its counters
lib.morna_debug_gemm_counters(None)
c = buf.view(148, 4).cpu()
lead = c[::2]                                                 # leader CTAs issue the MMAs
tot = lead[:, 2].float()
tiles = (16 * ((N - 8192 + 255) // 256)) / 74.0
print("epilogue warp 2 busy: %.0f k cycles per CTA = %.1f %% of the MMA thread's total, %.0f cycles per tile (a tile's MMAs: %.0f cycles)"
      % (c[:, 3].float().mean() / 1e3, 100 * c[:, 3].float().mean() / tot.mean(), c[:, 3].float().mean() / tiles, tot.mean() / tiles))
print("filter pass, per leader CTA: total %.0f k cycles; waiting for operand tiles %.1f %% (min %.1f, max %.1f), for a free accumulator %.1f %%"
      % (tot.mean() / 1e3, 100 * (lead[:, 0].float() / tot).mean(), 100 * (lead[:, 0].float() / tot).min(), 100 * (lead[:, 0].float() / tot).max(),
         100 * (lead[:, 1].float() / tot).mean()))
