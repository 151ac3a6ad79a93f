"""SM partition experiment, headline batch (50,000 x 3000, 4096 queries, k = 100): the re-rank as X SM-filling CTAs
(morna_debug_set_tuning key 30) beside the scoring kernels of the next batch on at most P CTA pairs (key 31), through the
streaming pipeline (two slots, CUDA graphs).  Serial per-phase times, the pipelined step, results checked against the scan."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from morna_b200.search import MornaSearch, make_phase_events, PHASE_NAMES
from morna_b200 import _lib, synth
lib = _lib.load()
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.gauss(N, D, "cuda")
q, rows = synth.queries(S, Q)
configs = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(0, 0), (44, 52), (60, 44), (52, 48), (36, 56), (0, 0)]
for cfg in configs:
    fat, pairs, evict_first, keep = (list(cfg) + [0, 0])[:4]          # keys 32 / 33: L2 policies of the row gather / the operand tiles
    lib.morna_debug_set_tuning(30, fat); lib.morna_debug_set_tuning(31, pairs)
    lib.morna_debug_set_tuning(32, evict_first); lib.morna_debug_set_tuning(33, keep)
    s = MornaSearch(vectors=S, stats=(N, N, D))
    s.enable_tensor_path()
    pick = torch.arange(0, Q, 61, device="cuda")
    ref_ids, ref_d = s.exact_search_device(q[pick], K, allow_single=False)
    events, arr = make_phase_events()
    for _ in range(3):
        ids, d = s.batched_search_device(q, K, phase_events=arr)
    torch.cuda.synchronize()
    acc = [0.0] * 6
    for _ in range(10):
        ids, d = s.batched_search_device(q, K, phase_events=arr); torch.cuda.synchronize()
        for i in range(6):
            acc[i] += events[i].elapsed_time(events[i + 1]) / 10
    ok = torch.equal(ids[pick], ref_ids) and torch.equal(d[pick], ref_d)
    ms = []
    for rep in range(4):
        m, last = bench.pipeline_ms(torch, s, [q] * 20, K)
        time.sleep(0.3)
        ms.append(m)
    ok2 = torch.equal(torch.as_tensor(last[0])[pick.cpu()], ref_ids.cpu()) and torch.equal(torch.as_tensor(last[1])[pick.cpu()], ref_d.cpu())
    print("re-rank on %3d SMs, GEMM on %2d pairs, rows evict_first=%d, tiles evict_last=%d: %s sum %.3f | pipeline %s ok=%s/%s" % (
        fat, pairs, evict_first, keep, ", ".join("%s %.3f" % (n_[:6], v) for n_, v in zip(PHASE_NAMES, acc)), sum(acc), " ".join("%.3f" % m for m in ms[1:]), ok, ok2), flush=True)
for key in (30, 31, 32, 33):
    lib.morna_debug_set_tuning(key, 0)
