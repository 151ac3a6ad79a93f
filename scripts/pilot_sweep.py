"""Headline batch (50,000 x 3000, 4096 queries, k = 100): pilot size (key 10) x first refinement point (key 11, absolute
row; 0 = none) -- serial per-phase times and the pipelined step, results checked against the exact scan."""
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
configs = [(8192, 0), (1024, 8192), (2048, 8192), (1024, 4096), (2048, 16384), (4096, 0), (4096, 16384), (8192, 0)]
for pilot, first in configs:
    lib.morna_debug_set_tuning(10, pilot); lib.morna_debug_set_tuning(11, first)
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
    for rep in range(3):
        m, _ = bench.pipeline_ms(torch, s, [q] * 20, K)
        time.sleep(0.3)
        ms.append(m)
    print("pilot %5d first block to %5d: %s sum %.3f | pipeline %s ok=%s" % (
        pilot, first, ", ".join("%s %.3f" % (n_[:6], v) for n_, v in zip(PHASE_NAMES, acc)), sum(acc), " ".join("%.3f" % m for m in ms[1:]), ok), flush=True)
lib.morna_debug_set_tuning(10, 8192); lib.morna_debug_set_tuning(11, 0)
