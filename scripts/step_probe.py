"""One headline batch (50,000 x 3000, 4096 queries, k = 100) under the round-2 schedules, burst regime (20 batches
after a pause, CUDA events on the device):
  serial      score + re-rank, one stream, per-phase times
  rerank      the re-rank kernels alone: warp-granular (rows per pass x phase MB) and CTA-per-query
  pipeline    MornaSearch.search_batches, depth 2, with and without the side job (helper warps in the GEMM kernel)
Every variant's results are compared with the exact scan on a sample of the queries."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from morna_b200.search import MornaSearch, make_phase_events, PHASE_NAMES
from morna_b200 import _lib, synth
lib = _lib.load()
N, D, Q, K = 50000, 3000, 4096, 100
S = synth.gauss(N, D, "cuda")
s = MornaSearch(vectors=S, stats=(N, N, D))
q, rows = synth.queries(S, Q)
s.enable_tensor_path()
pick = torch.arange(0, Q, 61, device="cuda")
ref_ids, ref_d = s.exact_search_device(q[pick], K, allow_single=False)

def check(ids, d, what):
    ok = torch.equal(ids[pick], ref_ids) and torch.equal(d[pick], ref_d)
    if not ok:
        print("!! %s: results differ from the exact scan" % what, flush=True)
    return ok

def timed(fn, reps=20, pause=0.5):
    fn(); torch.cuda.synchronize(); time.sleep(pause)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# ---- serial, per phase
events, arr = make_phase_events()
for _ in range(3):
    ids, d = s.batched_search_device(q, K, phase_events=arr)
torch.cuda.synchronize()
acc = [0.0] * 6
for _ in range(5):
    ids, d = s.batched_search_device(q, K, phase_events=arr); torch.cuda.synchronize()
    for i in range(6):
        acc[i] += events[i].elapsed_time(events[i + 1]) / 5
print("serial phases (ms): " + ", ".join("%s %.3f" % (n_, v) for n_, v in zip(PHASE_NAMES, acc)), "ok=%s" % check(ids, d, "serial"), flush=True)
print("serial step: %.3f ms" % timed(lambda: s.batched_search_device(q, K, check_overflow=False)), flush=True)

# ---- the re-rank alone
need = lib.morna_knn_batched_workspace_bytes(N, Q, D, K)
ws = _lib.workspace(need, "cuda"); ov = torch.zeros(Q, dtype=torch.uint8, device="cuda"); st = torch.zeros(4, dtype=torch.int32, device="cuda")
oi = torch.empty((Q, K), dtype=torch.int32, device="cuda"); od = torch.empty((Q, K), dtype=torch.float64, device="cuda")
def score():
    _lib.check(lib.morna_knn_batched_score(_lib.dev_ptr(s.hs), s.ld_h, _lib.dev_ptr(s.rho_max), N, D, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(ov), _lib.dev_ptr(st), _lib.dev_ptr(ws), ws.numel(), None, None, None, _lib.stream_ptr()), "score")
def rerank():
    _lib.check(lib.morna_knn_batched_rerank(_lib.dev_ptr(s.vectors), _lib.dev_ptr(s.pp), N, D, s.ld, 0, _lib.dev_ptr(q), Q, D, K,
               _lib.dev_ptr(oi), _lib.dev_ptr(od), _lib.dev_ptr(ov), _lib.dev_ptr(ws), ws.numel(), 0, _lib.stream_ptr()), "rerank")
score(); torch.cuda.synchronize()
print("score alone: %.3f ms" % timed(score), flush=True)
for kern, rows_list in ((0, (4, 8, 16)), (1, (8,))):
    for r in rows_list:
        for mb in ((0, 32, 48, 64, 96) if kern == 0 else (0,)):
            for ctas in ((0, 1) if kern == 0 and mb == 64 else (0,)):
                lib.morna_debug_set_tuning(14, kern); lib.morna_debug_set_tuning(5, r); lib.morna_debug_set_tuning(6, mb)
                lib.morna_debug_set_tuning(15, ctas)
                ms = timed(rerank, reps=10, pause=0.2)
                print("rerank kernel=%s rows/pass=%d phase_mb=%d ctas/sm=%d: %.3f ms ok=%s" % ("warp" if kern == 0 else "cta", r, mb, ctas, ms, check(oi, od, "rerank")), flush=True)
lib.morna_debug_set_tuning(14, 0); lib.morna_debug_set_tuning(5, 8); lib.morna_debug_set_tuning(6, -1); lib.morna_debug_set_tuning(15, 0)

# ---- the streaming pipeline, resident queries
def pipeline(nb=20):
    last = None
    for ids_, d_ in s.search_batches((q for _ in range(nb)), K, depth=2, side_job=True):
        last = (ids_, d_)
    return last
for side in (1, 0):
    lib.morna_debug_set_tuning(17, side)
    for r in (8, 4, 16):
        lib.morna_debug_set_tuning(5, r)
        pipeline(4); torch.cuda.synchronize(); time.sleep(0.5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); last = pipeline(20)
        pipe = s._pipes[(K, 2)]
        torch.cuda.current_stream().wait_stream(pipe.compute); [torch.cuda.current_stream().wait_stream(sl_.compute) for sl_ in pipe.slots]
        e1.record(); torch.cuda.synchronize()
        ids_h, d_h = last
        ok = torch.equal(torch.from_numpy(ids_h).cuda()[pick], ref_ids) and torch.equal(torch.from_numpy(d_h).cuda()[pick], ref_d)
        print("pipeline depth 2, side job %s, rows/pass %d: %.3f ms per batch ok=%s" % ("on" if side else "off", r, e0.elapsed_time(e1) / 20, ok), flush=True)
lib.morna_debug_set_tuning(17, 1); lib.morna_debug_set_tuning(5, 8)
