"""Single-query stream (morna_knn_single_stream) at 21,504 x 3000: us per query for each hand-over mode
(morna_debug_set_tuning key 29: 0 plain launches, 1 next scan starts after this scan, 2 after the first row pass),
m queries per call, direct launches and CUDA-graph replay, against one-at-a-time morna_knn_single."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from morna_b200 import _lib, synth          # noqa: E402
from morna_b200.search import MornaSearch   # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device("cuda:0")
    n, dim, k, m = 21504, 3000, 100, 32
    S = synth.gauss(n, dim, dev, 4321)
    srch = MornaSearch(vectors=S, stats=(n, n, dim), device=dev)
    Q = S[torch.arange(m, device=dev) * 600 + 7].to(torch.float64).contiguous()
    want_i, want_d = [], []
    for j in range(m):
        i_, d_ = srch.single_search_device(Q[j], k)
        want_i.append(i_[0]); want_d.append(d_[0])
    want_i, want_d = torch.stack(want_i), torch.stack(want_d)
    side = torch.cuda.Stream(device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for chain in (0, 1, 2, 3, 2, 3):
        lib.morna_debug_set_tuning(29, chain)
        with torch.cuda.stream(side):
            ids, d = srch.single_search_stream(Q, k)
            assert torch.equal(ids, want_i) and torch.equal(d, want_d), "chain %d differs" % chain
            ws = srch._single_stream_workspace(n)
            flags = torch.zeros(m, dtype=torch.int32, device=dev)

            def call():
                _lib.check(lib.morna_knn_single_stream(_lib.dev_ptr(srch.vectors), _lib.dev_ptr(srch.pp), n, dim, srch.ld, 0,
                                                       _lib.dev_ptr(Q), dim, m, k, _lib.dev_ptr(ids), _lib.dev_ptr(d), _lib.dev_ptr(flags),
                                                       _lib.dev_ptr(ws), ws.numel(), _lib.stream_ptr()), "stream")
            for _ in range(3):
                call()
            e0.record(side)
            for _ in range(10):
                call()
            e1.record(side)
            side.synchronize()
            direct = e0.elapsed_time(e1) * 1e3 / (10 * m)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                call()
            for _ in range(3):
                graph.replay()
            e0.record(side)
            for _ in range(10):
                graph.replay()
            e1.record(side)
            side.synchronize()
            replay = e0.elapsed_time(e1) * 1e3 / (10 * m)
            assert torch.equal(ids, want_i) and torch.equal(d, want_d), "chain %d differs after the replays" % chain
        gb = 4.0 * n * dim / 1e9
        print("chain %d: direct %.2f us/query (%.0f GB/s)   graph replay %.2f us/query (%.0f GB/s)"
              % (chain, direct, gb / (direct * 1e-6), replay, gb / (replay * 1e-6)), flush=True)
    lib.morna_debug_set_tuning(29, 2)
    # few queries at once: the FP64 scan (morna_knn_exact, several queries per pass over the rows) against the stream of
    # single-query kernels
    for nq in (2, 4, 8, 16):
        for name, fn in (("knn_exact", lambda: srch.exact_search_device(Q[:nq], k, allow_single=False)),
                         ("single stream", lambda: srch.single_search_stream(Q[:nq], k))):
            for _ in range(3):
                fn()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print("nq %2d  %-14s %.1f us per call" % (nq, name, e0.elapsed_time(e1) * 100), flush=True)


if __name__ == "__main__":
    main()
