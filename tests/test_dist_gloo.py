"""world_size-2 gloo run (CPU) of the sharded-search plumbing: shard bounds, global
ids, rank-major gather layout, merge rule.  The local scan and the merge are the
oracle here (the CUDA kernels need a GPU); test_gpu_parity covers those."""
import os
import socket

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from morna_b200 import dist as mdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_merge(ids, dists, k):
    from oracle import morna_oracle as mo
    oi = torch.full((ids.shape[0], k), -1, dtype=torch.int32)
    od = torch.full((ids.shape[0], k), float("inf"), dtype=torch.float64)
    for q in range(ids.shape[0]):
        valid = ids[q] >= 0
        i, d = mo.topk_rule(dists[q][valid].numpy(), k, ids=ids[q][valid].numpy())
        oi[q, : len(i)] = torch.from_numpy(i.astype(np.int32))
        od[q, : len(i)] = torch.from_numpy(d)
    return oi, od


def _worker(rank, world, port, S, Q, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    lo, hi = mdist.shard_bounds(S.shape[0], rank, world)

    def local_search(queries, kk):
        ids = torch.full((queries.shape[0], kk), -1, dtype=torch.int32)
        dists = torch.full((queries.shape[0], kk), float("inf"), dtype=torch.float64)
        for q in range(queries.shape[0]):
            if hi > lo:
                i, d = c_oracle.exact_search(S[lo:hi], queries[q].numpy(), kk)
                ids[q, : len(i)] = torch.from_numpy(i + lo)
                dists[q, : len(i)] = torch.from_numpy(d)
        return ids, dists

    ids, dists = mdist.sharded_exact_search(local_search, torch.from_numpy(Q), k, merge=_oracle_merge)
    out[rank] = (ids.numpy(), dists.numpy())
    td.destroy_process_group()


def test_shard_bounds_cover_rows_once():
    for n in (0, 1, 7, 8, 21504, 1000000):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_two_rank_sharded_search_equals_single_process():
    from oracle import c_oracle
    rng = np.random.default_rng(0)
    S = rng.standard_normal((101, 24)).astype(np.float32)
    S[60] = S[3]; S[90] = 2 * S[3]                      # ties across the shard boundary
    Q = np.concatenate([S[[3, 70]].astype(np.float64), rng.standard_normal((2, 24))])
    k = 12
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, _free_port(), S, Q, k, out), nprocs=2, join=True)
    for q in range(Q.shape[0]):
        want_i, want_d = c_oracle.exact_search(S, Q[q], k)
        for rank in range(2):
            assert np.array_equal(out[rank][0][q], want_i)
            assert np.array_equal(out[rank][1][q], want_d)
