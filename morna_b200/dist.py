"""Multi-GPU exact search: the sample matrix is row-sharded (row = internal id) in
contiguous blocks, one process per GPU; every rank answers the replicated queries
on its block, the per-rank top-k lists are all-gathered (NCCL over NVLink; gloo in
CPU tests) and merged under the reference order.  The union of per-shard exact
top-k lists contains the global top-k, so the result equals one process scanning
all rows (morna.py:697-712).  This is the only exchange step on the path.
"""
import torch

from . import _lib


def shard_bounds(n_rows, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: ceil(n/world) rows per rank."""
    per = -(-n_rows // world)
    lo = min(rank * per, n_rows)
    return lo, min(lo + per, n_rows)


def merge_topk(ids, dists, k, stream=None):
    """Exact top-k of concatenated candidate lists.  ids int32 [nq x m], dists float64
    [nq x m] on the GPU; entries with id < 0 are padding.  Order: distance ascending,
    equal distances id descending (morna.py:705-712)."""
    lib = _lib.load()
    assert ids.is_cuda and dists.is_cuda and ids.shape == dists.shape
    ids = ids.contiguous().to(torch.int32)
    dists = dists.contiguous().to(torch.float64)
    nq, m = ids.shape
    out_i = torch.empty((nq, k), dtype=torch.int32, device=ids.device)
    out_d = torch.empty((nq, k), dtype=torch.float64, device=ids.device)
    if nq == 0:
        return out_i, out_d
    with torch.cuda.device(ids.device):
        ws = _lib.workspace(lib.morna_select_topk_workspace_bytes(m, nq, k), ids.device)
        done = 0
        while done < nq:                     # grid.y limit of the select kernel
            cnt = min(nq - done, 65535)
            _lib.check(lib.morna_select_topk(_lib.dev_ptr(dists[done:]), _lib.dev_ptr(ids[done:]), m, m, 0, cnt, k,
                                             _lib.dev_ptr(out_i[done:]), _lib.dev_ptr(out_d[done:]),
                                             _lib.dev_ptr(ws), ws.numel(), _lib.stream_ptr(stream)),
                       "morna_select_topk")
            done += cnt
    return out_i, out_d


def merge_sorted_lists(ids, dists, k, stream=None):
    """Exact top-k of G sorted lists per query.  ids int32 / dists float64 [G x nq x k_in] on the GPU, every
    list sorted under the reference order (what each shard's search returns), padding id -1 / +inf."""
    lib = _lib.load()
    assert ids.is_cuda and dists.is_cuda and ids.shape == dists.shape and ids.dim() == 3
    ids = ids.contiguous().to(torch.int32)
    dists = dists.contiguous().to(torch.float64)
    g, nq, k_in = ids.shape
    out_i = torch.empty((nq, k), dtype=torch.int32, device=ids.device)
    out_d = torch.empty((nq, k), dtype=torch.float64, device=ids.device)
    with torch.cuda.device(ids.device):
        _lib.check(lib.morna_merge_sorted_topk(_lib.dev_ptr(dists), _lib.dev_ptr(ids), g, nq, k_in, k,
                                               _lib.dev_ptr(out_i), _lib.dev_ptr(out_d), _lib.stream_ptr(stream)),
                   "morna_merge_sorted_topk")
    return out_i, out_d


def all_gather_sorted(ids, dists, group=None):
    """Every rank's [nq x k] lists -> [world x nq x k] (one collective per tensor, no concatenation)."""
    import torch.distributed as td
    world = td.get_world_size(group)
    gi = torch.empty((world,) + tuple(ids.shape), dtype=ids.dtype, device=ids.device)
    gd = torch.empty((world,) + tuple(dists.shape), dtype=dists.dtype, device=dists.device)
    td.all_gather_into_tensor(gi, ids.contiguous(), group=group)
    td.all_gather_into_tensor(gd, dists.contiguous(), group=group)
    return gi, gd


def all_gather_topk(ids, dists, group=None):
    """Gather every rank's [nq x k] lists -> [nq x world*k], rank-major along dim 1."""
    import torch.distributed as td
    world = td.get_world_size(group)
    if world == 1:
        return ids, dists
    gi = [torch.empty_like(ids) for _ in range(world)]
    gd = [torch.empty_like(dists) for _ in range(world)]
    td.all_gather(gi, ids.contiguous(), group=group)
    td.all_gather(gd, dists.contiguous(), group=group)
    return torch.cat(gi, dim=1), torch.cat(gd, dim=1)


def sharded_exact_search(local_search, queries, k, group=None, merge=None):
    """`local_search(queries, k)` -> this rank's (ids, dists) with GLOBAL internal ids, sorted under the
    reference order.  Returns the merged global (ids, dists), identical on every rank.  On the GPU the
    lists are gathered as [world x nq x k] and merged by rank counting (morna_merge_sorted_topk); a
    `merge(ids [nq x world*k], dists, k)` callable replaces that (the CPU tests pass their own reference rule)."""
    import torch.distributed as td
    ids, dists = local_search(queries, k)
    if not (td.is_available() and td.is_initialized() and td.get_world_size(group) > 1):
        return ids, dists
    if merge is None and ids.is_cuda:
        gi, gd = all_gather_sorted(ids, dists, group)
        return merge_sorted_lists(gi, gd, k)
    ids, dists = all_gather_topk(ids, dists, group)
    return (merge or merge_topk)(ids, dists, k)


def sharded_batched_search(search, queries, k, group=None, check_overflow=True):
    """Rows-sharded exact search of a query batch on the tensor-core path, one process per GPU.  `search` holds this
    rank's row block (MornaSearch(..., shard=(rank, world))), `queries` (CUDA float64 [nq x dim]) are replicated.

      1. every rank scores the queries against its rows and writes, per query, lower bounds of the true cosines of
         its k best rows (fp16 score minus the rank's rigorous error bound);
      2. ONE all-gather (4 * nq * k bytes per rank) and a k-th-largest over the world * k values per query give a bound
         that at least k rows over ALL shards reach -- so the ranks together re-rank about k rows per query instead of
         k rows each;
      3. every rank re-ranks (exact FP64) and orders its surviving rows; ONE more all-gather moves the [nq x k] lists
         (ids and distances packed in one buffer) and every rank merges them by rank counting.

    Returns (ids, dists), identical on every rank and identical to one process scanning all rows."""
    import torch.distributed as td
    world = td.get_world_size(group) if (td.is_available() and td.is_initialized()) else 1
    if world == 1:
        return search.batched_search_device(queries, k, check_overflow=check_overflow)
    vals = search.batched_score_bound(queries, k)
    every = torch.empty((world,) + tuple(vals.shape), dtype=vals.dtype, device=vals.device)
    td.all_gather_into_tensor(every, vals, group=group)
    bound = search.union_kth_bound(every, k)
    nq = queries.shape[0]
    mine = torch.empty(nq * k * 12, dtype=torch.uint8, device=queries.device)      # the send buffer: the lists are written into it
    views = (mine[nq * k * 8:].view(torch.int32).view(nq, k), mine[:nq * k * 8].view(torch.float64).view(nq, k))
    search.batched_finish_bound(bound, check_overflow=check_overflow, out=views)
    return gather_merge_packed(views[0], views[1], k, group, packed=mine)


def gather_merge_packed(ids, dists, k, group=None, packed=None):
    """All-gather of every rank's sorted [nq x k] lists in ONE collective (distances and ids packed into one byte
    buffer per rank: `packed`, or packed here), then the rank-counting merge."""
    import torch.distributed as td
    world = td.get_world_size(group)
    nq, k_in = ids.shape
    nd, ni = nq * k_in * 8, nq * k_in * 4
    mine = packed
    if mine is None:
        mine = torch.empty(nd + ni, dtype=torch.uint8, device=ids.device)
        mine[:nd].view(torch.float64).copy_(dists.reshape(-1))
        mine[nd:].view(torch.int32).copy_(ids.reshape(-1))
    every = torch.empty((world, nd + ni), dtype=torch.uint8, device=ids.device)
    td.all_gather_into_tensor(every, mine, group=group)
    if (nq * k_in) % 2:                       # odd element count: the packed ids would not be 8-byte aligned per rank
        gd = every[:, :nd].contiguous().view(torch.float64).view(world, nq, k_in)
        gi = every[:, nd:].contiguous().view(torch.int32).view(world, nq, k_in)
        return merge_sorted_lists(gi, gd, k)
    lib = _lib.load()
    out_i = torch.empty((nq, k), dtype=torch.int32, device=ids.device)
    out_d = torch.empty((nq, k), dtype=torch.float64, device=ids.device)
    with torch.cuda.device(ids.device):
        _lib.check(lib.morna_merge_packed_topk(_lib.dev_ptr(every), world, nq, k_in, k, _lib.dev_ptr(out_i), _lib.dev_ptr(out_d),
                                               _lib.stream_ptr()), "morna_merge_packed_topk")
    return out_i, out_d
