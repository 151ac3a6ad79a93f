"""Tensor-core batched search: fp16 score error bound, and bit-identical results vs the exact scan."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import c_oracle
from oracle import morna_oracle as mo
from tests.helpers import GOLDEN, check_topk, tiny_lines

pytestmark = pytest.mark.gpu


def make_search(S, **kw):
    from morna_b200.search import MornaSearch
    return MornaSearch(vectors=S, stats=(S.shape[0], S.shape[0], S.shape[1]), **kw)


def sparse_rows(rng, n, d, nnz):
    S = np.zeros((n, d), np.float32)
    for i in range(n):
        cols = rng.choice(d, size=int(rng.integers(1, nnz + 1)), replace=False)
        S[i, cols] = (rng.integers(1, 50, size=len(cols)) * rng.choice([-1.0, 1.0], size=len(cols))
                      * rng.choice([0.098, 1.21, 1.415, 0.5], size=len(cols)))
    return S


@pytest.mark.parametrize("n,d,kind", [(2048, 3000, "gauss"), (1000, 37, "gauss"), (777, 130, "gauss"),
                                      (3000, 3000, "sparse"), (512, 64, "sparse")])
def test_fp16_scores_stay_within_the_rigorous_bound(n, d, kind):
    from morna_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n + d)
    S = rng.standard_normal((n, d)).astype(np.float32) if kind == "gauss" else sparse_rows(rng, n, d, 3)
    if kind == "gauss":
        S *= np.exp(rng.standard_normal((n, 1))).astype(np.float32)      # mixed row scales
    S[3] = 0.0
    srch = make_search(S)
    srch.enable_tensor_path()
    Q = np.concatenate([S[:100].astype(np.float64), S[100:150] + 0.05 * rng.standard_normal((50, d)),
                        rng.standard_normal((27, d))])
    Q[120] = 0.0
    nq = Q.shape[0]
    dq = torch.from_numpy(Q).cuda()
    ld_s = (n + 3) // 4 * 4
    scores = torch.zeros((nq, ld_s), dtype=torch.float32, device="cuda")
    eps = torch.zeros(nq, dtype=torch.float32, device="cuda")
    ws = _lib.workspace(lib.morna_knn_batched_workspace_bytes(n, nq, d, 10), "cuda")
    rc = lib.morna_debug_tensor_scores(_lib.dev_ptr(srch.hs), srch.ld_h, _lib.dev_ptr(srch.rho_max), n, d,
                                       _lib.dev_ptr(dq), nq, d, _lib.dev_ptr(scores), ld_s, _lib.dev_ptr(eps),
                                       _lib.dev_ptr(ws), ws.numel(), _lib.stream_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    got = scores.cpu().numpy()[:, :n].astype(np.float64)
    S64 = S.astype(np.float64)
    sn = np.linalg.norm(S64, axis=1)
    qn = np.linalg.norm(Q, axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        cos = (Q @ S64.T) / (qn[:, None] * sn[None, :])
    cos[~np.isfinite(cos)] = 0.0                       # zero-norm row or query: cosine defined as 0
    err = np.abs(got - cos)
    e = eps.cpu().numpy().astype(np.float64)
    assert np.all(err <= e[:, None]), "fp16 score error %g exceeds bound %g" % (err.max(), e.min())
    assert e.max() < 2e-3
    assert err.max() > 0          # it really is a reduced-precision pass
    print("max err %.3g, bound %.3g..%.3g" % (err.max(), e.min(), e.max()))


@pytest.mark.parametrize("n,d,nq,k", [(3000, 3000, 300, 100), (20000, 512, 130, 20), (9000, 64, 257, 50),
                                      (500, 3000, 5, 100), (8192, 256, 128, 10), (8449, 100, 129, 1)])
def test_batched_equals_exact_scan_bit_for_bit(n, d, nq, k):
    rng = np.random.default_rng(n * 7 + d)
    S = rng.standard_normal((n, d)).astype(np.float32)
    S[n // 2] = 0.0
    srch = make_search(S)
    rows = rng.permutation(n)[:nq]
    Q = S[rows].astype(np.float64)
    Q[1::3] += 0.05 * rng.standard_normal((len(Q[1::3]), d))
    Q[2::7] = rng.standard_normal((len(Q[2::7]), d))
    if nq > 4:
        Q[4] = 0.0
    q = torch.from_numpy(Q).cuda()
    e_ids, e_d = srch.exact_search_device(q, k)
    b_ids, b_d = srch.batched_search_device(q, k)
    torch.cuda.synchronize()
    # only the all-zero query (every row ties at sqrt(2)) may overflow into the exact scan
    assert srch.last_stats[0] <= (1 if nq > 4 else 0), "unexpected overflow on gaussian data: %r" % (srch.last_stats,)
    assert torch.equal(b_ids, e_ids)
    assert torch.equal(b_d, e_d)
    # and the exact scan is itself checked against the oracle for one query
    true_d = c_oracle.distances(S, Q[0])
    check_topk(true_d, b_ids[0].cpu().numpy()[: min(k, n)], b_d[0].cpu().numpy()[: min(k, n)], tol=1e-9)
    print("stats", srch.last_stats, "per query survivors %.1f final %.1f" % (srch.last_stats[1] / nq, srch.last_stats[2] / nq))


@pytest.mark.parametrize("pair,stages", [(1, 6), (1, 4), (0, 4), (2, 4)])
def test_gemm_variants_agree(pair, stages):
    """CTA-pair (cta_group::2), single-CTA and cluster-of-two-pairs (TMA multicast of the sample tile) GEMM variants give
    the same exact results."""
    from morna_b200 import _lib
    lib = _lib.load()
    try:
        assert lib.morna_debug_set_tuning(0, pair) == 0 and lib.morna_debug_set_tuning(1, stages) == 0
        rng = np.random.default_rng(77)
        n, d, nq, k = 17000, 1000, 700 if pair != 2 else 1024, 64      # (the cluster variant needs a multiple of 512 queries)
        S = rng.standard_normal((n, d)).astype(np.float32)
        srch = make_search(S)
        Q = S[rng.permutation(n)[:nq]].astype(np.float64) + 0.02 * rng.standard_normal((nq, d))
        q = torch.from_numpy(Q).cuda()
        e_ids, e_d = srch.exact_search_device(q, k)
        b_ids, b_d = srch.batched_search_device(q, k)
        assert srch.last_stats[0] == 0
        assert torch.equal(b_ids, e_ids) and torch.equal(b_d, e_d)
    finally:
        lib.morna_debug_set_tuning(0, 1)
        lib.morna_debug_set_tuning(1, 6)


def test_batched_with_massive_ties_falls_back_to_exact_scan():
    oracle = mo.go_index(tiny_lines(), features=3000, sample_threshold=100)
    S = oracle.matrix_f32()
    srch = make_search(S)
    assert srch.csr is not None
    srch.sparse_exact = False                # force the tensor-core path (a sparse index normally skips it)
    rows = np.array([0, 2040, 6595, 5, 4000, 17])
    q = torch.from_numpy(S[rows].astype(np.float64)).cuda()
    e_ids, e_d = srch.exact_search_device(q, 20)
    b_ids, b_d = srch.batched_search_device(q, 20)
    assert srch.last_stats[0] > 0            # thousands of rows tie at distance 0: lists overflow
    assert torch.equal(b_ids, e_ids) and torch.equal(b_d, e_d)
    with open(os.path.join(GOLDEN, "tiny_expected.json")) as fh:
        exp = json.load(fh)
    assert b_ids[0].tolist() == exp["queries"][0]["ids"]


def test_batched_row_blocks_merge():
    rng = np.random.default_rng(5)
    S = rng.standard_normal((2600, 96)).astype(np.float32)
    srch = make_search(S)
    srch.BATCH_BLOCK_ROWS = 1024          # force three row blocks + merge
    q = torch.from_numpy(S[:40].astype(np.float64)).cuda()
    e_ids, e_d = srch.exact_search_device(q, 30)
    b_ids, b_d = srch.batched_search_device(q, 30)
    assert torch.equal(b_ids, e_ids) and torch.equal(b_d, e_d)


@pytest.mark.parametrize("kernel,pipe,rows_per_warp,phase_mb", [(1, 0, 2, 0), (1, 0, 4, 1), (1, 0, 8, 0), (1, 1, 8, 1), (1, 1, 4, 0),
                                                                (1, 1, 16, 0), (0, 0, 8, 0), (0, 0, 16, 1), (0, 0, 2, 1)])
def test_rerank_variants_agree(kernel, pipe, rows_per_warp, phase_mb):
    """Which kernel re-ranks (CTA per query with the query in shared memory, plain or software-pipelined; warp-granular
    queue items with the query read from global memory), rows per warp pass and the L2 phase split only change the
    schedule of the re-rank, never a bit."""
    from morna_b200 import _lib
    lib = _lib.load()
    try:
        for key, val in ((14, kernel), (18, pipe), (5, rows_per_warp), (6, phase_mb)):
            assert lib.morna_debug_set_tuning(key, val) == 0
        rng = np.random.default_rng(123)
        n, d, nq, k = 4100, 300, 600, 50          # every row is re-ranked ~9 times: the phase split engages
        S = rng.standard_normal((n, d)).astype(np.float32)
        S[7, :5] = np.float32(1e-42)              # float32 denormals and signed zeros survive f32_scaled_f64
        S[8, :5] = np.float32(-0.0)
        srch = make_search(S)
        Q = S[rng.permutation(n)[:nq]].astype(np.float64) + 0.02 * rng.standard_normal((nq, d))
        Q[0] = S[7]
        Q[1, :] = 0.0
        Q[1, :5] = 1e-42
        q = torch.from_numpy(Q).cuda()
        e_ids, e_d = srch.exact_search_device(q, k)
        b_ids, b_d = srch.batched_search_device(q, k)
        assert torch.equal(b_ids, e_ids) and torch.equal(b_d, e_d)
    finally:
        for key, val in ((14, 1), (18, 1), (5, 4), (6, -1)):
            lib.morna_debug_set_tuning(key, val)


def test_l2_policy_knobs_change_no_bit():
    """L2 eviction policies of the re-rank's row loads (key 32) and of the GEMM's operand tiles (key 33) are hints only."""
    from morna_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)
    n, d, nq, k = 3000, 160, 300, 25
    S = rng.standard_normal((n, d)).astype(np.float32)
    srch = make_search(S)
    q = torch.from_numpy(S[rng.permutation(n)[:nq]].astype(np.float64) + 0.03 * rng.standard_normal((nq, d))).cuda()
    e_ids, e_d = srch.exact_search_device(q, k)
    try:
        for rows_policy, keep in ((0, 0), (1, 0), (2, 1), (3, 0), (4, 1)):
            assert lib.morna_debug_set_tuning(32, rows_policy) == 0 and lib.morna_debug_set_tuning(33, keep) == 0
            b_ids, b_d = srch.batched_search_device(q, k)
            assert torch.equal(b_ids, e_ids) and torch.equal(b_d, e_d)
    finally:
        lib.morna_debug_set_tuning(32, 0)
        lib.morna_debug_set_tuning(33, 0)


@pytest.mark.parametrize("fat_sms,pairs", [(8, 4), (60, 44), (2, 0)])
def test_sm_partition_changes_no_bit(fat_sms, pairs):
    """The re-rank as SM-filling CTAs in pairs drawing queries from the batch's counter (key 30) beside scoring kernels held
    to a number of CTA pairs (key 31) -- the SM-partition experiment -- returns the same lists, synchronously and streamed."""
    from morna_b200 import _lib
    lib = _lib.load()
    try:
        assert lib.morna_debug_set_tuning(30, fat_sms) == 0 and lib.morna_debug_set_tuning(31, pairs) == 0
        rng = np.random.default_rng(77)
        n, d, nq, k = 5000, 200, 700, 40
        S = rng.standard_normal((n, d)).astype(np.float32)
        srch = make_search(S)
        batches = [S[rng.permutation(n)[:nq]].astype(np.float64) + 0.05 * rng.standard_normal((nq, d)) for _ in range(4)]
        want = [srch.exact_search_device(torch.from_numpy(b).cuda(), k) for b in batches]
        b_ids, b_d = srch.batched_search_device(torch.from_numpy(batches[0]).cuda(), k)
        assert torch.equal(b_ids, want[0][0]) and torch.equal(b_d, want[0][1])
        for (ids, dd), (w_ids, w_d) in zip(srch.search_batches(batches, k), want):
            assert np.array_equal(ids, w_ids.cpu().numpy()) and np.array_equal(dd, w_d.cpu().numpy())
    finally:
        lib.morna_debug_set_tuning(30, 0)
        lib.morna_debug_set_tuning(31, 0)


@pytest.mark.parametrize("side", [1, 0])
def test_side_job_pipeline_equals_the_synchronous_call(side):
    """The scoring call of batch i+1 carrying batch i's re-rank as its side job (helper warps inside the GEMM kernel,
    then the resume call) returns the same lists as the plain call -- with the side job on (key 17) and ignored."""
    from morna_b200 import _lib
    lib = _lib.load()
    try:
        assert lib.morna_debug_set_tuning(17, side) == 0 and lib.morna_debug_set_tuning(14, 0 if side else 1) == 0
        rng = np.random.default_rng(17)
        n, d, k = 9000, 520, 30
        S = rng.standard_normal((n, d)).astype(np.float32)
        S[20:23] = S[19]
        srch = make_search(S)
        batches = []
        for nq in (700, 700, 333, 700, 64):
            Q = S[rng.permutation(n)[:nq]] + np.float32(0.03) * rng.standard_normal((nq, d)).astype(np.float32)
            Q[0] = S[19]
            batches.append(Q)
        want = [srch.exact_search_batch(B, k, tensor_cores=False) for B in batches]
        got = [(i.copy(), d_.copy()) for i, d_ in srch.search_batches(iter(batches), k, depth=2, side_job=True)]
        for (gi, gd), (wi, wd) in zip(got, want):
            assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
    finally:
        lib.morna_debug_set_tuning(17, 0)
        lib.morna_debug_set_tuning(14, 1)


def test_queries_too_large_for_the_scaled_rerank_go_to_the_exact_scan():
    rng = np.random.default_rng(9)
    n, d, nq, k = 3000, 128, 70, 10
    S = rng.standard_normal((n, d)).astype(np.float32)
    srch = make_search(S)
    Q = rng.standard_normal((nq, d))
    Q[3, 5] = 3e38                 # >= 2^127: q * 2^896 would overflow a double
    Q[11] *= 1e300
    q = torch.from_numpy(Q).cuda()
    e_ids, e_d = srch.exact_search_device(q, k)
    b_ids, b_d = srch.batched_search_device(q, k)
    assert srch.last_stats[0] == 2
    assert torch.equal(b_ids, e_ids) and torch.equal(b_d, e_d)


def test_streaming_batches_equal_the_synchronous_call():
    """MornaSearch.search_batches: several batches in flight on their own streams, same results."""
    rng = np.random.default_rng(31)
    n, d, k = 6000, 200, 25
    S = rng.standard_normal((n, d)).astype(np.float32)
    S[10:14] = S[9]                                  # a few exact ties
    srch = make_search(S)
    batches = []
    for b, nq in enumerate((300, 300, 64, 500, 300)):
        Q = S[rng.permutation(n)[:nq]] + np.float32(0.03) * rng.standard_normal((nq, d)).astype(np.float32)
        Q[0] = S[9]
        batches.append(Q if b % 2 == 0 else Q.astype(np.float64))
    batches.append(torch.from_numpy(batches[0]).pin_memory())        # caller-pinned source, no staging copy
    want = [srch.exact_search_batch(np.asarray(B), k, tensor_cores=False) for B in batches]
    member_ids = np.array([9, 10, 5999, 0, 77] * 20, dtype=np.int64)   # stored rows named by internal id (search -q)
    batches.append(member_ids)
    want.append(srch.exact_search_batch(S[member_ids], k, tensor_cores=False))
    for depth in (1, 2, 3):
        got = [(i.copy(), d_.copy()) for i, d_ in srch.search_batches(iter(batches), k, depth=depth)]
        assert len(got) == len(want)
        for (gi, gd), (wi, wd) in zip(got, want):
            assert np.array_equal(gi, wi) and np.array_equal(gd, wd)


def test_streaming_batches_handle_tie_overflow_and_row_blocks():
    oracle = mo.go_index(tiny_lines(), features=3000, sample_threshold=100)
    S = oracle.matrix_f32()
    srch = make_search(S)
    srch.sparse_exact = False                        # force the tensor-core path on this sparse index
    srch.BATCH_BLOCK_ROWS = 4096                     # two row blocks + merge, and thousands of ties at distance 0
    rows = np.array([0, 2040, 6595, 5, 4000, 17] * 12)
    B = S[rows]
    wi, wd = srch.exact_search_batch(B, 20, tensor_cores=False)
    for gi, gd in srch.search_batches([B, B.astype(np.float64)], 20):
        assert np.array_equal(gi, wi) and np.array_equal(gd, wd)


@pytest.mark.parametrize("block_rows", [256, 1024, 4096])
def test_many_row_blocks_in_one_call_refine_thresholds(block_rows):
    """One morna_knn_batched call over several internal row blocks: thresholds tightened and candidate lists
    compacted between blocks; results stay those of the exact scan, ties and zero rows included."""
    from morna_b200 import _lib
    lib = _lib.load()
    try:
        assert lib.morna_debug_set_tuning(9, block_rows) == 0
        rng = np.random.default_rng(block_rows)
        n, d, nq, k = 9000, 96, 300, 40
        S = rng.standard_normal((n, d)).astype(np.float32)
        S[100:104] = S[99]                      # exact ties across the list
        S[8000] = S[99]                         # ... and across blocks
        S[n // 2] = 0.0
        srch = make_search(S)
        rows = rng.permutation(n)[:nq]
        Q = S[rows].astype(np.float64)
        Q[1::3] += 0.05 * rng.standard_normal((len(Q[1::3]), d))
        Q[0] = S[99]
        q = torch.from_numpy(Q).cuda()
        e_ids, e_d = srch.exact_search_device(q, k)
        b_ids, b_d = srch.batched_search_device(q, k)
        assert srch.last_stats[0] == 0
        assert torch.equal(b_ids, e_ids) and torch.equal(b_d, e_d)
    finally:
        lib.morna_debug_set_tuning(9, 131072)


def test_headline_shape_properties():
    """BASELINE configs[2] at full size (50,000 x 3000, 4096 in-index queries, k = 100): properties that do not
    need the oracle -- every query finds itself first at distance exactly 0, lists are ordered under the reference
    rule with distinct ids, no query overflows on Gaussian data -- plus bit equality with the exact scan and with the
    single-query kernel on a sample of the queries."""
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
    n, d, nq, k = 50000, 3000, 4096, 100
    S = torch.randn((n, d), generator=g, device="cuda")
    srch = make_search(S)
    rows = torch.randperm(n, generator=torch.Generator().manual_seed(99))[:nq].cuda()
    q = S[rows].double()
    ids, dist = srch.batched_search_device(q, k)
    assert srch.last_stats[0] == 0
    assert torch.equal(ids[:, 0].long(), rows) and float(dist[:, 0].abs().max()) == 0.0
    assert bool((dist[:, 1:] >= dist[:, :-1]).all())
    tie = dist[:, 1:] == dist[:, :-1]
    assert bool((ids[:, 1:][tie] < ids[:, :-1][tie]).all())              # equal distances: larger id first
    srt = ids.sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all()) and int(ids.min()) >= 0 and int(ids.max()) < n
    pick = torch.arange(0, nq, 97, device="cuda")
    e_ids, e_d = srch.exact_search_device(q[pick], k, allow_single=False)
    assert torch.equal(ids[pick], e_ids) and torch.equal(dist[pick], e_d)
    for j in pick[:5].tolist():
        s_ids, s_d = srch.single_search_device(q[j], k)
        assert torch.equal(s_ids[0], ids[j]) and torch.equal(s_d[0], dist[j])
    # rows-sharded: the union of two shards' exact top-k merged equals the unsharded answer
    from morna_b200 import dist as mdist
    parts = []
    for r in range(2):
        sh = make_search(S, shard=(r, 2))
        parts.append(sh.batched_search_device(q[pick], k))
    m_ids, m_d = mdist.merge_topk(torch.cat([p[0] for p in parts], 1), torch.cat([p[1] for p in parts], 1), k)
    assert torch.equal(m_ids, ids[pick]) and torch.equal(m_d, dist[pick])


@pytest.mark.parametrize("k", [1, 257, 512])
def test_large_and_small_k_on_every_path(k):
    rng = np.random.default_rng(k)
    n, d, nq = 12000, 160, 150
    S = rng.standard_normal((n, d)).astype(np.float32)
    srch = make_search(S)
    Q = S[rng.permutation(n)[:nq]].astype(np.float64) + 0.03 * rng.standard_normal((nq, d))
    q = torch.from_numpy(Q).cuda()
    e_ids, e_d = srch.exact_search_device(q, k, allow_single=False)
    b_ids, b_d = srch.batched_search_device(q, k)
    assert torch.equal(b_ids, e_ids) and torch.equal(b_d, e_d)
    for j in (0, 77):
        s_ids, s_d = srch.single_search_device(q[j], k)
        assert torch.equal(s_ids[0], e_ids[j]) and torch.equal(s_d[0], e_d[j])
    true_d = c_oracle.distances(S, Q[3])
    check_topk(true_d, b_ids[3].cpu().numpy(), b_d[3].cpu().numpy(), tol=1e-9)


def _dense_twin(srch):
    """The same index without its CSR form: every search goes through the dense kernels."""
    from morna_b200.search import MornaSearch
    twin = MornaSearch.__new__(MornaSearch)
    twin.__dict__.update(srch.__dict__)
    twin.csr, twin._ws_pool = None, {}
    return twin


def test_sparse_exact_path_is_bit_identical_on_the_reference_fixture():
    """tests/tiny_intropolis.tsv through morna index: 6850 rows with 1-3 non-zero buckets, 3737 of them tie at distance
    0 from a typical query.  The CSR exact path (nnz multiply-adds per pair) returns the dense scan's bits and the
    oracle's golden lists; batches, single queries and the streaming API all take it."""
    oracle = mo.go_index(tiny_lines(), features=3000, sample_threshold=100)
    S = oracle.matrix_f32()
    srch = make_search(S)
    assert srch.csr is not None and srch.sparse_exact and srch.largest_tie_group > 1000
    dense = _dense_twin(srch)
    rng = np.random.default_rng(3)
    rows = np.concatenate([[0, 2040, 6595, 5, 4000, 17], rng.permutation(S.shape[0])[:250]])
    Q = S[rows].astype(np.float64)
    Q[10:20] += 0.01 * rng.standard_normal((10, 3000))          # dense, out-of-index queries
    Q[20, :] = 0.0
    q = torch.from_numpy(Q).cuda()
    for k in (1, 20, 100, 600):
        s_ids, s_d = srch.batched_search_device(q, k)
        d_ids, d_d = dense.exact_search_device(q, k, allow_single=False)
        assert torch.equal(s_ids, d_ids) and torch.equal(s_d, d_d)
    one_i, one_d = srch.exact_search_device(q[:1], 20)
    with open(os.path.join(GOLDEN, "tiny_expected.json")) as fh:
        exp = json.load(fh)
    assert one_i[0].tolist() == exp["queries"][0]["ids"]
    got = list(srch.search_batches([Q, Q.astype(np.float32)], 20))
    w_i, w_d = dense.exact_search_device(q, 20, allow_single=False)
    assert np.array_equal(got[0][0], w_i.cpu().numpy()) and np.array_equal(got[0][1], w_d.cpu().numpy())


@pytest.mark.parametrize("tile_mb", [1, 3, 1700])
def test_sparse_exact_path_query_tiles_of_any_size_agree(tile_mb):
    """The CSR exact path walks the queries in tiles sized by its distance scratch (one tile by default since the last
    session of round 2): many small tiles (unequal last tile included), a few, and one give the same ids and distances
    as the dense scan -- tie-heavy fixture rows, so the selection's tie-ranking and barrier-free phases both run."""
    from morna_b200 import _lib
    lib = _lib.load()
    oracle = mo.go_index(tiny_lines(), features=3000, sample_threshold=100)
    S = oracle.matrix_f32()
    srch = make_search(S)
    assert srch.csr is not None
    dense = _dense_twin(srch)
    rng = np.random.default_rng(11)
    Q = S[rng.permutation(S.shape[0])[:173]].astype(np.float64)
    Q[5:9] += 0.02 * rng.standard_normal((4, 3000))
    q = torch.from_numpy(Q).cuda()
    try:
        assert lib.morna_debug_set_tuning(35, tile_mb) == 0
        for k in (7, 100):
            s_ids, s_d = srch.exact_search_device(q, k)
            d_ids, d_d = dense.exact_search_device(q, k, allow_single=False)
            assert torch.equal(s_ids, d_ids) and torch.equal(s_d, d_d)
    finally:
        lib.morna_debug_set_tuning(35, 1700)


@pytest.mark.parametrize("n,d,max_nnz", [(5000, 3000, 3), (3000, 130, 16), (2000, 7, 4), (4000, 40000, 8)])
def test_sparse_exact_path_matches_dense_kernels_bit_for_bit(n, d, max_nnz):
    """Synthetic sparse rows with up to 16 non-zeros (several per summation lane, negative zeros, float32 denormals,
    all-zero rows, wide feature counts): CSR exact search == dense exact search, ids and distance bits, and one query
    against the oracle."""
    rng = np.random.default_rng(n + d)
    S = np.zeros((n, d), np.float32)
    for i in range(n):
        m = int(rng.integers(0, min(max_nnz, d) + 1))
        cols = rng.choice(d, size=m, replace=False)
        if m and rng.random() < 0.3:                                   # several entries in one lane class
            cols = (cols[0] // 4 * 4 + 128 * np.arange(m)) % d
            cols = np.unique(cols)
        S[i, cols] = (rng.integers(1, 50, size=len(cols)) * rng.choice([-1.0, 1.0], size=len(cols))
                      * rng.choice([0.098, 1.21, 1.415, 0.5], size=len(cols))).astype(np.float32)
    S[3, :2] = np.float32(1e-42)
    S[4, 0] = np.float32(-0.0)
    srch = make_search(S)
    assert srch.csr is not None
    dense = _dense_twin(srch)
    nq = 97
    Q = np.concatenate([S[rng.permutation(n)[:60]].astype(np.float64), rng.standard_normal((nq - 60, d))])
    Q[1] = 0.0
    Q[2, : min(d, 5)] = -0.0
    q = torch.from_numpy(Q).cuda()
    for k in (1, 33):
        s_ids, s_d = srch.exact_search_device(q, k)
        d_ids, d_d = dense.exact_search_device(q, k, allow_single=False)
        assert torch.equal(s_ids, d_ids) and torch.equal(s_d, d_d)
    true_d = c_oracle.distances(S, Q[70])
    check_topk(true_d, s_ids[70].cpu().numpy(), s_d[70].cpu().numpy(), tol=1e-9)
    # a non-finite query value keeps the dense arithmetic (0 * inf = nan there)
    Q[5, 0] = np.inf
    q = torch.from_numpy(Q[:8]).cuda()
    s_ids, s_d = srch.exact_search_device(q, 5)
    d_ids, d_d = dense.exact_search_device(q, 5, allow_single=False)
    assert torch.equal(s_ids, d_ids)


def test_approximate_mode_recall_and_distances():
    """MornaSearch.approx_search_device / search_nn: the tensor-core pass without the exact re-rank.  Recall@k against the
    exact answer stays high (ids differ only where cosines are closer than the fp16 error), distances are within the
    fp16 error of the true ones, well-separated data is answered exactly."""
    n, d, nq, k = 20000, 1000, 500, 50
    S = __import__("morna_b200.synth", fromlist=["x"]).gauss(n, d, "cuda", seed=21)
    srch = make_search(S)
    from morna_b200 import synth
    q, rows = synth.queries(S, nq, noise=0.05)
    e_ids, e_d = srch.batched_search_device(q, k)
    a_ids, a_d = srch.approx_search_device(q, k)
    hits = sum(len(set(a_ids[j].tolist()) & set(e_ids[j].tolist())) for j in range(nq))
    recall = hits / float(nq * k)
    assert recall >= 0.95, recall
    S64 = S.double()
    cos = (q @ S64.t()) / (q.norm(dim=1, keepdim=True) * S64.norm(dim=1)[None, :])
    true_d = (2 - 2 * cos).clamp_min(0).sqrt()
    err = (a_d - torch.gather(true_d, 1, a_ids.long())).abs().max()
    assert float(err) < 2e-3
    assert bool((a_d[:, 1:] >= a_d[:, :-1]).all())
    print("approximate mode: recall@%d = %.4f, max distance error %.2e" % (k, recall, float(err)))
    T = synth.tissue(8000, 600, "cuda", seed=5)
    ts = make_search(T)
    tq, _ = synth.queries(T, 200, noise=0.05)
    ta, _ = ts.approx_search_device(tq, 5)
    te, _ = ts.batched_search_device(tq, 5)
    assert float((ta == te).float().mean()) > 0.97
    # search_nn on the current query sample (what `morna.py search` without -e calls)
    srch.query_sample = q[0].cpu().tolist()
    ids, dists = srch.search_nn(k)
    assert ids == a_ids[0].tolist() and len(dists) == k


def test_streaming_batches_with_k_above_the_tensor_path_limit():
    """k = 600 > 512 sends every pipeline slot through the FP64 scan on the slot's own stream; each stream has its own
    scan workspace (two batches in flight used to share one), so depth 2 and 3 return the synchronous call's lists."""
    rng = np.random.default_rng(600)
    n, d, k = 5000, 96, 600
    S = rng.standard_normal((n, d)).astype(np.float32)
    srch = make_search(S)
    batches = [S[rng.permutation(n)[:nq]] + np.float32(0.02) * rng.standard_normal((nq, d)).astype(np.float32)
               for nq in (130, 130, 77, 130, 130, 9)]
    want = [srch.exact_search_batch(B, k, tensor_cores=False) for B in batches]
    for depth in (2, 3):
        got = [(i.copy(), d_.copy()) for i, d_ in srch.search_batches(iter(batches), k, depth=depth)]
        for (gi, gd), (wi, wd) in zip(got, want):
            assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
    # k beyond the select kernel's 2048 and beyond the number of rows: every row comes back, in order
    ids, dist = srch.exact_search_batch(batches[0][:3], n + 50, tensor_cores=False)
    assert ids.shape == (3, n + 50) and bool((ids[:, n:] == -1).all()) and bool(np.isinf(dist[:, n:]).all())
    assert sorted(ids[0, :n].tolist()) == list(range(n)) and bool((np.diff(dist[0, :n]) >= 0).all())
    true_d = c_oracle.distances(S, batches[0][0].astype(np.float64))
    assert np.abs(np.sort(true_d) - dist[0, :n]).max() < 1e-9
    assert srch.exact_search_batch(batches[0][:2], 0, tensor_cores=False)[0].shape == (2, 0)
