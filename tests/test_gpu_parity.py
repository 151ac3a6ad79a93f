"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  GPU only."""
import ctypes
import hashlib
import io
import json
import math
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import c_oracle
from oracle import morna_oracle as mo
from tests.helpers import GOLDEN, check_order_rule, check_topk, load_reference_vectors, tiny_lines

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from morna_b200 import _lib
    assert torch.cuda.is_available(), "these tests need the B200"
    return _lib.load()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------ K1 hashing
def test_hash_kernel_bit_exact(lib):
    from morna_b200 import _lib
    rng = np.random.default_rng(3)
    keys = ["", "a", "ab", "abc", "abcd", "abcde", "chr1 14830 14929", "chr1 14830 14969", "chr1 15039 15795"]
    for _ in range(5000):
        chrom = "chr" + str(rng.choice(list(range(1, 23)) + ["X", "Y"]))
        start = int(rng.integers(1, 2.4e8))
        keys.append("%s %d %d" % (chrom, start, start + int(rng.integers(50, 5e5))))
    for dim in (3000, 500, 30000, 7, 1):
        raw_o, bucket_o, sign_o = c_oracle.hash_rows(keys, dim)
        blobs = [k.encode() for k in keys]
        off = np.zeros(len(keys) + 1, np.int32)
        off[1:] = np.cumsum([len(b) for b in blobs])
        packed = np.frombuffer(b"".join(blobs) + b"\0", np.uint8).copy()
        raw = torch.empty(len(keys), dtype=torch.int32, device="cuda")
        bucket = torch.empty_like(raw)
        sign = torch.empty(len(keys), dtype=torch.int8, device="cuda")
        d_packed, d_off = dev(packed), dev(off)       # keep the inputs alive across the launch
        rc = lib.morna_hash_junctions(_lib.dev_ptr(d_packed), _lib.dev_ptr(d_off), len(keys), dim,
                                      _lib.dev_ptr(raw), _lib.dev_ptr(bucket), _lib.dev_ptr(sign), _lib.stream_ptr())
        assert rc == 0
        assert np.array_equal(raw.cpu().numpy(), raw_o)
        assert np.array_equal(bucket.cpu().numpy(), bucket_o)
        assert np.array_equal(sign.cpu().numpy(), sign_o)
    assert raw_o[6] == -28859081


# ------------------------------------------------------------------ index build
def build_both(lines, features, sample_count, threshold, store_skipped=True):
    from morna_b200.index import MornaIndex
    oracle = mo.go_index(lines, features=features, sample_count=sample_count, sample_threshold=threshold)
    idx = MornaIndex(oracle.sample_count, "unused", dim=features, sample_threshold=threshold,
                     store_skipped_rows=store_skipped)
    idx.add_lines(lines)
    idx.build(n_trees=20)
    return oracle, idx


@pytest.mark.parametrize("case_no", [0, 1, 2])
def test_index_build_reference_unittest_inputs(case_no):
    vec = load_reference_vectors()
    case = vec["cases"][case_no]
    lines = [l + "\n" for l in vec[case["input"]]]
    for features in (case["features"], 40, 7):
        oracle, idx = build_both(lines, features, case["sample_count"], case["sample_threshold"])
        assert idx.get_n_items() == oracle.new_internal_id == case["n_items"]
        assert idx.internal_id_map == oracle.internal_id_map
        assert idx.skipped == oracle.skipped and idx.junc_id == oracle.junc_id
        assert dict(idx.sample_frequencies) == dict(oracle.sample_frequencies)
        assert np.array_equal(idx.accumulator_f64(), oracle.matrix_f64())     # bit-exact doubles
        assert np.array_equal(idx.matrix_f32(), oracle.matrix_f32())


def test_index_build_tiny_fixture_bit_exact():
    with open(os.path.join(GOLDEN, "tiny_expected.json")) as fh:
        exp = json.load(fh)
    oracle, idx = build_both(tiny_lines(), 3000, None, 100)
    assert idx.get_n_items() == exp["n_kept"] == 6850
    assert idx.internal_id_map == oracle.internal_id_map
    raw, bucket, sign = (t.cpu().numpy() for t in idx.row_hash)
    for j, row in enumerate(exp["rows"]):
        assert (int(raw[j]), int(bucket[j]), int(sign[j])) == (row["hash"], row["bucket"], row["sign"])
    S = idx.matrix_f32()
    assert hashlib.sha256(np.ascontiguousarray(S).tobytes()).hexdigest() == exp["matrix_f32_sha256"]
    assert np.array_equal(idx.accumulator_f64(), oracle.matrix_f64())


def synthetic_rows(rng, n_rows, n_samples, dup_keys=True, descending=False):
    lines = []
    for j in range(n_rows):
        if dup_keys and j % 7 == 3 and j > 10:
            key = lines[int(rng.integers(0, j))].split("\t")[:3]
        else:
            start = int(rng.integers(1, 5000))
            key = ["chr%d" % rng.integers(1, 4), str(start), str(start + int(rng.integers(1, 300)))]
        size = int(min(n_samples, max(1, rng.lognormal(3.0, 1.5))))
        samples = np.sort(rng.choice(n_samples, size=size, replace=False) + 1)
        if descending:
            samples = samples[::-1]
        covs = 1 + rng.geometric(0.5, size=size) * rng.integers(1, 40, size=size)
        lines.append("\t".join(key + ["+", "GT", "AG", ",".join(map(str, samples)), ",".join(map(str, covs))]) + "\n")
    return lines


@pytest.mark.parametrize("features,threshold,n_samples", [(13, 3, 300), (500, 20, 2000), (3000, 1, 50), (64, 40, 30000)])
def test_index_build_synthetic_bit_exact(features, threshold, n_samples):
    rng = np.random.default_rng(features + threshold)
    lines = synthetic_rows(rng, 400, n_samples, descending=(features == 500))
    oracle, idx = build_both(lines, features, None, threshold)
    assert idx.internal_id_map == oracle.internal_id_map
    assert idx.get_n_items() == oracle.new_internal_id
    assert np.array_equal(idx.accumulator_f64(), oracle.matrix_f64())
    assert np.array_equal(idx.matrix_f32(), oracle.matrix_f32())
    # pad columns of the device rows are zero
    if idx.ld > idx.dim:
        assert float(idx.vectors[:, idx.dim:].abs().max()) == 0.0


def test_index_build_rows_with_repeated_or_unsorted_samples():
    """A row that is not monotonic (or lists a sample twice) takes the atomic variant of the scatter-add."""
    from morna_b200.index import MornaIndex
    rng = np.random.default_rng(33)
    lines = synthetic_rows(rng, 120, 400)
    lines.insert(5, "chr9\t5\t9\t+\tGT\tAG\t7,3,9,3,250,1\t2,4,1,8,5,1\n")       # sample 3 twice, unsorted
    lines.append("chr9\t50\t90\t+\tGT\tAG\t5,9,2,77\t1,1,1,1\n")                  # unsorted, distinct
    oracle = mo.go_index(lines, features=50, sample_threshold=2)
    idx = MornaIndex(oracle.sample_count, "unused", dim=50, sample_threshold=2)
    idx.add_lines(lines)
    idx.build()
    assert idx.internal_id_map == oracle.internal_id_map
    got, want = idx.accumulator_f64(), oracle.matrix_f64()
    np.testing.assert_allclose(got, want, rtol=1e-15, atol=0)      # repeated sample: last-ulp order freedom only
    assert np.array_equal(idx.matrix_f32(), oracle.matrix_f32())


@pytest.mark.parametrize("late_row", [10, 100, 2000])
def test_id_pass_early_exit_keeps_the_first_seen_order(late_row):
    """The id pass checks after 1/64 and 1/8 of the rows whether every sample id has been met and skips the rest if so.
    One sample is held back until `late_row` (inside the first stretch, the second, or near the end); ids 0 and 7 never
    pass the threshold at all in the last case's variant (an id space with holes never exits early)."""
    from morna_b200 import _lib
    from morna_b200.index import MornaIndex
    lib = _lib.load()
    rng = np.random.default_rng(late_row)
    n_rows, n_samples = 2048, 300
    rows = []
    for j in range(n_rows):
        pool = n_samples if j >= late_row else n_samples - 1
        samples = rng.choice(pool, size=40, replace=False)
        if j == late_row:
            samples[0] = n_samples - 1
        if j % 3:
            samples = np.sort(samples)
        rows.append(("chr1 %d %d" % (j, j + 9), samples.tolist(), rng.integers(1, 9, size=40).tolist()))
    want, order = {}, 0
    for _, samples, _ in rows:                              # morna.py:377-382
        for sid in samples:
            if sid not in want:
                want[sid] = order
                order += 1
    maps = []
    for early in (1, 0):
        lib.morna_debug_set_tuning(25, early)
        idx = MornaIndex(n_samples, "unused", dim=64, sample_threshold=1)
        for row in rows:
            idx.add_junction(*row)
        idx.build()
        maps.append(idx.internal_id_map)
    lib.morna_debug_set_tuning(25, 1)
    assert maps[0] == maps[1] == want


def test_index_build_id_range_shards_concatenate():
    from morna_b200.index import MornaIndex
    rng = np.random.default_rng(21)
    lines = synthetic_rows(rng, 300, 500)
    oracle = mo.go_index(lines, features=100, sample_threshold=5)
    parts = []
    n = oracle.new_internal_id
    for lo, hi in ((0, n // 3), (n // 3, n // 2), (n // 2, n)):
        idx = MornaIndex(oracle.sample_count, "unused", dim=100, sample_threshold=5)
        idx.add_lines(lines)
        idx.build(id_range=(lo, hi))
        assert idx.internal_id_map == oracle.internal_id_map
        parts.append(idx.matrix_f32())
    assert np.array_equal(np.concatenate(parts), oracle.matrix_f32())


def test_index_build_no_passing_rows_raises():
    from morna_b200.index import MornaIndex
    idx = MornaIndex(10, "unused", dim=40, sample_threshold=50)
    idx.add_junction("chr1 1 2", [1, 2, 3], [1, 1, 1])
    with pytest.raises(ValueError):
        idx.build()
    idx = MornaIndex(10, "unused", dim=40, sample_threshold=50, store_skipped_rows=True)
    idx.add_junction("chr1 1 2", [1, 2, 3], [1, 1, 1])
    with pytest.raises(ValueError):
        idx.build()


# ------------------------------------------------------------------ exact search
def make_search(S, **kw):
    from morna_b200.search import MornaSearch
    return MornaSearch(vectors=S, stats=(S.shape[0], S.shape[0], S.shape[1]), **kw)


def test_distances_match_oracle_and_self_distance_is_zero(lib):
    from morna_b200 import _lib
    rng = np.random.default_rng(8)
    for n, d in ((257, 3000), (100, 37), (33, 4), (1000, 130)):
        S = rng.standard_normal((n, d)).astype(np.float32)
        S[5] = 0.0
        S[7] = S[3] * 2
        srch = make_search(S)
        Q = np.concatenate([S[:9].astype(np.float64), rng.standard_normal((4, d))])
        Q[10] = 0.0
        dq = dev(Q)
        dist = torch.empty((Q.shape[0], n), dtype=torch.float64, device="cuda")
        rc = lib.morna_angular_distances(_lib.dev_ptr(srch.vectors), _lib.dev_ptr(srch.pp), n, d, srch.ld,
                                         _lib.dev_ptr(dq), Q.shape[0], d, _lib.dev_ptr(dist), n, _lib.stream_ptr())
        assert rc == 0
        got = dist.cpu().numpy()
        for qi in range(Q.shape[0]):
            want = c_oracle.distances(S, Q[qi])
            np.testing.assert_allclose(got[qi], want, rtol=0, atol=2e-7 if qi in (3, 7) else 1e-12)
        for i in range(9):
            if i != 5:
                assert got[i, i] == 0.0                      # stored row vs itself: exactly 0
        assert np.all(got[5] == math.sqrt(2.0)) and np.all(got[10] == math.sqrt(2.0))   # zero norm -> sqrt(2)
        assert np.all(got[:, 5] == math.sqrt(2.0))


def run_select(lib, keys, ids, k, id_base=0):
    from morna_b200 import _lib
    nq, n = keys.shape
    out_i = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    out_d = torch.empty((nq, k), dtype=torch.float64, device="cuda")
    ws = _lib.workspace(lib.morna_select_topk_workspace_bytes(n, nq, k), "cuda")
    dk = dev(keys)
    di = dev(ids) if ids is not None else None
    rc = lib.morna_select_topk(_lib.dev_ptr(dk), _lib.dev_ptr(di) if di is not None else None, n, n, id_base, nq, k,
                               _lib.dev_ptr(out_i), _lib.dev_ptr(out_d), _lib.dev_ptr(ws), ws.numel(),
                               _lib.stream_ptr())
    assert rc == 0
    return out_i.cpu().numpy(), out_d.cpu().numpy()


@pytest.mark.parametrize("n", [1, 5, 100, 1023, 1024, 1025, 8191, 8192, 8193, 21504, 100003])
def test_select_topk_exact_order(lib, n):
    rng = np.random.default_rng(n)
    for k in (1, 20, 100, 1000, 2048):
        nq = 3
        keys = rng.random((nq, n))
        keys[1] = np.round(keys[1] * 8) / 8            # heavy ties
        keys[2, : n // 2] = 0.0                        # half of the rows tie at distance 0
        ids_out, d_out = run_select(lib, keys, None, k, id_base=1000)
        for q in range(nq):
            want_i, want_d = mo.topk_rule(keys[q], k, ids=np.arange(n) + 1000)
            m = len(want_i)
            assert np.array_equal(ids_out[q, :m], want_i), (n, k, q)
            assert np.array_equal(d_out[q, :m], want_d)
            assert np.all(ids_out[q, m:] == -1) and np.all(np.isinf(d_out[q, m:]))


def test_select_topk_explicit_ids_with_padding(lib):
    rng = np.random.default_rng(4)
    n, k = 800, 100                                   # e.g. 8 ranks x 100 gathered entries
    keys = np.round(rng.random((2, n)) * 50) / 50
    ids = rng.permutation(10 * n)[: 2 * n].reshape(2, n).astype(np.int32)
    ids[0, 700:] = -1
    keys[0, 700:] = np.inf
    got_i, got_d = run_select(lib, keys, ids, k)
    for q in range(2):
        valid = ids[q] >= 0
        want_i, want_d = mo.topk_rule(keys[q][valid], k, ids=ids[q][valid])
        assert np.array_equal(got_i[q], want_i) and np.array_equal(got_d[q], want_d)


def test_exact_search_matches_oracle_gaussian():
    rng = np.random.default_rng(12)
    n, d, k = 3000, 3000, 100
    S = rng.standard_normal((n, d)).astype(np.float32)
    srch = make_search(S)
    qrows = rng.permutation(n)[:6]
    Q = np.concatenate([S[qrows].astype(np.float64), S[qrows[:3]] + 0.05 * rng.standard_normal((3, d))])
    ids, dist = srch.exact_search_batch(Q, k)
    for qi in range(Q.shape[0]):
        true_d = c_oracle.distances(S, Q[qi])
        want_i, want_d = c_oracle.exact_search(S, Q[qi], k)
        check_topk(true_d, ids[qi], dist[qi], tol=1e-9, dist_tol=1e-5)
        check_order_rule(ids[qi], dist[qi])
        assert np.array_equal(ids[qi], want_i)         # no near-ties in Gaussian data: identical ids
        np.testing.assert_allclose(dist[qi], want_d, rtol=0, atol=1e-9)


def test_exact_search_tiny_fixture_matches_golden():
    with open(os.path.join(GOLDEN, "tiny_expected.json")) as fh:
        exp = json.load(fh)
    oracle = mo.go_index(tiny_lines(), features=3000, sample_threshold=100)
    S = oracle.matrix_f32()
    srch = make_search(S)
    for q in exp["queries"]:
        query = S[q["internal_id"]].astype(np.float64)
        ids, dist = srch.exact_search_batch(query[None, :], 20)
        true_d = c_oracle.distances(S, query)
        check_topk(true_d, ids[0], dist[0], tol=1e-7, dist_tol=1e-5)
        check_order_rule(ids[0], dist[0])
    q0 = exp["queries"][0]       # rows with <= 2 non-zeros: sums are bit-equal, so ids are too
    ids, dist = srch.exact_search_batch(S[q0["internal_id"]].astype(np.float64)[None, :], 20)
    assert ids[0].tolist() == q0["ids"] and dist[0].tolist() == q0["dists"]


def test_exact_search_k_larger_than_n_and_shards():
    rng = np.random.default_rng(2)
    S = rng.standard_normal((50, 40)).astype(np.float32)
    srch = make_search(S)
    ids, dist = srch.exact_search_batch(S[:2].astype(np.float64), 64)
    assert np.all(ids[:, 50:] == -1) and np.all(np.isinf(dist[:, 50:]))
    want_i, _ = c_oracle.exact_search(S, S[0].astype(np.float64), 64)
    assert np.array_equal(ids[0, :50], want_i)
    # two row shards + merge == one shard
    from morna_b200 import dist as mdist
    parts = [make_search(S, shard=(r, 2)) for r in range(2)]
    q = torch.from_numpy(S[:5].astype(np.float64)).cuda()
    lists = [p.exact_search_device(q, 10) for p in parts]
    mi, md = mdist.merge_topk(torch.cat([l[0] for l in lists], 1), torch.cat([l[1] for l in lists], 1), 10)
    full_i, full_d = srch.exact_search_device(q, 10)
    assert torch.equal(mi, full_i) and torch.equal(md, full_d)


# ------------------------------------------------------------------ CLI end to end (BASELINE config 1)
def test_cli_index_then_exact_search_of_in_index_sample(tmp_path):
    from morna_b200 import cli, files
    with open(os.path.join(GOLDEN, "tiny_expected.json")) as fh:
        exp = json.load(fh)
    base = str(tmp_path / "tiny")
    out = io.StringIO()
    assert cli.main(["index", "--intropolis", os.path.join(GOLDEN, "tiny_intropolis.tsv"), "-x", base,
                     "--features", "3000"], stdout=out) == 0
    assert files.read_stats(base) == (6850, 6850, 3000)
    assert files.read_freq(base)["chr1 14830 14929"] == 2040
    assert files.read_map(base)[21504] == 6595
    S = files.read_vectors(base)
    assert hashlib.sha256(np.ascontiguousarray(S).tobytes()).hexdigest() == exp["matrix_f32_sha256"]
    out = io.StringIO()
    assert cli.main(["search", "-x", base, "-q", "12", "-e", "-d", "-r", "20"], stdout=out) == 0
    lines = out.getvalue().splitlines()
    assert lines[0] == "querying by sample id 12" and lines[1] == "this is internal id 0"
    got = [l.split("\t") for l in lines[2:]]
    assert [int(r[1]) for r in got] == exp["queries"][0]["ids"]
    assert [r[2] for r in got] == ["0.0"] * 20 and got[0][0] == "1."
    # out-of-index query from a raw junction list
    out = io.StringIO()
    query = "chr1\t14830\t14929\t3\nchr1\t14830\t14969\t5\nchr9\t1\t2\t100\n"
    assert cli.main(["search", "-x", base, "-f", "raw", "-e", "-d", "-r", "5"], stdin=io.StringIO(query), stdout=out) == 0
    freq = files.read_freq(base)
    qvec = mo.finalize_query({("chr1", 14830, 14929): 3, ("chr1", 14830, 14969): 5}, freq, 6850, 3000)
    want_i, want_d = c_oracle.exact_search(np.asarray(S), np.asarray(qvec), 5)
    got = [l.split("\t") for l in out.getvalue().splitlines()]
    true_d = c_oracle.distances(np.asarray(S), np.asarray(qvec))
    check_topk(true_d, [int(r[1]) for r in got], [float(r[2]) for r in got], tol=1e-7, dist_tol=1e-5)
    with pytest.raises(ValueError):
        cli.main(["search", "-x", base, "-q", "999999", "-e"], stdout=io.StringIO())
    # -m: index with a metadata file, search joins every result's sample id against basename.meta.mor (morna.py:494-520, 666-676)
    meta = tmp_path / "meta.tsv"
    inv = {v: k for k, v in files.read_map(base).items()}
    top = [inv[i] for i in exp["queries"][0]["ids"][:3]]
    meta.write_text("".join("%d\ttissue_%d blood\n" % (sid, sid) for sid in top[:2]) + "12 the query itself\n")
    assert cli.main(["index", "--intropolis", os.path.join(GOLDEN, "tiny_intropolis.tsv"), "-x", base, "-m", str(meta)],
                    stdout=io.StringIO()) == 0
    out = io.StringIO()
    assert cli.main(["search", "-x", base, "-q", "12", "-e", "-m", "-r", "3"], stdout=out) == 0
    rows = [l.split("\t") for l in out.getvalue().splitlines()[2:]]
    assert [int(r[1]) for r in rows] == exp["queries"][0]["ids"][:3]
    assert rows[0][2] == "(u'tissue_%d blood\\n',)" % top[0] and rows[1][2] == "(u'tissue_%d blood\\n',)" % top[1]
    assert rows[2][2] == "None"                                      # no metadata line for that sample


# ------------------------------------------------------------------ single query, fused FP64 scan + select
@pytest.mark.parametrize("n,d,k", [(21504, 3000, 100), (3000, 3000, 20), (5000, 130, 100), (900, 37, 64), (40, 16, 100)])
def test_single_query_path_equals_fp64_scan(n, d, k):
    _run_single_query_checks(n, d, k)


@pytest.mark.parametrize("rows_per_pass", [1, 2, 3, 4, 5])
def test_single_query_rows_per_pass_variants(rows_per_pass):
    from morna_b200 import _lib
    lib = _lib.load()
    try:
        assert lib.morna_debug_set_tuning(3, rows_per_pass) == 0
        _run_single_query_checks(7001, 300, 100)
    finally:
        lib.morna_debug_set_tuning(3, 0)


def test_single_query_many_rows_per_warp_and_huge_query_values():
    _run_single_query_checks(60011, 64, 100)          # more rows than one pass of every resident warp
    rng = np.random.default_rng(1)
    S = rng.standard_normal((900, 40)).astype(np.float32)
    srch = make_search(S)
    qv = rng.standard_normal(40)
    qv[3] = 2e38                                      # q * 2^896 is not finite: the generic scan answers
    q = torch.from_numpy(qv).cuda()
    s_ids, s_d = srch.single_search_device(q, 10)
    assert int(srch._sfallback.item()) == 1
    e_ids, e_d = srch.exact_search_device(q.view(1, -1), 10, allow_single=False)
    assert torch.equal(s_ids, e_ids) and torch.equal(s_d, e_d)
    s_ids, s_d = srch.single_search_device(torch.from_numpy(S[5].astype(np.float64)).cuda(), 10)
    assert int(srch._sfallback.item()) == 0 and int(s_ids[0, 0]) == 5      # the workspace is reusable afterwards


def _run_single_query_checks(n, d, k):
    rng = np.random.default_rng(n + d + k)
    S = rng.standard_normal((n, d)).astype(np.float32) * np.exp(rng.standard_normal((n, 1))).astype(np.float32)
    S[n // 2] = 0.0
    srch = make_search(S)
    qs = [S[7].astype(np.float64), S[n - 1] + 0.03 * rng.standard_normal(d), rng.standard_normal(d) * 1e-3]
    for qv in qs:
        q = torch.from_numpy(qv).cuda()
        s_ids, s_d = srch.single_search_device(q, k)
        e_ids, e_d = srch.exact_search_device(q.view(1, -1), k, allow_single=False)
        assert int(srch._sfallback.item()) == 0
        assert torch.equal(s_ids, e_ids) and torch.equal(s_d, e_d)
    true_d = c_oracle.distances(S, qs[-1])
    m = min(k, n)
    check_topk(true_d, s_ids[0].cpu().numpy()[:m], s_d[0].cpu().numpy()[:m], tol=1e-9)


@pytest.mark.parametrize("chain", [0, 1, 2, 3])
def test_single_query_stream_equals_one_query_at_a_time(chain):
    """morna_knn_single_stream (back-to-back kernels, programmatic dependent launch, alternating workspaces) returns for
    every query what morna_knn_single returns, which is pinned to the oracle above -- including queries that fall
    back (huge values), many queries in one call (each workspace half reused many times) and repeated calls."""
    from morna_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(17 + chain)
    n, d, k = 9001, 300, 50
    S = rng.standard_normal((n, d)).astype(np.float32) * np.exp(rng.standard_normal((n, 1))).astype(np.float32)
    srch = make_search(S)
    Q = np.concatenate([S[:20].astype(np.float64), rng.standard_normal((21, d))])
    Q[7, 3] = 2e38                                     # answered by the generic scan
    Qd = torch.from_numpy(Q).cuda()
    try:
        assert lib.morna_debug_set_tuning(29, chain) == 0
        for _ in range(3):
            ids, dist = srch.single_search_stream(Qd, k)
    finally:
        lib.morna_debug_set_tuning(29, 2)
    for j in range(Q.shape[0]):
        s_ids, s_d = srch.single_search_device(Qd[j], k)
        assert torch.equal(ids[j], s_ids[0]) and torch.equal(dist[j], s_d[0]), "query %d" % j
    true_d = c_oracle.distances(S, Q[30])
    check_topk(true_d, ids[30].cpu().numpy(), dist[30].cpu().numpy(), tol=1e-9)
    # k above the row count and an empty batch behave like the other entry points
    small = make_search(S[:30])
    i2, d2 = small.single_search_stream(Qd[:3], 40)
    e2, f2 = small.exact_search_device(Qd[:3], 40, allow_single=False)
    assert torch.equal(i2, e2) and torch.equal(d2, f2)
    i3, _ = srch.single_search_stream(Qd[:0], k)
    assert i3.shape == (0, k)
    if chain == 2:
        # very wide --features: the staged query would not fit beside a second CTA, every query is flagged and the generic
        # scan answers -- same lists
        W = rng.standard_normal((300, 13000)).astype(np.float32)
        wide = make_search(W)
        wq = torch.from_numpy(np.concatenate([W[:2].astype(np.float64), rng.standard_normal((1, 13000))])).cuda()
        i4, d4 = wide.single_search_stream(wq, 5)
        e4, f4 = wide.exact_search_device(wq, 5, allow_single=False)
        assert torch.equal(i4, e4) and torch.equal(d4, f4) and i4[0, 0] == 0 and i4[1, 0] == 1


def test_single_query_ties_fall_back_and_shard_offsets():
    oracle = mo.go_index(tiny_lines(), features=3000, sample_threshold=100)
    S = oracle.matrix_f32()
    srch = make_search(S)
    q = torch.from_numpy(S[2040].astype(np.float64)).cuda()   # sample 1: one non-zero bucket
    s_ids, s_d = srch.single_search_device(q, 20)
    assert int(srch._sfallback.item()) == 1                  # 3737 rows tie at distance 0
    e_ids, e_d = srch.exact_search_device(q.view(1, -1), 20, allow_single=False)
    assert torch.equal(s_ids, e_ids) and torch.equal(s_d, e_d)
    zero = torch.zeros(3000, dtype=torch.float64, device="cuda")
    z_ids, z_d = srch.single_search_device(zero, 5)
    assert z_ids[0].tolist() == [6849, 6848, 6847, 6846, 6845] and float(z_d[0, 0]) == math.sqrt(2.0)
    rng = np.random.default_rng(3)
    G = rng.standard_normal((1001, 64)).astype(np.float32)
    full = make_search(G)
    part = make_search(G, shard=(1, 2))
    qg = torch.from_numpy(G[900].astype(np.float64)).cuda()
    p_ids, p_d = part.single_search_device(qg, 10)
    assert int(p_ids[0, 0]) == 900 and p_ids.min() >= 501


# ------------------------------------------------------------------ native tokenizer feeding the index build
def test_index_from_text_blocks_equals_index_from_python_rows(tmp_path):
    import gzip
    from morna_b200 import parse
    from morna_b200.index import MornaIndex, go_index
    rng = np.random.default_rng(77)
    lines = synthetic_rows(rng, 500, 3000)
    lines.insert(40, " chr7\t5\t9\t+\tGT\tAG\t7,30,90\t2,4,1\n")                    # leading blank: Python's strip()
    lines.insert(90, "chr7\t50\t90\t+\tGT\tAG\t5,9,20,77,78,79\t1,1,1,1\n")           # unequal lists: zip()
    lines.insert(91, "chr7\t51\t91\t+\tGT\tAG\t1,2,3,4,05\t1,1,1,1,1\n")              # "05": its own count_samples string
    text = "".join(lines).encode()
    n_samples = parse.count_samples(lines)
    a = MornaIndex(n_samples, "unused", dim=200, sample_threshold=4)
    a.add_lines(lines)
    a.build()
    for block_bytes in (1 << 30, 5000):
        seen = set()
        b = MornaIndex(0, "unused", dim=200, sample_threshold=4)
        for block in parse.read_blocks(io.BytesIO(text), block_bytes=block_bytes):
            b.add_text(block, seen_samples=seen, n_threads=3)
        assert len(seen) == n_samples
        b.sample_count = len(seen)
        b.build()
        assert b.internal_id_map == a.internal_id_map and b.skipped == a.skipped and b.junc_id == a.junc_id
        assert b.sample_frequencies == a.sample_frequencies
        assert np.array_equal(b.matrix_f32(), a.matrix_f32())
    # and through the command's own entry point, gzipped, sample count discovered in the same pass
    path = str(tmp_path / "rows.tsv.gz")
    with gzip.open(path, "wb") as fh:
        fh.write(text)
    c = go_index(path, str(tmp_path / "idx"), 200, None, None, 4, 1024, False, None, out=io.StringIO(), junction_shards=False)
    assert c.sample_count == n_samples
    assert np.array_equal(c.matrix_f32(), a.matrix_f32())
    # with the junctions-by-sample shards on (the command's default) the row with fewer coverages than samples stops the
    # run as it stops the reference: update_junction_dbs reads coverages[i] for every sample (morna.py:269)
    with pytest.raises(IndexError):
        go_index(path, str(tmp_path / "idx2"), 200, None, None, 4, 1024, False, None, out=io.StringIO())


def test_index_build_properties_at_scale():
    """A few million pairs through the C ABI (too many for the Python oracle): the warp-per-range and the
    barrier-per-row scatter-add give the same doubles; doubling every coverage doubles every cell exactly (powers of
    two commute with each rounding the reference performs); two runs are identical (no atomics on the path)."""
    from morna_b200 import _lib
    from morna_b200.index import MornaIndex
    lib = _lib.load()
    rng = np.random.default_rng(11)
    n_samples, n_rows, dim = 6000, 9000, 500
    lens = np.clip(rng.lognormal(5.0, 1.2, size=n_rows).astype(np.int64), 1, n_samples)
    def build(scale, variant):
        lib.morna_debug_set_tuning(8, variant)
        try:
            idx = MornaIndex(n_samples, "unused", dim=dim, sample_threshold=50)
            r = np.random.default_rng(5)
            for j in range(n_rows):
                samples = np.sort(r.choice(n_samples, size=int(lens[j]), replace=False)) + 1
                covs = (1 + r.geometric(0.4, size=int(lens[j]))) * scale
                idx._rows.add("chr%d %d %d" % (j % 22 + 1, 1000 + 7 * j, 1500 + 7 * j), samples, covs)
                passing = int(lens[j] >= 50)
                idx._pass.append(passing); idx._running_freq.append(int(lens[j]) if passing else 0)
            idx.build()
            return idx.accumulator_f64(), idx.matrix_f32(), idx.internal_id_map
        finally:
            lib.morna_debug_set_tuning(8, 3)
    a1, m1, map1 = build(1, 4)          # 4: the warp-per-range variant whatever the rows-per-bucket heuristic says
    a1b, m1b, _ = build(1, 4)
    a0, m0, map0 = build(1, 0)          # 0: barrier-per-row variants only
    a2, m2, _ = build(2, 4)
    assert int((lens >= 50).sum()) > 5000 and float(np.abs(a1).max()) > 0
    assert map1 == map0 and np.array_equal(a1, a0) and np.array_equal(m1, m0)      # both scatter-add variants
    assert np.array_equal(a1, a1b) and np.array_equal(m1, m1b)                     # deterministic
    assert np.array_equal(a2, 2.0 * a1) and np.array_equal(m2, 2.0 * m1)           # exact linearity in the coverages


@pytest.mark.parametrize("n_lists,nq,k_in,k_out", [(8, 300, 100, 100), (2, 5, 7, 7), (3, 64, 20, 10), (4, 33, 16, 40)])
def test_merge_of_sorted_lists_equals_generic_selection(n_lists, nq, k_in, k_out):
    """morna_merge_sorted_topk (rank counting over per-shard sorted lists) == morna_select_topk over the
    concatenation, with ties across lists, short lists (padding) and k_out above and below k_in."""
    from morna_b200 import dist as mdist
    rng = np.random.default_rng(n_lists * 1000 + nq)
    total = n_lists * k_in
    ids = np.full((n_lists, nq, k_in), -1, np.int32)
    d = np.full((n_lists, nq, k_in), np.inf)
    for q in range(nq):
        perm = rng.permutation(10 * total)[:total].astype(np.int32)           # distinct ids across the lists
        vals = np.round(rng.random(total), 2)                                 # two decimals: many equal distances
        for g in range(n_lists):
            m = k_in if (q + g) % 5 else int(rng.integers(0, k_in))           # some lists are short
            li, lv = perm[g * k_in:g * k_in + m], vals[g * k_in:g * k_in + m]
            order = np.lexsort((-li, lv))                                     # distance ascending, id descending
            ids[g, q, :m], d[g, q, :m] = li[order], lv[order]
    ti, td_ = torch.from_numpy(ids).cuda(), torch.from_numpy(d).cuda()
    got_i, got_d = mdist.merge_sorted_lists(ti, td_, k_out)
    flat_i = ti.permute(1, 0, 2).reshape(nq, total)
    flat_d = td_.permute(1, 0, 2).reshape(nq, total)
    want_i, want_d = mdist.merge_topk(flat_i, flat_d, k_out)
    assert torch.equal(got_i, want_i) and torch.equal(got_d, want_d)
