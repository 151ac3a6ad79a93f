"""TEST INFRASTRUCTURE ONLY -- literal CPU restatement of morna's hot path.

Every function cites the reference lines (``/root/reference/morna.py``) it
follows.  Arithmetic is kept exactly as the reference does it (Python floats ==
IEEE doubles, sequential left-to-right sums, float32 rounding where Annoy's
``add_item`` rounds), so results are the ground truth the CUDA path is compared
with.  Slow by construction; ``oracle/oracle.c`` is the same arithmetic in C for
larger cases and for the CPU baseline, and the ``*_np`` helpers are vectorised
twins for mid-size checks.

Third-party arithmetic absent from the reference tree (unpinned there):
  * ``mmh3.hash`` -- MurmurHash3_x86_32, seed 0, signed 32-bit result, over the
    key's bytes (call sites morna.py:369, 591, 625).  Restated below from the
    published algorithm; cross-checked in tests against sklearn's
    ``murmurhash3_32`` and published known answers.
  * ``annoy`` -- on the exact path only ``add_item`` (cast to float32) and
    ``get_item_vector`` (returns those float32 values) matter (morna.py:406, 702).
"""
import bisect
import math
from collections import defaultdict

import numpy as np

_M32 = 0xFFFFFFFF


def murmur3_x86_32(data, seed=0):
    """MurmurHash3_x86_32 of ``data`` (bytes or str) -> signed int32.

    Stands in for ``mmh3.hash(junction)`` at morna.py:369.  Published algorithm:
    4-byte little-endian blocks, c1=0xcc9e2d51, c2=0x1b873593, rotl 15 / 13,
    h*5+0xe6546b64, 1..3 byte tail, h^=len, fmix32 (0x85ebca6b, 0xc2b2ae35).
    """
    if isinstance(data, str):
        data = data.encode("utf-8")
    n = len(data)
    h = seed & _M32
    c1, c2 = 0xCC9E2D51, 0x1B873593
    nblocks = n // 4
    for b in range(nblocks):
        k = int.from_bytes(data[4 * b:4 * b + 4], "little")
        k = (k * c1) & _M32
        k = ((k << 15) | (k >> 17)) & _M32
        k = (k * c2) & _M32
        h ^= k
        h = ((h << 13) | (h >> 19)) & _M32
        h = (h * 5 + 0xE6546B64) & _M32
    tail = data[4 * nblocks:]
    k = 0
    if len(tail) == 3:
        k ^= tail[2] << 16
    if len(tail) >= 2:
        k ^= tail[1] << 8
    if len(tail) >= 1:
        k ^= tail[0]
        k = (k * c1) & _M32
        k = ((k << 15) | (k >> 17)) & _M32
        k = (k * c2) & _M32
        h ^= k
    h ^= n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & _M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & _M32
    h ^= h >> 16
    return h - (1 << 32) if h & 0x80000000 else h


def bucket_and_sign(junction, dim):
    """(raw hash, bucket, sign) as morna.py:369-371: sign from the signed hash,
    bucket by Python floor-mod (non-negative for negative hashes)."""
    h = murmur3_x86_32(junction)
    return h, h % dim, (-1 if h < 0 else 1)


def tokenize_line(line):
    """One intropolis row -> (junction key, samples, coverages); morna.py:848-853.

    key = first three tab fields joined by single spaces (strand is not part of
    the key); samples / coverages = second-to-last / last field split on ','.
    """
    tokens = line.strip().split("\t")
    return (" ".join(tokens[:3]),
            [int(t) for t in tokens[-2].split(",")],
            [int(t) for t in tokens[-1].split(",")])


def count_samples(lines):
    """Number of distinct sample-id *strings* in field -2; morna.py:809-822."""
    seen = set()
    for line in lines:
        seen.update(line.split("\t")[-2].split(","))
    return len(seen)


class OracleIndex(object):
    """State and arithmetic of ``MornaIndex`` on the hot path (morna.py:164-219,
    344-388, 390-425) without Annoy's forest and without the sqlite side files."""

    def __init__(self, sample_count, dim=3000, sample_threshold=100):
        self.sample_count = sample_count
        self.dim = dim
        self.sample_threshold = sample_threshold
        self.internal_id_map = {}
        self.new_internal_id = 0
        self.sample_frequencies = defaultdict(int)
        self.rows = {}            # internal id -> list of dim Python floats
        self.skipped = 0
        self.junc_id = -1
        # per passing row: (raw hash, bucket, sign, idf) -- for kernel parity tests
        self.row_trace = []

    def add_junction(self, junction, samples, coverages):
        """morna.py:344-388 (minus update_junction_dbs at :359)."""
        self.junc_id += 1
        if len(samples) < self.sample_threshold:        # :361-363
            self.skipped += 1
            return
        self.sample_frequencies[junction] += len(samples)   # :365 cumulative
        h = murmur3_x86_32(junction)                     # :369
        sign = -1 if h < 0 else 1                        # :370
        bucket = h % self.dim                            # :371 floor-mod
        idf = math.log(float(self.sample_count)
                       / self.sample_frequencies[junction])  # :372-374
        self.row_trace.append((h, bucket, sign, idf))
        for sample_id, coverage in zip(samples, coverages):  # :376
            if sample_id not in self.internal_id_map:    # :377-382 first-seen order
                self.internal_id_map[sample_id] = self.new_internal_id
                self.new_internal_id += 1
            internal = self.internal_id_map[sample_id]
            row = self.rows.get(internal)
            if row is None:
                row = self.rows[internal] = [0.0] * self.dim   # :184-186
            row[bucket] += sign * (coverage * idf)       # :384-388

    def matrix_f64(self):
        """Dense [new_internal_id x dim] double matrix, row = internal id."""
        if self.new_internal_id == 0:                    # :399-403
            raise ValueError("No internal ids were assigned, indicating that no "
                             "samples were added to the index. Likely caused when "
                             "no junctions pass the sample threshold.")
        out = np.zeros((self.new_internal_id, self.dim), dtype=np.float64)
        for internal, row in self.rows.items():
            out[internal, :] = row
        return out

    def matrix_f32(self):
        """What ``add_item`` stores (morna.py:405-407, 422-424): float32 rows."""
        return self.matrix_f64().astype(np.float32)


def go_index(lines, features=3000, sample_count=None, sample_threshold=100):
    """Line loop of ``go_index`` (morna.py:824-861) over an iterable of text rows."""
    lines = list(lines)
    if not sample_count:
        sample_count = count_samples(lines)              # :830-832
    idx = OracleIndex(sample_count, dim=features, sample_threshold=sample_threshold)
    for line in lines:
        idx.add_junction(*tokenize_line(line))
    return idx


def finalize_query(query, sample_frequencies, sample_count, dim):
    """morna.py:609-629.  ``query``: {(chrom, start, end): summed coverage} in
    insertion order; returns the dim-long list of Python floats."""
    out = [0.0] * dim
    for junction, cov in query.items():
        key = " ".join(str(t) for t in junction)
        freq = sample_frequencies.get(key, 0)
        if freq == 0:
            idf = 0
        else:
            idf = math.log(float(sample_count) / freq)
        h = murmur3_x86_32(key)
        sign = -1 if h < 0 else 1
        out[h % dim] += sign * (cov * idf)
    return out


def cosine_distance(v1, v2, clamp=False):
    """morna.py:101-114 literally: one pass, three double sums in index order,
    sqrt(2 - 2*pq/sqrt(pp*qq)); sqrt(2) when pp*qq <= 0.  The reference does not
    clamp, so a rounding-negative radicand raises ValueError there (math domain
    error); ``clamp=True`` floors it at 0 (what Annoy itself and the CUDA path do).
    """
    pp = qq = pq = 0.0
    for a, b in zip(v1, v2):
        pp += a * a
        qq += b * b
        pq += a * b
    ppqq = pp * qq
    if ppqq > 0.0:
        d = 2.0 - 2.0 * pq / math.sqrt(ppqq)
    else:
        d = 2.0
    if clamp and d < 0.0:
        d = 0.0
    return math.sqrt(d)


def exact_search_nn(matrix_f32, query, num_neighbors, clamp=False):
    """morna.py:697-712: scan rows in id order, ``bisect_left`` insert, truncate.
    Order: distance ascending; among equal distances the later-scanned (higher id)
    row sits first and evicts earlier ones.  Returns (ids, distances)."""
    ids, dists = [], []
    q = [float(x) for x in query]
    for i in range(matrix_f32.shape[0]):
        d = cosine_distance(matrix_f32[i].tolist(), q, clamp=clamp)   # :701-703
        at = bisect.bisect_left(dists, d)                # :705
        if at < num_neighbors:                           # :707-709
            dists.insert(at, d)
            ids.insert(at, i)
        if len(dists) > num_neighbors:                   # :710-712
            dists = dists[:num_neighbors]
            ids = ids[:num_neighbors]
    return ids, dists


# ----------------------------------------------------------------------------
# vectorised twins (numpy) -- same definitions, pairwise/BLAS summation order, so
# equal to the literal functions to ~1e-15 relative in the sums, not bit-for-bit.
# ----------------------------------------------------------------------------

def distances_np(matrix_f32, query, clamp=True):
    s = np.asarray(matrix_f32, dtype=np.float64)
    q = np.asarray(query, dtype=np.float64)
    pp = np.einsum("ij,ij->i", s, s)
    qq = float(q @ q)
    pq = s @ q
    ppqq = pp * qq
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.where(ppqq > 0.0, 2.0 - 2.0 * pq / np.sqrt(ppqq), 2.0)
    if clamp:
        d = np.maximum(d, 0.0)
    return np.sqrt(d)


def topk_rule(dists, num_neighbors, ids=None):
    """Top-k of a distance array under the reference order (distance asc, id desc)."""
    dists = np.asarray(dists)
    if ids is None:
        ids = np.arange(dists.shape[0])
    order = np.lexsort((-np.asarray(ids, dtype=np.int64), dists))[:num_neighbors]
    return np.asarray(ids)[order], dists[order]


def exact_search_np(matrix_f32, query, num_neighbors):
    return topk_rule(distances_np(matrix_f32, query), num_neighbors)


def format_results(results):
    """``results_output`` (morna.py:116-127) as a string.  Python 2 ``str(float)``
    prints 12 significant digits; ``repr``-style shortest round-trip is Python 3's
    ``str``.  The drop-in CLI keeps Python 2's ``%.12g`` look."""
    out = []
    for i in range(len(results[0])):
        cells = [str(i + 1) + "."]
        for column in results:
            v = column[i]
            cells.append(py2_str(v))
        out.append("\t".join(cells) + "\n")
    return "".join(out)


def py2_str(v):
    """Python 2 ``str()`` of ints / floats (floats: ``%.12g`` plus a trailing
    ``.0`` when the result looks integral)."""
    if isinstance(v, (float, np.floating)):
        s = "%.12g" % float(v)
        if s.lstrip("-").isdigit():
            s += ".0"
        return s
    if isinstance(v, (np.integer,)):
        return str(int(v))
    return str(v)
