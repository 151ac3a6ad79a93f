"""Junctions-by-sample shards and the ``junctions`` subcommand -- host-side mirror of
/root/reference/morna.py:221-341 (``update_junction_dbs``), :457-488 (flush in ``save``) and :1486-1632 (``junctions``).

The reference updates 100 sqlite shards pair by pair while it reads the rows.  Every piece of state it keeps is
per sample (write buffers, last junction seen, the sample's table), so the same tables come out of the TRANSPOSE of
the rows: all (junction index, coverage) pairs of one sample in file order.  Here the rows' pairs are collected as
flat arrays during indexing, one stable sort by sample id on the GPU transposes them, the shard of every sample comes
from the device hash kernel (``mmh3.hash(str(sample_id)) % 100``, morna.py:236), and each sample's text is then
produced by ``sample_rows`` -- the reference's statements applied to one sample's sequence, including where its
buffer is cut into table rows (``sys.getsizeof`` of a CPython 2.7 list, modelled by ``_Py2Buffer``) and its quirks,
which its reader (``junctions``) sees: a sample first seen at junction 0 starts with "!1", which ``running_sum`` reads
as a run of ABSENT junctions; extending a run that lives in a buffer's first element replaces the whole element
("!5!1" -> "!2", morna.py:277-282), dropping the leading gap; a sample listed twice in one row gets a gap of -1,
``encode_64(-1) == "o"``.  Files: ``<basename>.shXX.junc.mor``, one table ``sample_<id>`` per sample
with columns (junctions TEXT, coverages TEXT), as the reference writes them.
"""
import gzip
import os
import sqlite3
from collections import defaultdict
from math import ceil

import numpy as np
import torch

from . import _lib

N_SHARDS = 100


# ---------------------------------------------------------------- base-64 run lengths (morna.py:38-75, 129-144)
def encode_64(num):
    """'0'..'o' as digits 0..63, most significant first; Python 2's floor division kept for negative gaps."""
    s = [chr(48 + num % 64)]
    num //= 64
    while num > 0:
        s.append(chr(48 + num % 64))
        num //= 64
    return "".join(s[::-1])


def decode_64(s):
    v = 0
    for ch in s:
        v = v * 64 + (ord(ch) - 48)
    return v


def increment_64(s):
    return encode_64(decode_64(s) + 1)


def running_sum(rls):
    """morna.py:129-144: even positions are runs of present junctions, odd positions are gaps."""
    tot = 0
    for i, item in enumerate(rls):
        length = decode_64(item)
        if i % 2 == 0:
            for t in range(length):
                yield tot + t
        tot += length


class _Py2Buffer(object):
    """The reference's per-sample token list together with what ``sys.getsizeof`` reports for it under 64-bit
    CPython 2.7 (72 + 8 * allocated slots; list_resize grows to (n >> 3) + (3 if n < 9 else 6) + n)."""
    __slots__ = ("items", "allocated")

    def __init__(self):
        self.items = []
        self.allocated = 0

    def append(self, x):
        n = len(self.items) + 1
        if n > self.allocated:
            self.allocated = (n >> 3) + (3 if n < 9 else 6) + n
        self.items.append(x)

    def sizeof(self):
        return 72 + 8 * self.allocated


def sample_rows(junc_ids, coverages, buffer_size=1024):
    """Rows [junctions, coverages] of one sample's table from its occurrences in file order
    (morna.py:246-341 per occurrence, :472-481 for what is left in the buffer)."""
    rows = []
    jbuf, cbuf = _Py2Buffer(), []
    last = -1
    for j, c in zip(junc_ids, coverages):
        if last == -1:                                          # first time seen: the table starts with a buffer
            jbuf.append("!1" if j <= 0 else "!" + encode_64(j) + "!1")
            cbuf.append(c)
        else:
            if last == j - 1:                                   # on a run of present junctions
                if jbuf.items:                                  # the whole last element becomes the longer run (:277-282)
                    tail = jbuf.items[-1]
                    jbuf.items[-1] = "!" + increment_64(tail[tail.rindex("!") + 1:])
                    cbuf.append(c)
                else:                                           # the run's tail already sits in the table (:288-305)
                    row = rows[-1]
                    cut = row[0].rindex("!")
                    row[0] = row[0][:cut] + "!" + increment_64(row[0][cut + 1:])
                    row[1] = row[1] + str(c) + ","
            else:
                jbuf.append("!" + encode_64(j - last - 1))
                jbuf.append("!1")
                cbuf.append(c)
            if jbuf.sizeof() > buffer_size:                     # :318-337
                rows.append(["".join(jbuf.items), ",".join(str(v) for v in cbuf) + ","])
                jbuf, cbuf = _Py2Buffer(), []
        last = j
    if jbuf.items:
        rows.append(["".join(jbuf.items), ",".join(str(v) for v in cbuf) + ","])
    return rows


def hash_strings(strings, device=None):
    """mmh3.hash of every string (signed int32) through the device hash kernel (K1)."""
    _lib.require_cuda()
    lib = _lib.load()
    blobs = [s.encode("utf-8") for s in strings]
    n = len(blobs)
    if n == 0:
        return np.zeros(0, np.int32)
    dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
    off = np.zeros(n + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(b) for b in blobs])
    packed = np.frombuffer(b"".join(blobs) + b"\0", dtype=np.uint8).copy()
    with torch.cuda.device(dev):
        d_keys, d_off = torch.from_numpy(packed).to(dev), torch.from_numpy(off).to(dev)
        raw = torch.empty(n, dtype=torch.int32, device=dev)
        bucket = torch.empty(n, dtype=torch.int32, device=dev)
        sign = torch.empty(n, dtype=torch.int8, device=dev)
        _lib.check(lib.morna_hash_junctions(_lib.dev_ptr(d_keys), _lib.dev_ptr(d_off), n, N_SHARDS, _lib.dev_ptr(raw),
                                            _lib.dev_ptr(bucket), _lib.dev_ptr(sign), _lib.stream_ptr()), "morna_hash_junctions")
        return raw.cpu().numpy()


def shard_ids(sample_ids, device=None):
    """``mmh3.hash(str(sample_id)) % 100`` (Python modulo: never negative) per sample id."""
    raw = hash_strings([str(int(s)) for s in sample_ids], device)
    return (raw.astype(np.int64) % N_SHARDS).astype(np.int32)


def shard_path(basename, shard_id):
    return basename + ".sh" + format(int(shard_id), "02d") + ".junc.mor"


def remove_shards(basename):
    """morna.py:200-205: a new index removes the shards of an old one."""
    for shard_id in range(N_SHARDS):
        try:
            os.remove(shard_path(basename, shard_id))
        except OSError:
            pass


class ShardRecorder(object):
    """Collects the pairs of EVERY row (the reference records a row before the sample threshold looks at it,
    morna.py:357-363) and writes the shards."""

    def __init__(self, basename, buffer_size=1024, device=None):
        self.basename = basename
        self.buffer_size = buffer_size
        self.device = device
        self._junc, self._sample, self._cov = [], [], []

    def add_row(self, junc_id, samples, coverages):
        n = len(samples)
        if len(coverages) < n:                  # the reference reads coverages[i] for every sample (morna.py:269, 284, 313)
            raise IndexError("list index out of range")
        if n:
            self._junc.append(np.full(n, junc_id, dtype=np.int64))
            self._sample.append(np.asarray(samples, dtype=np.int64))
            self._cov.append(np.asarray(coverages[:n], dtype=np.int64))

    def add_rows(self, first_junc_id, row_lens, samples, coverages):
        """A block of consecutive rows: ``row_lens[i]`` pairs of row ``first_junc_id + i``, flat pair arrays."""
        row_lens = np.asarray(row_lens, dtype=np.int64)
        if row_lens.sum():
            self._junc.append(np.repeat(np.arange(first_junc_id, first_junc_id + len(row_lens), dtype=np.int64), row_lens))
            self._sample.append(np.asarray(samples, dtype=np.int64).copy())
            self._cov.append(np.asarray(coverages, dtype=np.int64).copy())

    def transposed(self):
        """-> (sample ids in first-seen order, {sample id: (junction indexes, coverages)}) ; pairs of one sample keep
        file order (stable sort by sample id, on the GPU)."""
        if not self._junc:
            return [], {}
        junc, sample, cov = np.concatenate(self._junc), np.concatenate(self._sample), np.concatenate(self._cov)
        dev = torch.device(self.device if self.device is not None else "cuda:%d" % torch.cuda.current_device())
        keys = torch.from_numpy(sample).to(dev)
        order = torch.sort(keys, stable=True).indices
        sorted_keys = keys[order]
        uniq, counts = torch.unique_consecutive(sorted_keys, return_counts=True)
        starts = torch.cumsum(counts, 0) - counts
        first_pos = order[starts]                                   # pair position of every sample's first occurrence
        seen_order = torch.argsort(first_pos)
        order, uniq, counts, starts, seen_order = (t.cpu().numpy() for t in (order, uniq, counts, starts, seen_order))
        junc, cov = junc[order], cov[order]
        per_sample = {}
        for u, s0, c in zip(uniq.tolist(), starts.tolist(), counts.tolist()):
            per_sample[u] = (junc[s0:s0 + c].tolist(), cov[s0:s0 + c].tolist())
        return uniq[seen_order].tolist(), per_sample

    def tables(self):
        """-> (sample ids in table-creation order, {sample id: rows})."""
        created, per_sample = self.transposed()
        return created, {s: sample_rows(per_sample[s][0], per_sample[s][1], self.buffer_size) for s in created}

    def write(self):
        """The shard files of this index (tables in first-seen order inside each shard)."""
        remove_shards(self.basename)
        created, tables = self.tables()
        shards = shard_ids(created, self.device)
        conns = {}
        for sample_id, shard_id in zip(created, shards.tolist()):
            conn = conns.get(shard_id)
            if conn is None:
                conn = conns[shard_id] = sqlite3.connect(shard_path(self.basename, shard_id))
                conn.isolation_level = None
                conn.execute("BEGIN")
            conn.execute("CREATE TABLE sample_%d (junctions TEXT,coverages TEXT)" % sample_id)
            conn.executemany("INSERT INTO sample_%d VALUES (?, ?)" % sample_id, tables[sample_id])
        for conn in conns.values():                                  # morna.py:484-488
            conn.commit()
            conn.execute("VACUUM")
            conn.close()
        return len(created)


def read_sample_table(basename, sample_id, shard_id):
    """morna.py:1504-1533: (junction indexes, coverage strings) of one sample from its shard."""
    conn = sqlite3.connect(shard_path(basename, shard_id))
    try:
        this_one_juncs, this_one_covrs = zip(*list(conn.execute("SELECT * FROM sample_%d" % sample_id)))
    finally:
        conn.close()
    this_one_juncs = "".join(this_one_juncs)
    this_one_covrs = "".join(this_one_covrs)
    return [j for j in running_sum(this_one_juncs.split("!"))], this_one_covrs.strip(",").split(",")


def retained_junctions(result_juncs, result_covrs, frequency_filter, coverage_filter):
    """morna.py:1541-1573: junction indexes found in at least ``frequency_filter`` of the result samples, or covered
    at least ``coverage_filter`` times in one of them -> (sorted indexes, {index: result numbers it was found in})."""
    found_in_map = defaultdict(list)
    for i, junction_list in enumerate(result_juncs):
        for index in junction_list:
            found_in_map[index].append(i)
    min_count = int(ceil(frequency_filter * len(result_juncs)))
    retain = set()
    if result_juncs:                                                 # the reference's loop body runs once per result
        retain.update(index for index, where in found_in_map.items() if len(where) >= min_count)
    for i, coverages_list in enumerate(result_covrs):
        for j, coverage in enumerate(coverages_list):
            if int(coverage) >= coverage_filter:
                retain.add(result_juncs[i][j])                       # IndexError if the lists disagree, as in the reference
    return sorted(retain), found_in_map


def write_splicefile(splicefile, junction_file, ordered_junctions, found_in_map, result_sample_ids, stderr):
    """morna.py:1582-1632: the retained rows of the junction file, start - 2, restricted to the result samples, with
    the result numbers appended."""
    with open(splicefile, "w") as splices, gzip.open(junction_file, "rt") as junction_names:
        ordered_junctions = list(ordered_junctions)
        stderr.write(str(len(ordered_junctions)) + " junctions to begin with\n")
        junction_index = ordered_junctions.pop(0)                    # IndexError when nothing was retained (reference)
        stderr.flush()
        for i, line in enumerate(junction_names):
            if i != junction_index:
                continue
            retain_sample_ids = [result_sample_ids[r] for r in found_in_map[i]]
            tokens = line.strip().split("\t")
            tokens[1] = str(int(tokens[1]) - 2)
            old_samples = [int(x) for x in tokens[6].split(",")]
            old_covs = [int(x) for x in tokens[7].split(",")]
            position = {}
            for at, s in enumerate(old_samples):
                position.setdefault(s, at)                           # list.index: first occurrence
            new_samples = [s for s in retain_sample_ids if s in position]
            new_covs = [old_covs[position[s]] for s in new_samples]
            tokens[6] = ",".join(str(_) for _ in new_samples)
            tokens[7] = ",".join(str(_) for _ in new_covs)
            splices.write("\t".join(tokens) + "\t" + str(found_in_map[i]) + "\n")
            if not ordered_junctions:
                break
            junction_index = ordered_junctions.pop(0)


def go_junctions(args, searcher, results, stdout, stderr):
    """The ``junctions`` part of the reference's main (morna.py:1486-1632) after the search produced ``results``."""
    filter_whole = args.junction_filter.split(",")
    frequency_filter = float(filter_whole[0])
    coverage_filter = int(filter_whole[1])
    result_sample_ids = [searcher.inverse_lookup(result) for result in results[0]]
    shards = shard_ids(result_sample_ids, searcher.device).tolist() if result_sample_ids else []
    result_juncs, result_covrs = [], []
    for sample_id, shard_id in zip(result_sample_ids, shards):
        stdout.write("shard_id is " + format(shard_id, "02d") + "\n")
        juncs, covrs = read_sample_table(args.basename, sample_id, shard_id)
        result_juncs.append(juncs)
        result_covrs.append(covrs)
    stdout.write("result_juncs lengths: \n")
    stdout.write(str([len(ls) for ls in result_juncs]) + "\n")
    stdout.write("result_covrs lengths: \n")
    stdout.write(str([len(ls) for ls in result_covrs]) + "\n")
    ordered, found_in_map = retained_junctions(result_juncs, result_covrs, frequency_filter, coverage_filter)
    stdout.write("Number of retained junctions: " + str(len(ordered)) + "\n")
    write_splicefile(args.splicefile, args.junction_file, ordered, found_in_map, result_sample_ids, stderr)
