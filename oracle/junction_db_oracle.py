"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's junctions-by-sample database and of the
`junctions` subcommand's filter, statement by statement, for the parity tests of morna_b200/junctions.py.
Nothing on the product path may import this file.

Parity pin: **unpinned**.  The reference ships no expected output for this path (tests/shard_check.py,
tests/junction_integrity.py and tests/interpret_rle.py are integrity checkers, restated as properties in
tests/test_junction_shards.py) and cannot run here (Python 2, mmh3 and annoy are absent).  What is restated:

  encode_64 / decode_64 / increment_64      /root/reference/morna.py:38-75
  running_sum                               /root/reference/morna.py:129-144
  update_junction_dbs                       /root/reference/morna.py:221-341  (called for EVERY row, before the
                                            sample threshold, morna.py:357-359)
  final buffer flush in save()              /root/reference/morna.py:457-488
  `junctions` filter and splice file        /root/reference/morna.py:1486-1632

Two things the reference leaves to its interpreter are modelled explicitly:
  * ``sys.getsizeof(list)`` under 64-bit CPython 2.7 = 72 + 8 * allocated slots, with list_resize's growth rule
    ``new_allocated = (newsize >> 3) + (newsize < 9 ? 3 : 6) + newsize`` (Objects/listobject.c) -- it decides where a
    sample's text is cut into table rows (morna.py:318-319);
  * Python 2 integer division in encode_64 (``num /= 64`` floors).
sqlite is replaced by a dict of row lists per table: the statements used (CREATE, INSERT, SELECT/UPDATE of the last
row) have no other semantics.  The reference's quirks are kept, because its reader (`junctions`) sees them:
  * a sample first seen at junction 0 starts with "!1", which running_sum reads as a run of ABSENT junctions;
  * extending a run that lives in a buffer's first element replaces the whole element ("!5!1" -> "!2"), dropping the
    leading gap (the regex of morna.py:277 keeps only the last run);
  * a sample listed twice in one row gets a gap of -1, encode_64(-1) == "o".
"""
import re
from collections import defaultdict
from math import ceil


def encode_64(num):
    s = [chr(48 + num % 64)]
    num //= 64
    while num > 0:
        s.append(chr(48 + num % 64))
        num //= 64
    return "".join(s[::-1])


def decode_64(s):
    return sum([(ord(s[idx]) - 48) * (64 ** (len(s) - idx - 1)) for idx in range(len(s))])


def increment_64(s):
    return encode_64(decode_64(s) + 1)


def running_sum(rls):
    tot = 0
    for i, item in enumerate(rls):
        length = decode_64(item)
        if i % 2:
            tot += length
        else:
            for i in range(length):
                yield tot + i
            tot += length


class Py2List(object):
    """A list that knows what sys.getsizeof would say about it under CPython 2.7 (64-bit)."""

    def __init__(self):
        self.items = []
        self.allocated = 0

    def append(self, x):
        newsize = len(self.items) + 1
        if newsize > self.allocated:
            self.allocated = (newsize >> 3) + (3 if newsize < 9 else 6) + newsize
        self.items.append(x)

    def getsizeof(self):
        return 72 + 8 * self.allocated

    def __bool__(self):
        return bool(self.items)

    __nonzero__ = __bool__


class JunctionDbOracle(object):
    """update_junction_dbs + the end-of-save flush; ``tables[sample_id]`` = list of [junctions, coverages] rows,
    ``created`` = sample ids in table-creation order, ``shard_of`` filled by the caller's hash."""

    def __init__(self, buffer_size=1024):
        self.buffer_size = buffer_size
        self.junc_id = -1
        self.last_present_junction = defaultdict(lambda: -1)
        self.jns_write_buffer = defaultdict(Py2List)
        self.cov_write_buffer = defaultdict(list)
        self.tables = {}
        self.created = []

    def add_junction(self, samples, coverages):
        self.junc_id += 1                                                  # morna.py:357
        self.update_junction_dbs(samples, coverages)

    def update_junction_dbs(self, samples, coverages):
        for i, sample_id in enumerate(samples):
            if self.last_present_junction[sample_id] == -1:                # :248
                self.tables[sample_id] = []                                # CREATE TABLE sample_%d
                self.created.append(sample_id)
                brand_new_junctions = "!1"
                if self.junc_id > 0:
                    brand_new_junctions = "!" + encode_64(self.junc_id) + "!1"
                self.jns_write_buffer[sample_id].append(brand_new_junctions)
                self.cov_write_buffer[sample_id].append(coverages[i])
            else:
                if self.last_present_junction[sample_id] == self.junc_id - 1:      # :273
                    if self.jns_write_buffer[sample_id]:
                        buf = self.jns_write_buffer[sample_id].items
                        m = re.search("!([0-o]+)$", buf[-1])
                        buf[-1] = "!" + increment_64(m.group(1))
                        self.cov_write_buffer[sample_id].append(coverages[i])
                    else:                                                          # the run's tail is already in the table
                        row = self.tables[sample_id][-1]
                        m = re.search("(.*?)!([0-o]+)$", row[0])
                        row[0] = m.group(1) + "!" + increment_64(m.group(2))
                        row[1] = row[1] + str(coverages[i]) + ","
                else:                                                              # :309
                    self.jns_write_buffer[sample_id].append(
                        "!" + encode_64((self.junc_id - self.last_present_junction[sample_id]) - 1))
                    self.jns_write_buffer[sample_id].append("!1")
                    self.cov_write_buffer[sample_id].append(coverages[i])
                if self.jns_write_buffer[sample_id].getsizeof() > self.buffer_size:    # :318
                    self.tables[sample_id].append(["".join(self.jns_write_buffer[sample_id].items),
                                                   ",".join(str(c) for c in self.cov_write_buffer[sample_id]) + ","])
                    self.jns_write_buffer[sample_id] = Py2List()
                    self.cov_write_buffer[sample_id] = []
            self.last_present_junction[sample_id] = self.junc_id           # :341

    def finish(self):
        """morna.py:457-488: what is left in the buffers becomes one more row per sample."""
        for sample_id in self.last_present_junction.keys():
            if self.jns_write_buffer[sample_id]:
                self.tables[sample_id].append(["".join(self.jns_write_buffer[sample_id].items),
                                               ",".join(str(c) for c in self.cov_write_buffer[sample_id]) + ","])
        return self.tables


def read_sample(rows):
    """morna.py:1513-1533: a table's rows -> (junction indexes, coverage strings)."""
    this_one_juncs = "".join(r[0] for r in rows)
    this_one_covrs = "".join(r[1] for r in rows)
    return [j for j in running_sum(this_one_juncs.split("!"))], this_one_covrs.strip(",").split(",")


def retained_junctions(result_juncs, result_covrs, frequency_filter, coverage_filter):
    """morna.py:1541-1573 -> (sorted retained junction indexes, found_in_map)."""
    retain_junctions = set()
    frequency_counts = defaultdict(int)
    found_in_map = defaultdict(list)
    min_count = int(ceil(frequency_filter * len(result_juncs)))
    for i, junction_list in enumerate(result_juncs):
        for index in junction_list:
            frequency_counts[index] += 1
            found_in_map[index].append(i)
    for i, junction_list in enumerate(result_juncs):
        for index in frequency_counts:
            if frequency_counts[index] >= min_count:
                retain_junctions.add(index)
    for i, coverages_list in enumerate(result_covrs):
        for j, coverage in enumerate(coverages_list):
            if int(coverage) >= coverage_filter:
                retain_junctions.add(result_juncs[i][j])
    return sorted(retain_junctions), found_in_map


def splice_lines(junction_lines, ordered_junctions, found_in_map, result_sample_ids):
    """morna.py:1582-1632: the lines of the splice file."""
    out = []
    ordered_junctions = list(ordered_junctions)
    junction_index = ordered_junctions.pop(0)
    for i, line in enumerate(junction_lines):
        if i == junction_index:
            retain_sample_ids = []
            for result_index in found_in_map[i]:
                retain_sample_ids.append(result_sample_ids[result_index])
            tokens = line.strip().split("\t")
            tokens[1] = str(int(tokens[1]) - 2)
            old_samples = [int(x) for x in tokens[6].split(",")]
            old_covs = [int(x) for x in tokens[7].split(",")]
            new_samples = []
            new_covs = []
            for sample_id in retain_sample_ids:
                if sample_id in old_samples:
                    new_samples.append(sample_id)
                    new_covs.append(old_covs[old_samples.index(sample_id)])
            tokens[6] = ",".join([str(_) for _ in new_samples])
            tokens[7] = ",".join([str(_) for _ in new_covs])
            out.append("\t".join(tokens) + "\t" + str(found_in_map[i]) + "\n")
            try:
                junction_index = ordered_junctions.pop(0)
            except IndexError:
                break
    return out
