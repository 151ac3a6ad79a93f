"""Host-side text parsing either side of the hot path: intropolis rows (index input)
and query junction streams (search input).  Mirrors go_index's tokenising
(morna.py:841-861), count_samples (:789-822) and utils.py:194-290.
"""
import gzip
import re

import numpy as np


def open_intropolis(path):
    """The reference insists on gzip (morna.py:831, 841); plain text is accepted too
    (its own fixture tests/tiny_intropolis.tsv is not gzipped)."""
    with open(path, "rb") as probe:
        magic = probe.read(2)
    if magic == b"\x1f\x8b":
        return gzip.open(path, "rt")
    return open(path, "rt")


def tokenize_line(line):
    """(junction key, samples, coverages) of one row; morna.py:848-853."""
    tokens = line.strip().split("\t")
    return (" ".join(tokens[:3]),
            [int(t) for t in tokens[-2].split(",")],
            [int(t) for t in tokens[-1].split(",")])


def count_samples(lines, verbose=False, out=None):
    """Distinct sample-id strings in field -2 (morna.py:809-822)."""
    seen = set()
    for i, line in enumerate(lines):
        if verbose and out is not None and i % 100 == 0:
            out.write("%d lines into sample count, %d samples so far.\r" % (i, len(seen)))
            out.flush()
        seen.update(line.split("\t")[-2].split(","))
    return len(seen)


class RowBatch(object):
    """Pre-tokenised rows in the binary CSR form the kernels stream:
    packed key bytes + int32 offsets, int64 pair offsets, int32 samples / coverages."""

    def __init__(self):
        self.keys = []            # list of bytes
        self.lens = []            # pairs per row
        self.samples = []         # list of int32 arrays
        self.coverages = []

    def add(self, key, samples, coverages):
        n = min(len(samples), len(coverages))        # zip() semantics, morna.py:376
        self.keys.append(key.encode("utf-8") if isinstance(key, str) else bytes(key))
        self.lens.append(n)
        self.samples.append(np.asarray(samples[:n], dtype=np.int32))
        self.coverages.append(np.asarray(coverages[:n], dtype=np.int32))

    def __len__(self):
        return len(self.keys)

    def finish(self):
        n_rows = len(self.keys)
        key_off = np.zeros(n_rows + 1, dtype=np.int32)
        if n_rows:
            key_off[1:] = np.cumsum([len(k) for k in self.keys])
        packed = np.frombuffer(b"".join(self.keys), dtype=np.uint8).copy() if n_rows else np.zeros(0, np.uint8)
        row_off = np.zeros(n_rows + 1, dtype=np.int64)
        if n_rows:
            row_off[1:] = np.cumsum(self.lens, dtype=np.int64)
        cat = lambda parts: (np.concatenate(parts) if parts else np.zeros(0, np.int32)).astype(np.int32, copy=False)
        return packed, key_off, row_off, cat(self.samples), cat(self.coverages)


# ---------------------------------------------------------------- query streams
def junctions_from_raw_stream(stream):
    """utils.py:194-204: chrom, start, end (1-based inclusive), coverage."""
    for line in stream:
        tokens = line.strip().split("\t")
        yield (tokens[0], int(tokens[1]), int(tokens[2]), int(tokens[3]))


def junctions_from_bed_stream(stream):
    """utils.py:206-252: BED12 blocks -> junctions between consecutive blocks."""
    for line in stream:
        tokens = line.rstrip().split("\t")
        if len(tokens) < 12:
            continue
        chrom, chrom_start, coverage = tokens[0], int(tokens[1]), int(tokens[4])
        sizes = [t for t in tokens[10].split(",")]
        starts = [t for t in tokens[11].split(",")]
        if sizes and not _is_int(sizes[-1]):
            sizes = sizes[:-1]
        if starts and not _is_int(starts[-1]):
            starts = starts[:-1]
        if len(sizes) < 2:
            continue
        assert len(sizes) == len(starts)
        sizes = [int(t) for t in sizes]
        starts = [int(t) for t in starts]
        for i in range(len(sizes) - 1):
            left = chrom_start + starts[i] + sizes[i]       # end of block i (0-based exclusive)
            right = chrom_start + starts[i + 1]             # start of block i+1
            yield (chrom, left + 1, right, coverage)


def _is_int(text):
    try:
        int(text)
        return True
    except ValueError:
        return False


_CIGAR_OP = re.compile(r"(\d+)([MINDS])")


def junctions_from_sam_stream(stream):
    """utils.py:254-290: one junction per N operation of a primary, mapped
    alignment; start = first skipped base, end = last skipped base, coverage 1."""
    for line in stream:
        if line[0] == "@":
            continue
        tokens = line.strip().split("\t")
        if len(tokens) < 10:
            raise IndexError("Error found on line: " + line)
        flag = int(tokens[1])
        if flag & 4:
            continue
        rname, pos, cigar = tokens[2], int(tokens[3]), tokens[5]
        if "N" not in cigar or flag & 256:
            continue
        if _CIGAR_OP.sub("", cigar):
            raise RuntimeError("Accepted CIGAR characters are only in [MINDS].")
        for size, op in _CIGAR_OP.findall(cigar):
            size = int(size)
            if op == "N":
                yield (rname, pos, pos + size - 1, 1)
                pos += size
            elif op in "MD":
                pos += size
