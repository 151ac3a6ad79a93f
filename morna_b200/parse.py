"""Host-side text parsing either side of the hot path: intropolis rows (index input)
and query junction streams (search input).  Mirrors go_index's tokenising
(morna.py:841-861), count_samples (:789-822) and utils.py:194-290.
"""
import gzip
import re
import sys

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
    """Pre-tokenised rows in the binary CSR form the kernels stream: packed key bytes + int32 offsets,
    int64 pair offsets, int32 samples / coverages.  Rows arrive one at a time (``add``) or as whole
    tokenised blocks (``add_block``); ``finish`` concatenates in arrival order."""

    def __init__(self):
        self.key_chunks = []      # bytes / uint8 arrays
        self.keylen_chunks = []   # int64 arrays: key length per row
        self.len_chunks = []      # int64 arrays: pairs per row
        self.sample_chunks = []   # int32 arrays
        self.cov_chunks = []
        self.n_rows = 0

    def add(self, key, samples, coverages):
        n = min(len(samples), len(coverages))        # zip() semantics, morna.py:376
        kb = key.encode("utf-8") if isinstance(key, str) else bytes(key)
        self.key_chunks.append(np.frombuffer(kb, dtype=np.uint8))
        self.keylen_chunks.append(np.array([len(kb)], dtype=np.int64))
        self.len_chunks.append(np.array([n], dtype=np.int64))
        self.sample_chunks.append(np.asarray(samples[:n], dtype=np.int32))
        self.cov_chunks.append(np.asarray(coverages[:n], dtype=np.int32))
        self.n_rows += 1

    def add_block(self, packed_keys, key_lens, pair_lens, samples, coverages):
        """Rows of one tokenised block, already concatenated (arrays are kept, not copied)."""
        self.key_chunks.append(packed_keys)
        self.keylen_chunks.append(np.asarray(key_lens, dtype=np.int64))
        self.len_chunks.append(np.asarray(pair_lens, dtype=np.int64))
        self.sample_chunks.append(samples)
        self.cov_chunks.append(coverages)
        self.n_rows += len(key_lens)

    def __len__(self):
        return self.n_rows

    def finish(self):
        n_rows = self.n_rows
        cat = lambda parts, dt: (np.concatenate(parts) if parts else np.zeros(0, dt)).astype(dt, copy=False)
        key_off = np.zeros(n_rows + 1, dtype=np.int64)
        if n_rows:
            key_off[1:] = np.cumsum(cat(self.keylen_chunks, np.int64))
        if key_off[-1] > 0x7fffffff:
            raise ValueError("junction keys exceed 2 GiB")
        packed = cat(self.key_chunks, np.uint8)
        row_off = np.zeros(n_rows + 1, dtype=np.int64)
        if n_rows:
            row_off[1:] = np.cumsum(cat(self.len_chunks, np.int64))
        return (packed, key_off.astype(np.int32), row_off, cat(self.sample_chunks, np.int32),
                cat(self.cov_chunks, np.int32))


def tokenize_buffer(buf, n_threads=None):
    """Native tokenizer (morna_b200/csrc/tokenize.cpp) over a bytes object holding whole lines.
    Returns (packed_keys u8, key_off i32[n+1], row_off i64[n+1], sample i32, cov i32, line_off i64[n+1],
    needs_python u8[n]); rows with needs_python set are empty and must go through ``tokenize_line``."""
    import ctypes
    import os
    from . import _lib
    lib = _lib.load()
    if n_threads is None:
        n_threads = max(1, min(32, os.cpu_count() or 1))
    n = ctypes.c_int64()
    kb = ctypes.c_int64()
    pr = ctypes.c_int64()
    nbytes = len(buf)
    _lib.check(lib.morna_tokenize_count(buf, nbytes, n_threads, ctypes.byref(n), ctypes.byref(kb), ctypes.byref(pr)),
               "morna_tokenize_count")
    n, kb, pr = n.value, kb.value, pr.value
    keys = np.empty(max(kb, 1), dtype=np.uint8)
    key_off = np.empty(n + 1, dtype=np.int32)
    row_off = np.empty(n + 1, dtype=np.int64)
    sample = np.empty(max(pr, 1), dtype=np.int32)
    cov = np.empty(max(pr, 1), dtype=np.int32)
    line_off = np.empty(n + 1, dtype=np.int64)
    needs = np.empty(max(n, 1), dtype=np.uint8)
    _lib.check(lib.morna_tokenize_fill(buf, nbytes, n_threads, keys.ctypes.data, key_off.ctypes.data, row_off.ctypes.data,
                                       sample.ctypes.data, cov.ctypes.data, line_off.ctypes.data, needs.ctypes.data),
               "morna_tokenize_fill")
    return keys[:kb], key_off, row_off, sample[:pr], cov[:pr], line_off, needs[:n]


def _read_blocks_sync(fh, block_bytes):
    tail = b""
    while True:
        chunk = fh.read(block_bytes)
        if not chunk:
            if tail:
                yield tail
            return
        cut = chunk.rfind(b"\n")
        if cut < 0:
            tail += chunk
            continue
        yield tail + chunk[:cut + 1]
        tail = chunk[cut + 1:]


def read_blocks(fh, block_bytes=32 << 20, prefetch=2):
    """Binary file object -> bytes blocks ending at a line boundary.  A reader thread stays ``prefetch``
    blocks ahead (zlib releases the GIL), so gunzip overlaps the tokenising and bookkeeping of the
    previous block -- gunzip is the slowest stage of an index build end to end."""
    if prefetch <= 0:
        for block in _read_blocks_sync(fh, block_bytes):
            yield block
        return
    import queue
    import threading
    q = queue.Queue(maxsize=prefetch)
    done = object()

    def reader():
        try:
            for block in _read_blocks_sync(fh, block_bytes):
                q.put(block)
            q.put(done)
        except BaseException as exc:            # surface read errors in the consumer
            q.put(exc)

    t = threading.Thread(target=reader, daemon=True)
    t.start()
    while True:
        item = q.get()
        if item is done:
            break
        if isinstance(item, BaseException):
            raise item
        yield item
    t.join()


def open_intropolis_binary(path):
    with open(path, "rb") as probe:
        magic = probe.read(2)
    return gzip.open(path, "rb") if magic == b"\x1f\x8b" else open(path, "rb")


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
        try:                                   # utils.py:268-290: a short line prints to stderr, then the IndexError propagates
            tokens = line.strip().split("\t")
            flag = int(tokens[1])
            if flag & 4:                       # unmapped reads are skipped before any other column is touched
                continue
            rname, cigar, pos = tokens[2], tokens[5], int(tokens[3])
            tokens[9]                          # the reference reads SEQ here
        except IndexError:
            sys.stderr.write("Error found on line: " + line + "\n")
            raise
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
