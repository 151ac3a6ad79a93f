"""On-disk index contract: basename.stats.mor / .freq.mor / .map.mor (the reference's
files, morna.py:443-455 written, :535-550 read) plus the dense sample-vector store
basename.vec.mor that stands where the reference keeps rows inside Annoy's file.
"""
import os
import pickle
import struct
from collections import defaultdict

import numpy as np

VEC_MAGIC = b"MORNAVEC"
VEC_VERSION = 1
_VEC_HEADER = struct.Struct("<8sIIqq")      # magic, version, dtype code (1 = float32), n, dim


def write_stats(basename, sample_count, index_size, dim):
    """Three text lines: input sample count, retained samples, dimension (morna.py:443-446)."""
    with open(basename + ".stats.mor", "w") as fh:
        fh.write("%d\n%d\n%d\n" % (sample_count, index_size, dim))


def read_stats(basename):
    with open(basename + ".stats.mor") as fh:             # morna.py:535-538
        return int(fh.readline()), int(fh.readline()), int(fh.readline())


def write_freq(basename, sample_frequencies):
    """Pickle protocol 2 of a defaultdict(int) -- loadable by the reference's
    cPickle.load under Python 2 (morna.py:449-451, 546-547)."""
    freq = defaultdict(int)
    freq.update(sample_frequencies)
    with open(basename + ".freq.mor", "wb") as fh:
        pickle.dump(freq, fh, protocol=2)


def read_freq(basename):
    with open(basename + ".freq.mor", "rb") as fh:
        freq = pickle.load(fh, encoding="latin1")       # also reads Python 2 pickles
    if not isinstance(freq, defaultdict):
        d = defaultdict(int)
        d.update(freq)
        freq = d
    return freq


def write_map(basename, internal_id_map):
    """Pickle protocol 2 of {input sample id: internal id} (morna.py:453-455)."""
    with open(basename + ".map.mor", "wb") as fh:
        pickle.dump({int(k): int(v) for k, v in internal_id_map.items()}, fh, protocol=2)


def read_map(basename):
    with open(basename + ".map.mor", "rb") as fh:
        return pickle.load(fh, encoding="latin1")


def write_vectors(basename, matrix_f32):
    """Dense row-major float32 [n x dim]; row i is internal id i (what
    get_item_vector(i) returns in the reference, morna.py:702)."""
    m = np.ascontiguousarray(matrix_f32, dtype="<f4")
    with open(basename + ".vec.mor", "wb") as fh:
        fh.write(_VEC_HEADER.pack(VEC_MAGIC, VEC_VERSION, 1, m.shape[0], m.shape[1]))
        m.tofile(fh)


def read_vectors(basename, mmap=True):
    path = basename + ".vec.mor"
    with open(path, "rb") as fh:
        magic, version, dtype_code, n, dim = _VEC_HEADER.unpack(fh.read(_VEC_HEADER.size))
    if magic != VEC_MAGIC or version != VEC_VERSION or dtype_code != 1:
        raise ValueError("%s is not a morna vector store" % path)
    if os.path.getsize(path) != _VEC_HEADER.size + 4 * n * dim:
        raise ValueError("%s is truncated" % path)
    if mmap:
        return np.memmap(path, dtype="<f4", mode="r", offset=_VEC_HEADER.size, shape=(n, dim))
    return np.fromfile(path, dtype="<f4", offset=_VEC_HEADER.size).reshape(n, dim)


def read_annoy_item_vectors(path, n_items, dim):
    """Item rows of a reference-built basename.annoy.mor.  Annoy's angular node is
    {int32 n_descendants; int32 children[2] (or float norm); float v[dim]} = 12 + 4*dim
    bytes, item nodes first, so row i starts at i*(12+4*dim)+12.  The layout comes from
    Annoy's published source, which is not in the reference tree: unpinned."""
    stride = 12 + 4 * dim
    raw = np.memmap(path, dtype=np.uint8, mode="r")
    if raw.shape[0] < stride * n_items:
        raise ValueError("%s is too small for %d items of dimension %d" % (path, n_items, dim))
    rows = np.lib.stride_tricks.as_strided(raw[12:], shape=(n_items, 4 * dim), strides=(stride, 1))
    return np.ascontiguousarray(rows).view("<f4").reshape(n_items, dim)


def write_meta(basename, metafile):
    """basename.meta.mor, the sqlite table the reference's search joins results against (morna.py:494-520): one row
    per line of `metafile` -- first whitespace-separated column the sample id, the rest of the line (newline included,
    as the reference keeps it) the keywords.  An existing table is dropped first."""
    import sqlite3
    conn = sqlite3.connect(basename + ".meta.mor")
    cursor = conn.cursor()
    cursor.execute("SELECT name FROM sqlite_master WHERE type='table' AND name='metadata'")
    if cursor.fetchone():
        cursor.execute("DROP TABLE metadata")
    cursor.execute("CREATE TABLE metadata (sample_id real, keywords text)")
    with open(metafile) as fh:
        for line in fh:
            parts = line.split(None, 1)
            if len(parts) < 2:
                raise IndexError("list index out of range")           # the reference indexes [1] of a one-column line
            cursor.execute("INSERT INTO metadata VALUES (?,?)", (parts[0], parts[1]))
    conn.commit()
    conn.close()


def read_meta(basename, sample_ids):
    """The join of morna.py:666-676: for each sample id the first matching row's (keywords,) tuple, or None."""
    import sqlite3
    conn = sqlite3.connect(basename + ".meta.mor")
    cursor = conn.cursor()
    out = []
    for sample_id in sample_ids:
        cursor.execute("SELECT keywords FROM metadata WHERE sample_id=?", (str(sample_id),))
        out.append(cursor.fetchone())
    conn.close()
    return out
