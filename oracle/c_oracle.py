"""TEST INFRASTRUCTURE ONLY -- ctypes front for oracle/oracle.c (the C oracle).

``build()`` is the committed recipe: one gcc call, no fast-math, no FMA
contraction, so the C sums are bit-identical to the Python oracle's.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "oracle.c")
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
           "-fno-fast-math", "-pthread", "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_murmur3_32.restype = ctypes.c_int32
        _lib.oracle_exact_search.restype = ctypes.c_int32
        _lib.oracle_index_accumulate.restype = ctypes.c_int32
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def murmur3_32(key, seed=0):
    if isinstance(key, str):
        key = key.encode("utf-8")
    return int(lib().oracle_murmur3_32(key, ctypes.c_int32(len(key)), ctypes.c_uint32(seed)))


def hash_rows(keys, dim):
    """keys: list of str/bytes -> (raw int32[J], bucket int32[J], sign int8[J])."""
    blobs = [k.encode("utf-8") if isinstance(k, str) else k for k in keys]
    off = np.zeros(len(blobs) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(b) for b in blobs])
    packed = np.frombuffer(b"".join(blobs) + b"\0", dtype=np.uint8).copy()
    raw = np.empty(len(blobs), np.int32)
    bucket = np.empty(len(blobs), np.int32)
    sign = np.empty(len(blobs), np.int8)
    lib().oracle_hash_rows(_p(packed), _p(off), ctypes.c_int64(len(blobs)),
                           ctypes.c_int32(dim), _p(raw), _p(bucket), _p(sign))
    return raw, bucket, sign


def distances(S, q, clamp=True):
    S = np.ascontiguousarray(S, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.empty(S.shape[0], np.float64)
    lib().oracle_distances(_p(S), ctypes.c_int64(S.shape[0]), ctypes.c_int32(S.shape[1]),
                           ctypes.c_int64(S.shape[1]), _p(q), ctypes.c_int(int(clamp)), _p(out))
    return out


def exact_search(S, q, k, clamp=True):
    S = np.ascontiguousarray(S, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float64)
    ids = np.full(k, -1, np.int32)
    dists = np.full(k, np.inf, np.float64)
    n = lib().oracle_exact_search(_p(S), ctypes.c_int64(S.shape[0]), ctypes.c_int32(S.shape[1]),
                                  ctypes.c_int64(S.shape[1]), _p(q), ctypes.c_int32(k),
                                  ctypes.c_int(int(clamp)), _p(ids), _p(dists))
    return ids[:n], dists[:n]


def exact_search_batch(S, Q, k, n_threads=1, clamp=True):
    S = np.ascontiguousarray(S, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    nq = Q.shape[0]
    ids = np.full((nq, k), -1, np.int32)
    dists = np.full((nq, k), np.inf, np.float64)
    lib().oracle_exact_search_batch(_p(S), ctypes.c_int64(S.shape[0]), ctypes.c_int32(S.shape[1]),
                                    ctypes.c_int64(S.shape[1]), _p(Q), ctypes.c_int64(nq),
                                    ctypes.c_int32(k), ctypes.c_int(int(clamp)),
                                    ctypes.c_int32(n_threads), _p(ids), _p(dists))
    return ids, dists


def index_accumulate(row_off, passing, bucket, sign, idf, sample, cov, dim, max_sample_id,
                     capacity):
    """Sequential first-seen id assignment + scatter-add.  Returns
    (id_of_sample int32[max_sample_id+1], n_kept, acc float64[n_kept x dim])."""
    row_off = np.ascontiguousarray(row_off, np.int64)
    passing = np.ascontiguousarray(passing, np.uint8)
    bucket = np.ascontiguousarray(bucket, np.int32)
    sign = np.ascontiguousarray(sign, np.int8)
    idf = np.ascontiguousarray(idf, np.float64)
    sample = np.ascontiguousarray(sample, np.int32)
    cov = np.ascontiguousarray(cov, np.int32)
    id_of = np.full(max_sample_id + 1, -1, np.int32)
    acc = np.zeros((capacity, dim), np.float64)
    n_kept = lib().oracle_index_accumulate(_p(row_off), _p(passing), _p(bucket), _p(sign), _p(idf),
                                           ctypes.c_int64(len(row_off) - 1), _p(sample), _p(cov),
                                           ctypes.c_int32(dim), _p(id_of), ctypes.c_int32(0), _p(acc))
    return id_of, int(n_kept), acc[:n_kept]
