"""CPU-only checks: the C-ABI library loads and exports the header's symbols,
file formats, parsers, CLI plumbing.  No compute calls (no GPU here)."""
import ctypes
import io
import os
import pickle
import re

import numpy as np
import pytest

from morna_b200 import _lib, cli, files, parse
from tests.helpers import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "morna_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(morna_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.SO_PATH):
        from morna_b200 import build
        build.build()
    lib = ctypes.CDLL(_lib.SO_PATH)
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), name + " declared in include/morna_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), "ctypes table and header disagree"
    assert _lib.load().morna_abi_version() == _lib.ABI_VERSION == 4
    assert _lib.load().morna_status_string(-2) == b"workspace too small"


def test_idf_host_matches_math_log():
    import math
    lib = _lib.load()
    freq = np.array([2040, 6210, 1664, 5, 0], dtype=np.int64)
    ok = np.array([1, 1, 1, 0, 1], dtype=np.uint8)
    out = np.empty(5)
    assert lib.morna_idf_host(freq.ctypes.data, ok.ctypes.data, 5, 6850, out.ctypes.data) == 0
    assert out[:3].tolist() == [math.log(6850.0 / f) for f in (2040, 6210, 1664)]
    assert out[3] == 0.0 and out[4] == 0.0


def test_product_path_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from morna_b200.index import MornaIndex
    with pytest.raises(_lib.MornaLibraryError):
        MornaIndex(10, "x")
    from morna_b200.search import MornaSearch
    with pytest.raises(_lib.MornaLibraryError):
        MornaSearch(vectors=np.zeros((1, 4), np.float32), stats=(1, 1, 4))


def test_product_never_imports_oracle():
    for dirpath, _, names in os.walk(os.path.join(ROOT, "morna_b200")):
        for name in names:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert "oracle" not in text.replace("test oracle", ""), os.path.join(dirpath, name)


def test_index_files_round_trip_and_py2_protocol(tmp_path):
    base = str(tmp_path / "idx")
    files.write_stats(base, 6850, 6850, 3000)
    assert open(base + ".stats.mor").read() == "6850\n6850\n3000\n"
    assert files.read_stats(base) == (6850, 6850, 3000)
    files.write_freq(base, {"chr1 14830 14929": 2040})
    raw = open(base + ".freq.mor", "rb").read()
    assert raw[:2] == b"\x80\x02" and b"defaultdict" in raw          # protocol 2, defaultdict
    freq = files.read_freq(base)
    assert freq["chr1 14830 14929"] == 2040 and freq["absent"] == 0
    files.write_map(base, {12: 0, np.int64(1): np.int32(2040)})
    assert files.read_map(base) == {12: 0, 1: 2040}
    m = np.random.default_rng(0).standard_normal((5, 7)).astype(np.float32)
    files.write_vectors(base, m)
    assert np.array_equal(files.read_vectors(base), m)
    # a Python 2 style pickle (str keys as bytes) still loads
    with open(base + ".map.mor", "wb") as fh:
        fh.write(pickle.dumps({3: 1}, protocol=2))
    assert files.read_map(base) == {3: 1}


def test_annoy_item_rows(tmp_path):
    import struct
    m = np.arange(12, dtype=np.float32).reshape(3, 4)
    blob = b"".join(struct.pack("<iii", 1, 0, 0) + m[i].tobytes() for i in range(3))
    path = str(tmp_path / "a.annoy.mor")
    open(path, "wb").write(blob + b"\0" * 28)
    assert np.array_equal(files.read_annoy_item_vectors(path, 3, 4), m)


def test_tokenizer_and_row_batch():
    key, s, c = parse.tokenize_line("chr10\t101039923\t101040322\t-\tGC\tAG\t1,2,3,4\t1,1,1,5\n")
    assert key == "chr10 101039923 101040322" and s == [1, 2, 3, 4] and c == [1, 1, 1, 5]
    batch = parse.RowBatch()
    batch.add(key, s, c)
    batch.add("chrX 1 2", [7], [9])
    packed, key_off, row_off, samples, covs = batch.finish()
    assert bytes(packed[key_off[1]:key_off[2]]) == b"chrX 1 2"
    assert row_off.tolist() == [0, 4, 5] and samples.tolist() == [1, 2, 3, 4, 7] and covs.tolist() == [1, 1, 1, 5, 9]
    assert parse.count_samples(["a\tb\t1,2\t3\n", "a\tb\t2,03\t3\n"]) == 3   # strings, not ints


def test_query_stream_parsers():
    raw = list(parse.junctions_from_raw_stream(io.StringIO("chr1\t10\t20\t3\n")))
    assert raw == [("chr1", 10, 20, 3)]
    bed = "chr1\t100\t500\tx\t7\t+\t100\t500\t0\t3\t50,60,40,\t0,150,360,\n"
    assert list(parse.junctions_from_bed_stream(io.StringIO(bed))) == [("chr1", 151, 250, 7), ("chr1", 311, 460, 7)]
    sam = ("@HD\tVN:1.0\n"
           "r1\t0\tchr2\t1000\t60\t10M100N5M2D3M50N7M\t*\t0\t0\tACGT\t*\n"
           "r2\t4\tchr2\t1000\t60\t10M100N5M\t*\t0\t0\tACGT\t*\n"
           "r3\t256\tchr2\t1000\t60\t10M100N5M\t*\t0\t0\tACGT\t*\n"
           "r4\t0\tchr2\t1000\t60\t25M\t*\t0\t0\tACGT\t*\n")
    assert list(parse.junctions_from_sam_stream(io.StringIO(sam))) == [("chr2", 1010, 1109, 1), ("chr2", 1120, 1169, 1)]


def test_sam_short_lines_behave_like_the_reference(capsys):
    """utils.py:268-290: an unmapped read is skipped before any other column is read (short lines included); a short
    mapped line prints 'Error found on line' to stderr and the IndexError propagates."""
    ok = "r1\t4\tchr2\n" + "r2\t0\tchr2\t1000\t60\t10M100N5M\t*\t0\t0\tACGT\t*\n"
    assert list(parse.junctions_from_sam_stream(io.StringIO(ok))) == [("chr2", 1010, 1109, 1)]
    bad = "r3\t0\tchr2\t1000\t60\t10M100N5M\n"
    with pytest.raises(IndexError):
        list(parse.junctions_from_sam_stream(io.StringIO(bad)))
    assert "Error found on line: r3" in capsys.readouterr().err


def test_cli_parser_defaults_match_reference():
    p = cli.build_parser()
    a = p.parse_args(["index", "--intropolis", "x.gz"])
    assert (a.basename, a.features, a.n_trees, a.sample_count, a.sample_threshold, a.buffer_size) == \
        ("morna", 3000, 200, None, 100, 1024)
    s = p.parse_args(["search", "-x", "idx", "-e", "-d", "-q", "12"])
    assert (s.results, s.search_k, s.format, s.exact, s.distances, s.query_id) == (20, 100, "sam", True, True, 12)
    out = io.StringIO()
    cli.results_output(([3, 1], [0.0, 1.4142135623730951]), out)
    assert out.getvalue() == "1.\t3\t0.0\n2.\t1\t1.41421356237\n"


# ------------------------------------------------------------------ native tokenizer (host code of the .so, no GPU)
def _python_rows(text):
    from morna_b200 import parse
    return [parse.tokenize_line(line) for line in text.decode().split("\n") if line != ""]


def _native_rows(text, n_threads):
    from morna_b200 import parse
    keys, key_off, row_off, sample, cov, line_off, needs = parse.tokenize_buffer(text, n_threads)
    rows = []
    kb = keys.tobytes()
    for r in range(len(needs)):
        line = text[line_off[r]:line_off[r + 1]]
        if needs[r]:
            assert row_off[r + 1] == row_off[r] and key_off[r + 1] == key_off[r]
            rows.append(("python", line))
        else:
            rows.append((kb[key_off[r]:key_off[r + 1]].decode(), sample[row_off[r]:row_off[r + 1]].tolist(),
                         cov[row_off[r]:row_off[r + 1]].tolist()))
    return rows


@pytest.mark.parametrize("n_threads", [1, 3, 16])
def test_native_tokenizer_matches_python_on_the_reference_fixture(n_threads):
    text = open(os.path.join(GOLDEN, "tiny_intropolis.tsv"), "rb").read()
    want = _python_rows(text)
    got = _native_rows(text, n_threads)
    assert len(got) == len(want) == 3
    assert got == want


def test_native_tokenizer_leaves_anything_unusual_to_python():
    good = b"chr1\t10\t20\t+\tGT\tAG\t1,5,9\t3,2,1\n"
    lines = [
        good,
        b"chr2\t7\t9\t-\tGT\tAG\t4\t11\r\n",                  # CRLF is what strip() removes: native
        b" chr1\t10\t20\t+\tGT\tAG\t1,5\t3,2\n",              # leading blank
        b"chr1\t10\t20\t+\tGT\tAG\t1,5\t3,2\t\n",             # trailing tab
        b"chr1\t10\t20\t+\tGT\tAG\t1,05\t3,2\n",              # leading zero: another count_samples string
        b"chr1\t10\t20\t+\tGT\tAG\t1,5,7\t3,2\n",             # unequal lists
        b"chr1\t10\t20\t+\tGT\tAG\t1,-5\t3,2\n",              # sign
        b"chr1\t10\t20\t+\tGT\tAG\t1, 5\t3,2\n",              # blank inside
        b"chr1\t10\t20\n",                                      # too few fields
        b"chr1\t10\t20\t+\tGT\tAG\t1,5\t3,99999999999\n",     # does not fit int32
        b"chrX\t1\t2\tx\ty\t12,13\t1,1",                        # last line without newline, five fields
    ]
    text = b"".join(lines)
    got = _native_rows(text, 4)
    assert [g[0] == "python" for g in got] == [False, False, True, True, True, True, True, True, True, True, False]
    assert got[0] == ("chr1 10 20", [1, 5, 9], [3, 2, 1])
    assert got[1] == ("chr2 7 9", [4], [11])
    assert got[10] == ("chrX 1 2", [12, 13], [1, 1])
    for g, line in zip(got, lines):
        if g[0] == "python":
            assert g[1] == line                                 # the caller gets the exact bytes of the row back


def test_native_tokenizer_random_rows_and_thread_counts():
    from morna_b200 import parse
    rng = np.random.default_rng(5)
    rows = []
    for j in range(2000):
        n = int(rng.integers(1, 60))
        s = np.sort(rng.choice(100000, size=n, replace=False))
        c = rng.integers(1, 5000, size=n)
        rows.append("chr%d\t%d\t%d\t+\tGT\tAG\t%s\t%s\n" % (rng.integers(1, 23), rng.integers(1, 1e8), rng.integers(1, 1e8),
                                                          ",".join(map(str, s)), ",".join(map(str, c))))
    text = "".join(rows).encode()
    want = _python_rows(text)
    for n_threads in (1, 2, 7, 32):
        assert _native_rows(text, n_threads) == want
    blocks = list(parse.read_blocks(io.BytesIO(text), block_bytes=4096))
    assert b"".join(blocks) == text and all(b.endswith(b"\n") for b in blocks)


def test_metadata_db_round_trip(tmp_path):
    """basename.meta.mor as morna.py:494-520 writes it and :666-676 reads it: numeric ids match through sqlite's REAL
    affinity, keywords keep their newline, a second index run replaces the table, Python 2 prints the row as (u'..',)."""
    from morna_b200 import cli, files
    base = str(tmp_path / "idx")
    meta = tmp_path / "meta.txt"
    meta.write_text("12 liver  adult\n21504\tblood\n7 x\n")
    files.write_meta(base, str(meta))
    assert files.read_meta(base, [12, 21504, 99, 7]) == [("liver  adult\n",), ("blood\n",), None, ("x\n",)]
    meta.write_text("5 only\n")
    files.write_meta(base, str(meta))
    assert files.read_meta(base, [12, 5]) == [None, ("only\n",)]
    assert cli.py2_str(("blood\n",)) == "(u'blood\\n',)" and cli.py2_str(None) == "None"
    meta.write_text("13\n")
    with pytest.raises(IndexError):
        files.write_meta(base, str(meta))


# ------------------------------------------------------------------ bench.py contract (no GPU needed)
def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` times the CPU port on the host cores and prints ONE JSON line with the keys the driver
    reads (metric / unit / config of the GPU arm, impl, cpu_baseline of this run, e2e with zero copy bytes)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("exact kNN queries/s") and "50000 samples x 3000 features" in line["config"]["workload"]
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_gpu_arm_refuses_to_run_without_cuda():
    """No CPU fallback: without a GPU the product arm of bench.py stops with an error instead of timing something else."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "CUDA" in out.stderr or "cuda" in out.stderr
