"""Junctions-by-sample shards and the `junctions` subcommand (morna.py:221-341, 457-488,
1486-1632) against oracle/junction_db_oracle.py, plus the reference's integrity checkers restated as properties
(tests/shard_check.py, tests/junction_integrity.py, tests/interpret_rle.py)."""
import gzip
import io
import os
import random
import re
import sqlite3
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import junction_db_oracle as jo
from oracle import morna_oracle as mo
from morna_b200 import junctions as mj
from tests.helpers import GOLDEN, tiny_lines


def test_base64_run_lengths():
    # digits '0'..'o' (morna.py:38-75); -1 is the gap of a sample listed twice in one row
    assert [mj.encode_64(v) for v in (0, 9, 10, 63, 64, 4095, 4096, -1)] == ["0", "9", ":", "o", "10", "oo", "100", "o"]
    for v in list(range(0, 5000, 7)) + [64 ** 3 - 1, 64 ** 3, 10 ** 9]:
        assert mj.encode_64(v) == jo.encode_64(v)
        assert mj.decode_64(mj.encode_64(v)) == v == jo.decode_64(jo.encode_64(v))
        assert mj.increment_64(mj.encode_64(v)) == jo.increment_64(jo.encode_64(v)) == mj.encode_64(v + 1)
    for text in ("!5!1!2!3", "!1", "!2!7!1", "", "!o!o"):
        assert list(mj.running_sum(text.split("!"))) == list(jo.running_sum(text.split("!")))
    assert list(mj.running_sum("!5!2!3!1".split("!"))) == [5, 6, 10]


def test_python2_list_size_model():
    # sys.getsizeof([]) == 72 and the over-allocation sequence of CPython 2.7's list.append
    buf = mj._Py2Buffer()
    assert buf.sizeof() == 72
    grown = []
    for i in range(130):
        buf.append("x")
        if not grown or grown[-1] != buf.allocated:
            grown.append(buf.allocated)
    assert grown[:12] == [4, 8, 16, 25, 35, 46, 58, 72, 88, 106, 126, 148]
    # the default 1024-byte buffer is cut when the 107th token arrives
    buf = mj._Py2Buffer()
    n = 0
    while buf.sizeof() <= 1024:
        buf.append("x")
        n += 1
    assert n == 107


def random_rows(rng, n_rows, n_samples, density, dup=False):
    rows = []
    for _ in range(n_rows):
        k = rng.randint(0, max(1, int(n_samples * density)))
        samples = rng.sample(range(n_samples), k)
        if rng.random() < 0.7:
            samples.sort()
        if dup and samples and rng.random() < 0.3:
            samples.append(samples[0])
        rows.append((samples, [rng.randint(1, 40) for _ in samples]))
    return rows


def transposed(rows):
    per, created = {}, []
    for j, (samples, covs) in enumerate(rows):
        for s, c in zip(samples, covs):
            if s not in per:
                per[s] = ([], [])
                created.append(s)
            per[s][0].append(j)
            per[s][1].append(c)
    return created, per


@pytest.mark.parametrize("buffer_size,density,dup", [(1024, 0.5, False), (1024, 0.9, False), (200, 0.5, False),
                                                      (104, 0.8, False), (96, 0.6, False), (1024, 0.7, True), (136, 0.95, True)])
def test_sample_rows_equal_the_interleaved_reference_loop(buffer_size, density, dup):
    rng = random.Random(buffer_size * 31 + int(density * 100))
    rows = random_rows(rng, 400, 23, density, dup)
    oracle = jo.JunctionDbOracle(buffer_size)
    for samples, covs in rows:
        oracle.add_junction(samples, covs)
    want = oracle.finish()
    created, per = transposed(rows)
    assert created == oracle.created
    cut_rows = 0
    for s in created:
        got = mj.sample_rows(per[s][0], per[s][1], buffer_size)
        assert got == want[s], "sample %d" % s
        cut_rows += len(got) > 1
        assert mj.sample_rows(per[s][0], per[s][1], buffer_size) == got          # no hidden state
    assert cut_rows > 0 or buffer_size >= 1024


def ones_from_the_end(text):
    """tests/junction_integrity.py: runs parsed from the end, the last one is a run of present junctions."""
    ones = total = 0
    i = 0
    while text:
        m = re.search("(.*)!([0-o]+)$", text)
        total += jo.decode_64(m.group(2))
        if i % 2 == 0:
            ones += jo.decode_64(m.group(2))
        i += 1
        text = m.group(1)
    return ones, total


def test_fixture_tables_pass_the_reference_integrity_check():
    rows = [mo.tokenize_line(line)[1:] for line in tiny_lines()]
    created, per = transposed(rows)
    assert len(created) == 6850
    for s in created:
        table = mj.sample_rows(per[s][0], per[s][1])
        text = "".join(r[0] for r in table)
        n_cov = sum(len(r[1].strip(",").split(",")) for r in table)
        ones, total = ones_from_the_end(text)
        assert ones == n_cov == len(per[s][0])              # junction_integrity.py: num_pos_juncs == num_covs
        assert total <= len(rows)                           # ... and min_num_juncs <= junctions in the input


def test_retained_junctions_and_splice_lines_equal_the_oracle(tmp_path):
    rng = random.Random(7)
    for trial in range(30):
        n_results = rng.randint(1, 12)
        result_juncs, result_covrs = [], []
        for _ in range(n_results):
            juncs = sorted(rng.sample(range(60), rng.randint(1, 40)))
            result_juncs.append(juncs)
            result_covrs.append([str(rng.randint(1, 9)) for _ in juncs])
        ff, cf = rng.choice([0.05, 0.3, 0.5, 1.0]), rng.choice([1, 5, 8, 100])
        got, got_map = mj.retained_junctions(result_juncs, result_covrs, ff, cf)
        want, want_map = jo.retained_junctions(result_juncs, result_covrs, ff, cf)
        assert got == want
        assert {k: v for k, v in got_map.items() if v} == {k: v for k, v in want_map.items() if v}
        if not want:
            continue
        ids = [rng.randint(0, 50) for _ in range(n_results)]
        lines = []
        for j in range(60):
            samples = sorted(rng.sample(range(51), rng.randint(1, 30)))
            lines.append("chr1\t%d\t%d\t+\tGT\tAG\t%s\t%s\n" % (100 + j, 200 + j, ",".join(map(str, samples)),
                                                                 ",".join(str(rng.randint(1, 30)) for _ in samples)))
        gz = tmp_path / ("j%d.gz" % trial)
        with gzip.open(gz, "wt") as fh:
            fh.writelines(lines)
        out = tmp_path / ("s%d.txt" % trial)
        err = io.StringIO()
        mj.write_splicefile(str(out), str(gz), got, got_map, ids, err)
        assert out.read_text() == "".join(jo.splice_lines(lines, want, want_map, ids))
        assert err.getvalue() == "%d junctions to begin with\n" % len(want)
    with pytest.raises(IndexError):                          # nothing retained: ordered_junctions.pop(0) (morna.py:1588)
        mj.write_splicefile(str(tmp_path / "none.txt"), str(gz), [], {}, [], io.StringIO())
    with pytest.raises(IndexError):                          # more coverages than junction indexes (morna.py:1567)
        mj.retained_junctions([[3]], [["9", "9"]], 1.0, 5)


# ------------------------------------------------------------------ on the GPU: the shard files and the subcommand
@pytest.mark.gpu
def test_index_writes_the_shards_the_reference_loop_would(tmp_path):
    from morna_b200 import cli
    base = str(tmp_path / "tiny")
    stale = mj.shard_path(base, 7)
    open(stale, "w").write("left over from an older index")          # morna.py:200-205 removes it
    assert cli.main(["index", "--intropolis", os.path.join(GOLDEN, "tiny_intropolis.tsv"), "-x", base, "-b", "256"],
                    stdout=io.StringIO()) == 0
    oracle = jo.JunctionDbOracle(256)
    for line in tiny_lines():
        _, samples, covs = mo.tokenize_line(line)
        oracle.add_junction(samples, covs)
    want = oracle.finish()
    seen = {}
    for shard_id in range(100):
        path = mj.shard_path(base, shard_id)
        if not os.path.exists(path):
            continue
        conn = sqlite3.connect(path)
        names = [r[0] for r in conn.execute("SELECT name FROM sqlite_master WHERE type='table'")]
        for name in names:
            sample_id = int(re.search("sample_([0-9]*)$", name).group(1))
            # tests/shard_check.py: the table sits in the shard its sample id hashes to
            assert mo.murmur3_x86_32(str(sample_id).encode()) % 100 == shard_id
            seen[sample_id] = [list(r) for r in conn.execute("SELECT * FROM %s" % name)]
        conn.close()
    assert seen == want
    created_by_shard = {}
    for s in oracle.created:
        created_by_shard.setdefault(mo.murmur3_x86_32(str(s).encode()) % 100, []).append(s)
    conn = sqlite3.connect(mj.shard_path(base, 7))                    # tables in creation order
    names = [int(r[0][7:]) for r in conn.execute("SELECT name FROM sqlite_master WHERE type='table'")]
    conn.close()
    assert names == created_by_shard[7]
    # --no-junction-shards leaves none behind
    base2 = str(tmp_path / "bare")
    assert cli.main(["index", "--intropolis", os.path.join(GOLDEN, "tiny_intropolis.tsv"), "-x", base2,
                     "--no-junction-shards"], stdout=io.StringIO()) == 0
    assert not [f for f in os.listdir(str(tmp_path)) if f.startswith("bare.sh")]


@pytest.mark.gpu
def test_recorder_transposes_blocks_and_single_rows_like_the_reference_loop(tmp_path):
    """Many rows, unsorted and repeated samples, small buffers: table rows get cut and runs continue in the table."""
    import numpy as np
    rng = random.Random(11)
    rows = random_rows(rng, 500, 37, 0.8, dup=True)
    for buffer_size in (136, 1024):
        oracle = jo.JunctionDbOracle(buffer_size)
        rec = mj.ShardRecorder(str(tmp_path / ("b%d" % buffer_size)), buffer_size)
        j = 0
        while j < len(rows):                                  # alternate block adds and single-row adds
            if rng.random() < 0.5:
                rec.add_row(j, rows[j][0], rows[j][1])
                j += 1
            else:
                block = rows[j:j + rng.randint(1, 9)]
                rec.add_rows(j, [len(r[0]) for r in block], np.array(sum((r[0] for r in block), []), dtype=np.int32),
                             np.array(sum((r[1] for r in block), []), dtype=np.int32))
                j += len(block)
        for samples, covs in rows:
            oracle.add_junction(samples, covs)
        want = oracle.finish()
        created, tables = rec.tables()
        assert created == oracle.created
        assert tables == want
        assert buffer_size == 1024 or any(len(t) > 1 for t in tables.values())
        assert rec.write() == len(created)
        got = {}
        for shard_id in range(100):
            path = mj.shard_path(rec.basename, shard_id)
            if os.path.exists(path):
                conn = sqlite3.connect(path)
                for (name,) in list(conn.execute("SELECT name FROM sqlite_master WHERE type='table'")):
                    got[int(name[7:])] = [list(r) for r in conn.execute("SELECT * FROM %s" % name)]
                    juncs, covrs = mj.read_sample_table(rec.basename, int(name[7:]), shard_id)
                    assert (juncs, covrs) == jo.read_sample(want[int(name[7:])])
                conn.close()
        assert got == want


@pytest.mark.gpu
def test_junctions_subcommand_end_to_end(tmp_path):
    from morna_b200 import cli
    from morna_b200.search import MornaSearch
    base = str(tmp_path / "tiny")
    tsv = os.path.join(GOLDEN, "tiny_intropolis.tsv")
    assert cli.main(["index", "--intropolis", tsv, "-x", base], stdout=io.StringIO()) == 0
    gz = str(tmp_path / "tiny.tsv.gz")
    with open(tsv, "rb") as src, gzip.open(gz, "wb") as dst:
        dst.write(src.read())
    # a first-pass SAM with two spliced reads over the fixture's junctions (intron = junction start .. end)
    sam = tmp_path / "pass1.sam"
    sam.write_text("@HD\tVN:1.0\n"
                   "r1\t0\tchr1\t14800\t60\t30M100N30M\t*\t0\t0\t" + "A" * 60 + "\t" + "I" * 60 + "\tXS:A:-\n"
                   "r2\t0\tchr1\t15009\t60\t30M757N30M\t*\t0\t0\t" + "A" * 60 + "\t" + "I" * 60 + "\tXS:A:-\n")
    splice = str(tmp_path / "splices.txt")
    out, err = io.StringIO(), io.StringIO()
    assert cli.main(["junctions", "-x", base, "-i", "unused", "-p1", str(sam), "-e", "-d", "-r", "6",
                     "--junction-filter", ".5,5", "--junction-file", gz, "-sf", splice], stdout=out, stderr=err) == 0
    text = out.getvalue()
    # the search part prints what `search` prints for the same query
    ref_out = io.StringIO()
    assert cli.main(["search", "-x", base, "-f", "sam", "-e", "-d", "-r", "6"], stdin=open(str(sam)), stdout=ref_out) == 0
    assert text.startswith(ref_out.getvalue())
    internal = [int(line.split("\t")[1]) for line in ref_out.getvalue().splitlines()]
    searcher = MornaSearch(basename=base)
    sample_ids = [searcher.inverse_lookup(i) for i in internal]
    # oracle: tables -> lists -> filter -> splice lines
    oracle = jo.JunctionDbOracle(1024)
    for line in tiny_lines():
        _, samples, covs = mo.tokenize_line(line)
        oracle.add_junction(samples, covs)
    tables = oracle.finish()
    juncs, covrs = zip(*[jo.read_sample(tables[s]) for s in sample_ids])
    ordered, found = jo.retained_junctions(list(juncs), list(covrs), 0.5, 5)
    rest = text[len(ref_out.getvalue()):]
    want_rest = "".join("shard_id is %02d\n" % (mo.murmur3_x86_32(str(s).encode()) % 100) for s in sample_ids)
    want_rest += "result_juncs lengths: \n%s\nresult_covrs lengths: \n%s\n" % ([len(x) for x in juncs], [len(x) for x in covrs])
    want_rest += "Number of retained junctions: %d\n" % len(ordered)
    assert rest == want_rest
    assert open(splice).read() == "".join(jo.splice_lines(tiny_lines(), ordered, found, sample_ids))
    assert ("%d junctions to begin with\n" % len(ordered)) in err.getvalue()
