"""MornaIndex -- host-side mirror of the reference class (morna.py:146-520) whose
per-pair Python loop runs as CUDA kernels on the B200.

Same constructor arguments, same ``add_junction`` / ``build`` / ``save`` surface,
same errors.  ``add_junction`` only records the row (threshold test, cumulative
frequency, key bytes); ``build`` ships the rows to the device once as binary CSR
and runs hash -> first-seen ids -> order-faithful scatter-add -> float32 store.
Annoy's forest (``n_trees``) is not built; the junctions-by-sample shards are written by
``morna_b200.junctions`` when ``junction_shards`` is set, the metadata table by ``files.write_meta``.
"""
import numpy as np
import torch

from . import _lib, files
from .parse import RowBatch


def _pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.pin_memory() if t.numel() and torch.cuda.is_available() else t


def round_up(x, m):
    return (x + m - 1) // m * m


class MornaIndex(object):
    def __init__(self, sample_count, basename, dim=3000, sample_threshold=100,
                 metafile=None, buffer_size=1024, device=None, store_skipped_rows=False, junction_shards=False):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.sample_count = sample_count                 # morna.py:168
        self.sample_count_is_exact = False               # True once count_samples produced it from this very input
        self.basename = basename
        self.internal_id_map = {}                        # :178
        self.new_internal_id = 0                         # :179
        self.sample_frequencies = {}                     # :182 (saved as defaultdict(int))
        self.metafile = metafile
        self.dim = self.dimension_count = dim            # :190
        self.sample_threshold = sample_threshold         # :194
        self.skipped = 0                                 # :197
        self.buffer_size = buffer_size
        self.junc_id = -1                                # :215
        self._rows = RowBatch()
        self._shards = None          # junctions-by-sample shards (morna.py:200-219), written by save()
        if junction_shards:
            from .junctions import ShardRecorder, remove_shards
            remove_shards(basename)
            self._shards = ShardRecorder(basename, buffer_size, self.device)
        self._store_skipped = store_skipped_rows   # ship under-threshold rows too (kernel tests)
        self._pass = []
        self._running_freq = []
        self.vectors = None          # device float32 [n_kept x ld]
        self.ld = round_up(dim, 4)
        self.row_hash = None         # device (raw, bucket, sign) of the rows, kept for inspection

    # ------------------------------------------------------------------ rows in
    def add_junction(self, junction, samples, coverages):
        """morna.py:344-388: threshold filter, cumulative frequency; the hash, idf and
        per-pair scatter-add happen on the device in build()."""
        self.junc_id += 1
        if self._shards is not None:                     # :359, before the threshold
            self._shards.add_row(self.junc_id, samples, coverages)
        n = len(samples)
        if n < self.sample_threshold:                    # :361-363
            self.skipped += 1
            if self._store_skipped:
                self._rows.add(junction, samples, coverages)
                self._pass.append(0)
                self._running_freq.append(0)
            return
        freq = self.sample_frequencies.get(junction, 0) + n   # :365
        self.sample_frequencies[junction] = freq
        self._rows.add(junction, samples, coverages)
        self._pass.append(1)
        self._running_freq.append(freq)

    def add_lines(self, lines, verbose=False, out=None):
        """go_index's row loop (morna.py:841-861) over text rows."""
        from .parse import tokenize_line
        for i, line in enumerate(lines):
            if verbose and out is not None and i % 1000 == 0:
                out.write("%d lines into index making\r" % i)
                out.flush()
            self.add_junction(*tokenize_line(line))

    def add_text(self, buf, seen_samples=None, n_threads=None):
        """The same row loop over a bytes block of whole lines, tokenised natively
        (csrc/tokenize.cpp): the threshold test and the cumulative frequency stay here
        (morna.py:361-365), rows go to the CSR store in bulk.  Rows the native tokenizer does not
        take (anything not plainly canonical) go through ``tokenize_line`` in their place.
        ``seen_samples``: optional set updated with the distinct sample-id strings of every row
        (count_samples, morna.py:809-822) so one pass over the file serves both."""
        from . import parse
        keys, key_off, row_off, sample, cov, line_off, needs = parse.tokenize_buffer(buf, n_threads)
        n = len(needs)
        odd = np.nonzero(needs)[0].tolist()
        start = 0
        for stop in odd + [n]:
            if stop > start:
                self._add_block(keys, key_off, row_off, sample, cov, start, stop, seen_samples)
            if stop < n:
                line = bytes(buf[line_off[stop]:line_off[stop + 1]]).decode("utf-8")
                if seen_samples is not None:
                    seen_samples.update(line.split("\t")[-2].split(","))
                self.add_junction(*parse.tokenize_line(line))
            start = stop + 1

    def _add_block(self, keys, key_off, row_off, sample, cov, r0, r1, seen_samples):
        lens = np.diff(row_off[r0:r1 + 1])
        passing = lens >= self.sample_threshold
        running = np.zeros(r1 - r0, dtype=np.int64)
        kbytes = keys.tobytes()
        freqs = self.sample_frequencies
        for i in np.nonzero(passing)[0].tolist():        # one dict update per passing row (morna.py:365)
            key = kbytes[key_off[r0 + i]:key_off[r0 + i + 1]].decode("utf-8")
            freq = freqs.get(key, 0) + int(lens[i])
            freqs[key] = freq
            running[i] = freq
        p0, p1 = int(row_off[r0]), int(row_off[r1])
        if self._shards is not None:
            self._shards.add_rows(self.junc_id + 1, lens, sample[p0:p1], cov[p0:p1])
        self.junc_id += r1 - r0
        self.skipped += int((~passing).sum())
        if seen_samples is not None and p1 > p0:         # canonical integers: distinct strings == distinct values
            seen_samples.update(np.unique(sample[p0:p1]).astype(str).tolist())
        if self._store_skipped or passing.all():
            self._rows.add_block(keys[key_off[r0]:key_off[r1]], np.diff(key_off[r0:r1 + 1]), lens, sample[p0:p1], cov[p0:p1])
            self._pass.extend(passing.astype(np.uint8).tolist())
            self._running_freq.extend(running.tolist())
        else:                                            # rows under the threshold are not shipped
            rows = np.nonzero(passing)[0]
            pair_mask = np.repeat(passing, lens)
            key_lens = np.diff(key_off[r0:r1 + 1])
            key_mask = np.repeat(passing, key_lens)
            self._rows.add_block(keys[key_off[r0]:key_off[r1]][key_mask], key_lens[rows], lens[rows],
                                 sample[p0:p1][pair_mask], cov[p0:p1][pair_mask])
            self._pass.extend([1] * len(rows))
            self._running_freq.extend(running[rows].tolist())

    # ------------------------------------------------------------------ device build
    def build(self, n_trees=None, verbose=False, id_range=None, stream=None):
        """Runs the device pipeline.  ``n_trees`` is accepted for signature parity and
        ignored (no Annoy forest).  ``id_range=(lo, hi)`` keeps only that slice of
        internal ids on this GPU (multi-GPU index build); ids and the map are global."""
        if stream is not None and stream != torch.cuda.current_stream(self.device):
            with torch.cuda.stream(stream):              # copies, allocations, kernels and host reads all on `stream`
                return self.build(n_trees, verbose, id_range)
        lib, dev = self.lib, self.device
        packed, key_off, row_off, sample, cov = self._rows.finish()
        n_rows, nnz = len(self._rows), int(row_off[-1])
        if n_rows == 0 or nnz == 0:
            raise ValueError("No internal ids were assigned, indicating that no samples were added "
                             "to the index. Likely caused when no junctions pass the sample threshold.")
        if sample.min() < 0:
            raise ValueError("negative sample id")
        max_sample_id = int(sample.max())
        passing = np.asarray(self._pass, dtype=np.uint8)
        running = np.asarray(self._running_freq, dtype=np.int64)
        idf = np.empty(n_rows, dtype=np.float64)
        _lib.check(lib.morna_idf_host(running.ctypes.data, passing.ctypes.data, n_rows,
                                      int(self.sample_count), idf.ctypes.data), "morna_idf_host")

        with torch.cuda.device(dev):
            sp = _lib.stream_ptr()
            to_dev = lambda a: _pinned(a).to(dev, non_blocking=True)
            d_keys, d_key_off = to_dev(packed), to_dev(key_off)
            d_row_off, d_sample, d_cov = to_dev(row_off), to_dev(sample), to_dev(cov)
            d_pass, d_idf = to_dev(passing), to_dev(idf)
            d_raw = torch.empty(n_rows, dtype=torch.int32, device=dev)
            d_bucket = torch.empty(n_rows, dtype=torch.int32, device=dev)
            d_sign = torch.empty(n_rows, dtype=torch.int8, device=dev)
            _lib.check(lib.morna_hash_junctions(_lib.dev_ptr(d_keys), _lib.dev_ptr(d_key_off), n_rows, self.dim,
                                                _lib.dev_ptr(d_raw), _lib.dev_ptr(d_bucket), _lib.dev_ptr(d_sign), sp),
                       "morna_hash_junctions")
            self.row_hash = (d_raw, d_bucket, d_sign)

            d_id_of = torch.empty(max_sample_id + 1, dtype=torch.int32, device=dev)
            d_n_kept = torch.zeros(1, dtype=torch.int32, device=dev)
            ws = _lib.workspace(lib.morna_assign_internal_ids_workspace_bytes(n_rows, nnz, max_sample_id), dev)
            _lib.check(lib.morna_assign_internal_ids(
                _lib.dev_ptr(d_row_off), _lib.dev_ptr(d_pass), n_rows, _lib.dev_ptr(d_sample), nnz, max_sample_id,
                int(self.sample_count) if self.sample_count_is_exact else 0, _lib.dev_ptr(d_id_of), _lib.dev_ptr(d_n_kept), _lib.dev_ptr(ws), ws.numel(), sp),
                "morna_assign_internal_ids")
            n_kept = int(d_n_kept.item())
            if n_kept == 0:                              # morna.py:399-403
                _lib.check(_lib.ERR_NO_SAMPLES, "build")
            self.new_internal_id = n_kept
            id_of = d_id_of.cpu().numpy()
            seen = np.nonzero(id_of >= 0)[0]
            self.internal_id_map = dict(zip(seen.tolist(), id_of[seen].tolist()))

            lo, hi = (0, n_kept) if id_range is None else (max(0, id_range[0]), min(n_kept, id_range[1]))
            self.id_range = (lo, hi)
            width = max(hi - lo, 0)
            acc_ld = max(round_up(width, 32), 32)
            d_acc = torch.empty(self.dim * acc_ld, dtype=torch.float64, device=dev)
            ws = _lib.workspace(lib.morna_index_accumulate_workspace_bytes(n_rows, nnz, self.dim), dev)
            _lib.check(lib.morna_index_accumulate(
                _lib.dev_ptr(d_row_off), _lib.dev_ptr(d_pass), _lib.dev_ptr(d_bucket), _lib.dev_ptr(d_sign),
                _lib.dev_ptr(d_idf), n_rows, _lib.dev_ptr(d_sample), _lib.dev_ptr(d_cov), nnz,
                _lib.dev_ptr(d_id_of), max_sample_id, lo, hi, self.dim, _lib.dev_ptr(d_acc), acc_ld,
                _lib.dev_ptr(ws), ws.numel(), sp), "morna_index_accumulate")
            self.vectors = torch.empty((width, self.ld), dtype=torch.float32, device=dev)
            _lib.check(lib.morna_round_store(_lib.dev_ptr(d_acc), acc_ld, width, self.dim,
                                             _lib.dev_ptr(self.vectors), self.ld, sp), "morna_round_store")
            self._acc = d_acc
            self._acc_ld = acc_ld
            if verbose:
                import sys
                sys.stderr.write("\nAdded a total of %d samples to the index.\n" % n_kept)
                sys.stderr.write("%d junctions skipped for not meeting sample threshold\n" % self.skipped)
        return self

    # ------------------------------------------------------------------ accessors
    def get_n_items(self):
        return self.new_internal_id

    def matrix_f32(self):
        """Host copy of the stored rows [hi-lo x dim] (what add_item holds, morna.py:406)."""
        return self.vectors[:, :self.dim].cpu().numpy()

    def accumulator_f64(self):
        """Host copy of the double accumulator as [hi-lo x dim] (pre-rounding cells)."""
        lo, hi = self.id_range
        a = self._acc.view(self.dim, self._acc_ld)[:, :hi - lo]
        return a.t().contiguous().cpu().numpy()

    def get_item_vector(self, internal_id):
        lo, hi = self.id_range
        return self.vectors[internal_id - lo, :self.dim].cpu().tolist()

    def save(self, basename):
        """morna.py:427-455: .stats/.freq/.map as the reference writes them, plus the
        dense vector store basename.vec.mor in place of basename.annoy.mor."""
        if self.vectors is None:
            raise RuntimeError("build() must run before save()")
        if self.id_range != (0, self.new_internal_id):
            raise RuntimeError("save() needs the full id range on this device")
        files.write_vectors(basename, self.matrix_f32())
        files.write_stats(basename, self.sample_count, self.new_internal_id, self.dim)
        files.write_freq(basename, self.sample_frequencies)
        files.write_map(basename, self.internal_id_map)
        if self._shards is not None:                       # morna.py:457-488
            self._shards.write()
        if self.metafile:                                  # morna.py:494-520
            files.write_meta(basename, self.metafile)


def go_index(intropolis, basename, features, n_trees, sample_count, sample_threshold, buffer_size,
             verbose, metafile, out=None, junction_shards=True):
    """morna.py:824-865.  ``junction_shards=False`` skips the junctions-by-sample shards (only ``junctions`` reads them)."""
    import sys
    out = out or sys.stdout
    from . import parse
    seen = None if sample_count else set()              # one pass serves count_samples too (morna.py:789-822)
    index = MornaIndex(sample_count or 0, basename, dim=features, sample_threshold=sample_threshold,
                       metafile=metafile, buffer_size=buffer_size, junction_shards=junction_shards)
    done = 0
    with parse.open_intropolis_binary(intropolis) as fh:
        for block in parse.read_blocks(fh):
            index.add_text(block, seen_samples=seen)
            if verbose:
                done += block.count(b"\n")
                out.write("%d lines into index making\r" % done)
                out.flush()
    if seen is not None:
        index.sample_count = len(seen)
        index.sample_count_is_exact = True               # distinct id strings >= distinct ids: the id pass may stop early
    if verbose:
        out.write("\nThere are %d samples.\n" % index.sample_count)
    if verbose:
        out.write("Finished making index; now building\n")
    index.build(n_trees, verbose=verbose)
    index.save(basename)
    return index
