"""Index-build kernels on BASELINE configs[4]: synthetic junction rows (default 500 M (sample, coverage) pairs,
21,504 samples, --sample-threshold 100), --features swept; pre-tokenised binary CSR resident in HBM.
Reports per-kernel device time, pairs/s and the fraction of the HBM roofline for the scatter-add
(algorithmic bytes 8*nnz + 40*J + 4*N*D)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from morna_b200 import _lib



def run(pairs=500e6, samples=21504, dims=(500, 1000, 3000, 10000, 30000), threshold=100, cpu_pairs=5e6, splits=(), peak=None,
        quiet=False, tunings=()):
    """Returns {"pairs", "rows", "assign_internal_ids_ms", "per_features": {D: {...}}, "cpu_one_core_pairs_per_s"}."""
    def say(*a, **k):
        if not quiet:
            print(*a, **k)
    result = {"per_features": {}}
    lib = _lib.load()
    dev = torch.device("cuda")
    if peak is None:
        peak = 6550.4
        try:
            peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
    N = samples
    rng = np.random.default_rng(7)
    # row lengths ~ lognormal(5, 1.5) clipped to [1, N]; rows appended until the pair budget is met
    lens = []
    total = 0
    while total < pairs:
        chunk = np.clip(rng.lognormal(5.0, 1.5, size=200000).astype(np.int64), 1, N)
        lens.append(chunk); total += int(chunk.sum())
    lens = np.concatenate(lens)
    cut = int(np.searchsorted(np.cumsum(lens), pairs)) + 1
    lens = lens[:cut]
    J = len(lens)
    row_off = np.zeros(J + 1, np.int64); row_off[1:] = np.cumsum(lens)
    nnz = int(row_off[-1])
    chrom = rng.integers(1, 25, size=J); start = rng.integers(10000, 240000000, size=J); end = start + rng.integers(50, 500000, size=J)
    keys = [("chr%s %d %d" % (c if c < 23 else "XY"[c - 23], s, e)).encode() for c, s, e in zip(chrom, start, end)]
    key_off = np.zeros(J + 1, np.int32); key_off[1:] = np.cumsum([len(k) for k in keys])
    packed = np.frombuffer(b"".join(keys) + b"\0", np.uint8).copy()
    passing = (lens >= threshold).astype(np.uint8)
    running = lens.astype(np.int64)                    # keys are unique: running frequency = row length
    idf = np.empty(J)
    _lib.check(lib.morna_idf_host(running.ctypes.data, passing.ctypes.data, J, N, idf.ctypes.data), "idf")
    say("rows J=%d, pairs nnz=%d (%.0f%% in passing rows), samples N=%d" % (J, nnz, 100.0 * lens[passing == 1].sum() / nnz, N))

    d_row_off = torch.from_numpy(row_off).to(dev); d_lens = torch.from_numpy(lens.astype(np.int32)).to(dev)
    rows_of_pair = torch.repeat_interleave(torch.arange(J, dtype=torch.int32, device=dev), d_lens.to(torch.int64))
    pos = torch.arange(nnz, dtype=torch.int64, device=dev) - d_row_off[rows_of_pair.long()]
    g = torch.Generator(device=dev); g.manual_seed(7)
    base = torch.randint(0, N, (J,), generator=g, device=dev)
    step = torch.randint(0, N // 42, (J,), generator=g, device=dev) * 42 + 1      # coprime to 21504 = 2^10 * 3 * 7
    ids = (base[rows_of_pair.long()] + pos * step[rows_of_pair.long()]) % N + 1                      # distinct ids within a row
    del pos, base, step
    ids = torch.sort(rows_of_pair.long() * (N + 2) + ids)[0]        # intropolis lists each row's samples in ascending order
    d_sample = (ids % (N + 2)).to(torch.int32)
    del ids
    d_cov = (1 + torch.empty(nnz, device=dev).exponential_(0.7, generator=g) * (1 + 30 * (torch.rand(nnz, generator=g, device=dev) < 0.02))).to(torch.int32)
    del rows_of_pair
    d_keys, d_key_off = torch.from_numpy(packed).to(dev), torch.from_numpy(key_off).to(dev)
    d_pass, d_idf = torch.from_numpy(passing).to(dev), torch.from_numpy(idf).to(dev)
    sp = _lib.stream_ptr()

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    d_id_of = torch.empty(N + 1, dtype=torch.int32, device=dev); d_n_kept = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_ids = _lib.workspace(lib.morna_assign_internal_ids_workspace_bytes(J, nnz, N), dev)
    t_ids = timed(lambda: _lib.check(lib.morna_assign_internal_ids(_lib.dev_ptr(d_row_off), _lib.dev_ptr(d_pass), J, _lib.dev_ptr(d_sample), nnz, N, N,
                  _lib.dev_ptr(d_id_of), _lib.dev_ptr(d_n_kept), _lib.dev_ptr(ws_ids), ws_ids.numel(), sp), "ids"))
    n_kept = int(d_n_kept.item())
    nnz_pass = int(lens[passing == 1].sum())
    say("assign_internal_ids: %.3f ms (n_kept=%d)" % (t_ids, n_kept))
    result.update({"pairs": nnz, "pairs_in_passing_rows": nnz_pass, "rows": J, "samples": N, "assign_internal_ids_ms": t_ids,
                   "algorithmic": "8*nnz_passing + 40*J + 4*N*D bytes against the measured HBM copy peak; ms = hash + scatter-add + float32 store, "
                                  "ms_with_ids adds the first-seen id assignment"})
    for dim in list(dims):
        d_raw = torch.empty(J, dtype=torch.int32, device=dev); d_bucket = torch.empty_like(d_raw); d_sign = torch.empty(J, dtype=torch.int8, device=dev)
        t_hash = timed(lambda: _lib.check(lib.morna_hash_junctions(_lib.dev_ptr(d_keys), _lib.dev_ptr(d_key_off), J, dim, _lib.dev_ptr(d_raw),
                       _lib.dev_ptr(d_bucket), _lib.dev_ptr(d_sign), sp), "hash"))
        acc_ld = (n_kept + 31) // 32 * 32
        d_acc = torch.empty(dim * acc_ld, dtype=torch.float64, device=dev)
        ws_acc = _lib.workspace(lib.morna_index_accumulate_workspace_bytes(J, nnz, dim), dev)
        for tune in list(tunings):                      # "key=value": one timed accumulate per setting, defaults restored after
            key, value = (int(x) for x in tune.split("="))
            lib.morna_debug_set_tuning(key, value)
            t_s = timed(lambda: _lib.check(lib.morna_index_accumulate(_lib.dev_ptr(d_row_off), _lib.dev_ptr(d_pass), _lib.dev_ptr(d_bucket), _lib.dev_ptr(d_sign),
                        _lib.dev_ptr(d_idf), J, _lib.dev_ptr(d_sample), _lib.dev_ptr(d_cov), nnz, _lib.dev_ptr(d_id_of), N, 0, n_kept, dim, _lib.dev_ptr(d_acc),
                        acc_ld, _lib.dev_ptr(ws_acc), ws_acc.numel(), sp), "acc"), reps=3)
            say("D=%5d: tuning %s -> accumulate %.2f ms" % (dim, tune, t_s), flush=True)
            lib.morna_debug_set_tuning(key, {12: 10, 25: 1, 8: 3}.get(key, 0))
        for split in list(splits):
            lib.morna_debug_set_tuning(12, split)
            t_s = timed(lambda: _lib.check(lib.morna_index_accumulate(_lib.dev_ptr(d_row_off), _lib.dev_ptr(d_pass), _lib.dev_ptr(d_bucket), _lib.dev_ptr(d_sign),
                        _lib.dev_ptr(d_idf), J, _lib.dev_ptr(d_sample), _lib.dev_ptr(d_cov), nnz, _lib.dev_ptr(d_id_of), N, 0, n_kept, dim, _lib.dev_ptr(d_acc),
                        acc_ld, _lib.dev_ptr(ws_acc), ws_acc.numel(), sp), "acc"), reps=2)
            say("D=%5d: sample-id range width 2^%d -> accumulate %.2f ms" % (dim, split, t_s), flush=True)
            lib.morna_debug_set_tuning(12, 10)
        t_acc = timed(lambda: _lib.check(lib.morna_index_accumulate(_lib.dev_ptr(d_row_off), _lib.dev_ptr(d_pass), _lib.dev_ptr(d_bucket), _lib.dev_ptr(d_sign),
                      _lib.dev_ptr(d_idf), J, _lib.dev_ptr(d_sample), _lib.dev_ptr(d_cov), nnz, _lib.dev_ptr(d_id_of), N, 0, n_kept, dim, _lib.dev_ptr(d_acc),
                      acc_ld, _lib.dev_ptr(ws_acc), ws_acc.numel(), sp), "acc"), reps=2)
        ld = (dim + 3) // 4 * 4
        d_vec = torch.empty((n_kept, ld), dtype=torch.float32, device=dev)
        t_store = timed(lambda: _lib.check(lib.morna_round_store(_lib.dev_ptr(d_acc), acc_ld, n_kept, dim, _lib.dev_ptr(d_vec), ld, sp), "store"))
        algo = 8.0 * nnz_pass + 40.0 * J + 4.0 * n_kept * dim
        total_ms = t_hash + t_acc + t_store
        say("D=%5d: hash %.3f ms, accumulate %.2f ms, round/store %.3f ms | %.2f G pairs/s, %.1f%% of the HBM roofline (%.2f GB algorithmic)"
              % (dim, t_hash, t_acc, t_store, nnz_pass / (total_ms * 1e-3) / 1e9, 100 * algo / (total_ms * 1e-3) / (peak * 1e9), algo / 1e9))
        result["per_features"][str(dim)] = {"ms": total_ms, "ms_with_ids": total_ms + t_ids, "accumulate_ms": t_acc, "store_ms": t_store,
                                            "pairs_per_s": nnz_pass / (total_ms * 1e-3), "frac": algo / (total_ms * 1e-3) / (peak * 1e9),
                                            "frac_with_ids": (algo + 4.0 * nnz) / ((total_ms + t_ids) * 1e-3) / (peak * 1e9),
                                            "algorithmic_bytes": algo}
        del d_acc, d_vec, ws_acc
    # CPU baseline: the C port of the per-pair loop (morna.py:376-388) on a slice of the same stream, one core
    from oracle import c_oracle
    cpu_rows = int(np.searchsorted(row_off, cpu_pairs))
    ro = row_off[:cpu_rows + 1]; sm = d_sample[:ro[-1]].cpu().numpy(); cv = d_cov[:ro[-1]].cpu().numpy()
    raw, bucket, sign = c_oracle.hash_rows([k.decode() for k in keys[:cpu_rows]], 3000)
    t0 = time.perf_counter()
    c_oracle.index_accumulate(ro, passing[:cpu_rows], bucket, sign, idf[:cpu_rows], sm, cv, 3000, N, N)
    dt = time.perf_counter() - t0
    say("CPU (C port of add_junction's pair loop, 1 core, D=3000): %.1f M pairs/s on %d pairs" % (ro[-1] / dt / 1e6, ro[-1]))
    result["cpu_one_core_pairs_per_s"] = float(ro[-1] / dt)
    result["cpu_sample"] = "C port of add_junction's per-pair loop (morna.py:376-388), one core, %d pairs, D=3000" % int(ro[-1])

    return result


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=float, default=500e6)
    ap.add_argument("--samples", type=int, default=21504)
    ap.add_argument("--dims", type=str, default="500,1000,3000,10000,30000")
    ap.add_argument("--threshold", type=int, default=100)
    ap.add_argument("--cpu-pairs", type=float, default=5e6)
    ap.add_argument("--splits", type=str, default="", help="sweep morna_debug_set_tuning key 12 (log2 width of the sample-id ranges of the warp-per-range scatter-add)")
    ap.add_argument("--tunings", type=str, default="", help="comma list of key=value for morna_debug_set_tuning, each timed on the accumulate call")
    a = ap.parse_args()
    run(a.pairs, a.samples, [int(x) for x in a.dims.split(",")], a.threshold, a.cpu_pairs, [int(x) for x in a.splits.split(",") if x],
        tunings=[x for x in a.tunings.split(",") if x])
