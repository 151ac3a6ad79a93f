"""bench.py's single-query section (one at a time, two in flight, end to end) with the L2 prefetch of the first-pass rows
on and off (morna_debug_set_tuning key 27), alternating in one process."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from morna_b200 import _lib, synth
lib = _lib.load()
dev = torch.device("cuda:0")
peaks = bench.measured_peaks()
for rep in range(3):
    for v in [int(a) for a in sys.argv[1:]] or (0, 6144):
        lib.morna_debug_set_tuning(27, v)
        r = bench.single_query_line(torch, lib, _lib, synth, dev, peaks)
        print("prefetch %7d: one at a time %.1f us, two in flight %.1f us, e2e %.1f us" % (
            v, r["us_per_query"], r["two_in_flight"]["us_per_query"], r["e2e"]["us_per_query"]), flush=True)
lib.morna_debug_set_tuning(27, 6144)
