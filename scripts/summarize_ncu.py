"""Turns ncu exports brought back in gpurun_out/ into the committed summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
  python scripts/summarize_ncu.py full gpurun_out/step_r01.ncu-rep profiles/r01_step_summary
"""
import collections
import csv
import json
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        agg.setdefault(row["Kernel Name"], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as fh:
        fh.write("# ncu launch list (gpu__time_duration.sum, --clock-control none): per-launch device time, "
                 "cold-cache and serialised -- compare SHARES\n\nsource: `%s`\n\n" % src)
        fh.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            fh.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k.split("(")[0][:70], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))
    print(open(dst).read())


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def full(src, dst_prefix):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out, summary = [], {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        rec = {"kernel": name}
        for w in WANT:
            if w in idx:
                rec[w] = "%s %s" % (r[idx[w]], units[idx[w]])
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        rec["dram_bytes_per_launch"] = rd + wr
        out.append(rec)
    with open(dst_prefix + ".md", "w") as fh:
        fh.write("# ncu --set full --clock-control none, one launch of each kernel of a step\n\nsource: `%s`\n\n" % src)
        for rec in out:
            fh.write("## %s\n\n" % rec["kernel"])
            for k, v in rec.items():
                if k != "kernel":
                    fh.write("- %s = %s\n" % (k, v))
            fh.write("\n")
    gemm = [r for r in out if "knn_gemm" in r["kernel"]]
    if gemm:
        big = max(gemm, key=lambda r: float(r["gpu__time_duration.sum"].split()[0]))
        summary["knn_gemm2_filter_dram_bytes"] = big["dram_bytes_per_launch"]
    rr = [r for r in out if "rerank_dist" in r["kernel"]]
    if rr:
        summary["rerank_dist_dram_bytes"] = rr[0]["dram_bytes_per_launch"]
    try:                                   # keys written by other captures (single-query kernel) are kept
        with open(dst_prefix + ".json") as fh:
            old = json.load(fh)
        for k, v in old.items():
            summary.setdefault(k, v)
    except Exception:
        pass
    with open(dst_prefix + ".json", "w") as fh:
        json.dump(summary, fh, indent=1)
    print(open(dst_prefix + ".md").read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
