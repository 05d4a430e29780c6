#!/usr/bin/env python
"""Aggregate the warp-stall samples of one ncu --set full --import-source capture by source line and by SASS opcode.
   python tools/ncu_source_summary.py gpurun_out/prof_grad.ncu-rep [min_pct]"""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, what):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", what], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, min_pct=1.2):
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, units, vals = raw[0], raw[1], raw[2]
    want = ("gpu__time_duration.sum", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active")
    print("## raw page\n")
    for i, h in enumerate(hdr):
        if h in want:
            print("%-90s %-16s %s" % (h, units[i], vals[i]))
        elif "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                if float(vals[i].replace(",", "")) >= 0.1:
                    print("%-90s %-16s %s" % (h, units[i], vals[i]))
            except ValueError:
                pass
    rows = page(rep, "cuda,sass")
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "File Path":
            cur = {"file": r[1], "agg": collections.OrderedDict()}
            blocks.append(cur)
        elif cur is not None and len(r) >= 8:
            try:
                ln, s, inst = int(r[0]), int(r[4] or 0), int(r[7] or 0)
            except ValueError:
                continue
            a = cur["agg"].setdefault(ln, [r[1][:110], 0, 0])
            a[1] += s
            a[2] += inst
    tot = sum(v[1] for b in blocks for v in b["agg"].values())
    print("\n## source lines with >= %.1f %% of the %d warp-stall samples\n" % (min_pct, tot))
    for b in blocks:
        for ln, v in sorted(b["agg"].items(), key=lambda x: -x[1][1]):
            if v[1] >= tot * min_pct / 100.0:
                print("%-16s %5d %6d %5.1f%%  inst %10d  %s" % (b["file"].split("/")[-1][:16], ln, v[1], 100.0 * v[1] / tot, v[2], v[0]))
    rows = page(rep, "sass")
    hdr = next(r for r in rows if r and r[0] == "Address")
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]

    def gi(r, k):
        try:
            return int(r[ix[k]] or 0)
        except (ValueError, KeyError):
            return 0
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("\n## stall reasons (all samples)\n")
    print(sorted(((k, sum(gi(r, k) for r in data)) for k in reasons), key=lambda x: -x[1])[:8])
    byop = collections.Counter()
    for r in data:
        m = re.match(r"\s*(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        byop[m.group(1).split(".")[0] if m else "?"] += gi(r, "Warp Stall Sampling (All Samples)")
    print("\n## samples by SASS opcode\n")
    print(byop.most_common(12))
    print("\n## hottest SASS instructions\n")
    for i in sorted(sorted(range(len(data)), key=lambda i: -gi(data[i], "Warp Stall Sampling (All Samples)"))[:10]):
        r = data[i]
        print("%6d  %s  | %s" % (gi(r, "Warp Stall Sampling (All Samples)"), r[ix["Source"]][:70], {k: gi(r, k) for k in reasons if gi(r, k) > 20}))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.2)
