#!/usr/bin/env python
"""ncu CSV (per-launch gpu__time_duration, DMMA sub-pipe %, DRAM bytes of every GEMM launch) -> markdown table by kernel
instance and grid class, with the time-weighted DMMA activity.    python tools/summarize_gemm_launches.py gpurun_out/gemm_launches.csv"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "ID")
    ix = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rows:
        if len(r) != len(hdr) or r[0] == "ID":
            continue
        d = launches.setdefault(r[ix["ID"]], {"name": re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", ""),
                                              "grid": int(re.findall(r"\d+", r[ix["Grid Size"]])[0])})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        m, u = r[ix["Metric Name"]], r[ix["Metric Unit"]]
        if m == "gpu__time_duration.sum":
            d["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1e-3)
        elif "dmma" in m:
            d["dmma"] = v
        elif m.startswith("dram__bytes"):
            d["dram"] = d.get("dram", 0.0) + v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    L = [d for d in launches.values() if "us" in d]
    tot = sum(d["us"] for d in L)
    wavg = sum(d["us"] * d.get("dmma", 0.0) for d in L) / tot
    print("# Round 2 — FP64 tensor-pipe utilisation of every GEMM launch of one n=8192 evaluation\n")
    print("`tools/gpu_profile_gemms.sh` (ncu --metrics gpu__time_duration.sum, sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active, "
          "dram__bytes_read.sum, dram__bytes_write.sum -k regex:gemm_kernel -s 1400 -c 460; pipelined plan, plain streams under ncu, serialised, cold caches).\n")
    print("%d launches, %.2f ms summed, **time-weighted DMMA sub-pipe utilisation %.1f %%**, DRAM traffic %.2f GB.\n" %
          (len(L), tot * 1e-3, wavg, sum(d.get("dram", 0.0) for d in L) / 1e9))
    agg = collections.OrderedDict()
    for d in L:
        cls = ">=592 (bulk)" if d["grid"] >= 592 else ("64-591" if d["grid"] >= 64 else "<64 CTAs")
        a = agg.setdefault((d["name"], cls), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d["us"]; a[2] += d["us"] * d.get("dmma", 0.0); a[3] += d.get("dram", 0.0)
    print("| kernel | grid class | launches | total us | share | DMMA pipe % (time-weighted) | DRAM GB |\n|---|---|---:|---:|---:|---:|---:|")
    for (nm, cls), (c, us, w, dr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %s | %d | %.0f | %.3f | %.1f | %.2f |" % (nm, cls, c, us, us / tot, w / us, dr / 1e9))
    bulk = [d for d in L if d["grid"] >= 592]
    if bulk:
        bu = sum(d["us"] for d in bulk)
        print("\nBulk launches (>= 592 CTAs) alone: %d launches, %.2f ms, time-weighted DMMA %.1f %%." %
              (len(bulk), bu * 1e-3, sum(d["us"] * d.get("dmma", 0.0) for d in bulk) / bu))


if __name__ == "__main__":
    main(sys.argv[1])
