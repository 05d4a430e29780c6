#!/usr/bin/env python
"""ncu launch list (gpu__time_duration.sum per launch, CSV) -> per-kernel table (markdown)."""
import collections
import csv
import re
import sys


def main(path, title):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "")
        t = float(d["Metric Value"]) * (1e-3 if d["Metric Unit"] == "ns" else 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print("# %s\n" % title)
    print("%d launches captured, %.2f ms summed (cold-cache, serialised by ncu: compare shares, not absolutes)\n" % (len(data), tot * 1e-3))
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f | %.3f |" % (k, c, t, t / c, t / tot))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
