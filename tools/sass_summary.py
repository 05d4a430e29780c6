#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass): the evidence table for which hardware
paths the kernels use -- DMMA (FP64 tensor MMA, `mma.sync.m8n8k4.f64`), LDGSTS (`cp.async`), UTMALDG (TMA,
`cp.async.bulk.tensor`), SYNCS (mbarrier), DFMA/DADD/DMUL (vector FP64), MUFU.  tcgen05 (UTC*MMA / LDTM / STTM) has no f64
kind, so none is expected.    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "causalstump_b200", "csrc", "libcausalstump_b200.so")
OPS = ("DMMA", "DFMA", "DADD", "DMUL", "LDGSTS", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "BAR", "MUFU", "UTC", "LDTM", "STTM", "HMMA")


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    fn, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(.*", "", fn).replace("void ", "")
            counts[fn] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[fn][o] += 1
                    total[o] += 1
                    break
    print("SASS mnemonic counts per kernel of causalstump_b200/csrc/libcausalstump_b200.so (cuobjdump -sass, sm_100a)\n")
    cols = [o for o in OPS if total[o] or o in ("UTC", "LDTM", "STTM", "HMMA", "UTMALDG")]
    print("%-78s " % "kernel" + " ".join("%7s" % c for c in cols))
    for fn, c in counts.items():
        if sum(c.values()) == 0:
            continue
        print("%-78s " % fn[:78] + " ".join("%7d" % c[o] for o in cols))
    print("%-78s " % "TOTAL" + " ".join("%7d" % total[o] for o in cols))
    print("\nUTC*MMA / LDTM / STTM (tcgen05) = 0 by construction: tcgen05.mma has no f64 kind (SURVEY.md 2); the FP64 tensor path "
          "on sm_100a is DMMA.8x8x4.  UTMALDG = TMA loads (cp.async.bulk.tensor) of the covariate tiles in the Gram and "
          "trace-gradient kernels; SYNCS = their mbarrier waits.  LDGSTS = cp.async operand staging of the DMMA GEMM kernel.")


if __name__ == "__main__":
    sys.exit(main())
