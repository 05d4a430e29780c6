#!/usr/bin/env python
"""Selected raw-page metrics of one ncu --set full capture as text (the format of profiles/r0x_ncu_*.txt).
   python tools/ncu_raw_summary.py gpurun_out/prof_grad.ncu-rep "header line" > profiles/r02_ncu_grad.txt"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(FBSP\.TriageCompute\.dram__throughput|dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|launch__(block_size|grid_size|"
                  r"occupancy_limit_\w+|registers_per_thread\w*)|lts__t_sector_hit_rate\.pct|sm__inst_executed_pipe_tensor_subpipe_dmma\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|"
                  r"smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|syslts__t_sector_hit_rate\.pct)")


def main(rep, header):
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, units, vals = raw[0], raw[1], raw[2]
    ix = {h: i for i, h in enumerate(hdr)}
    print(header)
    print("kernel: %s\n" % vals[ix["Kernel Name"]][:200])
    for i, h in enumerate(hdr):
        if KEEP.match(h):
            print("%-100s %-18s %s" % (h, units[i], vals[i]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
