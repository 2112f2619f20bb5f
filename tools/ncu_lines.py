#!/usr/bin/env python3
"""Per-source-line share of executed instructions and stall samples of one kernel in an .ncu-rep captured with
--import-source on (needs `ncu` on PATH; runs here, no GPU).  usage: python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"],
                         check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    lines = [r for r in rows if len(r) > ii and r[0].isdigit()]
    tot_i = sum(int(r[ii]) for r in lines) or 1
    tot_s = sum(int(r[si]) for r in lines) or 1
    print(f"{kern}: {tot_i} warp instructions, {tot_s} samples")
    for r in sorted(lines, key=lambda r: -int(r[si]))[:top]:
        print(f"{int(r[0]):5d} {100 * int(r[ii]) / tot_i:5.1f}% inst {100 * int(r[si]) / tot_s:5.1f}% samples  {r[1].strip()[:120]}")


if __name__ == "__main__":
    main()
