#!/usr/bin/env python3
"""Per-SASS-instruction share of executed instructions and stall samples of one kernel in an .ncu-rep captured with
--import-source on (needs `ncu` on PATH; runs here, no GPU): the hottest instructions, and the same summed over windows of
consecutive instructions (loops show up as plateaus of equal execution counts).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top] [window]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    win = int(sys.argv[4]) if len(sys.argv) > 4 else 50
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", f"regex:{kern}"],
                         check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    isrc, ii, si = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = [r for r in rows if len(r) > max(ii, si) and r[si].isdigit() and r[ii].isdigit()]
    if len(body) % 2 == 0 and body[0][isrc] == body[len(body) // 2][isrc]:
        body = body[:len(body) // 2]  # some ncu versions print the kernel twice
    tot_i = sum(int(r[ii]) for r in body) or 1
    tot_s = sum(int(r[si]) for r in body) or 1
    print(f"{kern}: {len(body)} SASS instructions, {tot_i} warp instructions executed, {tot_s} samples")
    print("-- hottest instructions (by samples)")
    for n in sorted(range(len(body)), key=lambda n: -int(body[n][si]))[:top]:
        r = body[n]
        print(f"{n:6d} {100 * int(r[ii]) / tot_i:5.2f}% inst {100 * int(r[si]) / tot_s:5.2f}% samples  exec {r[ii]:>10s}  {r[isrc].strip()[:90]}")
    print(f"-- windows of {win} instructions")
    for a in range(0, len(body), win):
        w = body[a:a + win]
        ei, es = sum(int(r[ii]) for r in w), sum(int(r[si]) for r in w)
        if ei:
            print(f"{a:6d}-{a + len(w) - 1:6d} {100 * ei / tot_i:5.1f}% inst {100 * es / tot_s:5.1f}% samples  exec/instr {ei // len(w):>10d}")


if __name__ == "__main__":
    main()
