#!/usr/bin/env python3
"""Disassembly evidence for profiles/: per kernel of aruco3_b200/libaruco3_b200.so the SASS instruction count and the counts
of the mnemonics that prove (or disprove) what DESIGN.md claims — TMA tensor loads (UTMALDG), bulk copies (UBLKCP), mbarrier
traffic (SYNCS), local-memory spills (STL / LDL), integer dot products (IDP), shuffles, popcounts, atomics — beside ptxas's
registers / spill bytes from the build logs.  No GPU needed.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "aruco3_b200" / "libaruco3_b200.so"
BUILD = ROOT / "aruco3_b200" / "csrc" / "build"
WATCH = ("UTMALDG", "UBLKCP", "SYNCS", "STL", "LDL", "IDP", "SHFL", "POPC", "ATOM", "RED", "LDG", "STG", "LDS", "STS", "BAR", "DFMA", "DMUL", "DADD", "MUFU")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def ptxas_info():
    info = {}
    for log in sorted(BUILD.glob("*_ptxas.log")):
        cur = None
        for ln in log.read_text().splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m:
                cur = m.group(1)
                info[cur] = {}
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m and cur:
                info[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
            m = re.search(r"Used (\d+) registers", ln)
            if m and cur:
                info[cur]["regs"] = int(m.group(1))
                sm = re.search(r"(\d+) bytes smem", ln)
                info[cur]["smem"] = int(sm.group(1)) if sm else 0
    return info


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", ln)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            kernels[cur][op] += 1
    names = demangle(list(kernels))
    pinfo = ptxas_info()
    head = subprocess.run(["git", "-C", str(ROOT), "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# SASS summary of {LIB.relative_to(ROOT)} (cuobjdump -sass, sm_100a), tree at {head}; regenerate: python tools/sass_summary.py")
    print("# per kernel: total SASS instructions | ptxas registers, spill stores / loads (bytes), static smem | watched mnemonics")
    for k, c in kernels.items():
        p = pinfo.get(k, {})
        short = re.sub(r"\s+", " ", names.get(k, k))
        short = short.replace("a3::(anonymous namespace)::", "")
        watched = "  ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"\n{short[:150]}")
        print(f"    instructions {c['_total']:6d} | regs {p.get('regs', '?')}, spill st/ld {p.get('spill_st', '?')}/{p.get('spill_ld', '?')} B, stack {p.get('stack', '?')} B, "
              f"smem {p.get('smem', '?')} B | {watched}")
    k1 = [c for k, c in kernels.items() if "k1_strips_kernel" in k]
    print(f"\n# k1_strips_kernel instantiations: {len(k1)}; with UTMALDG: {sum(1 for c in k1 if c['UTMALDG'])}; with SYNCS: {sum(1 for c in k1 if c['SYNCS'])}; "
          f"with local-memory traffic (STL/LDL): {sum(1 for c in k1 if c['STL'] or c['LDL'])}")
    print(f"# kernels using tensor cores (HMMA / UTCMMA / tcgen05): {sum(1 for c in kernels.values() if any(op.startswith(('HMMA', 'UTC', 'IMMA')) for op in c))} "
          "(expected 0: nothing on this path is a contraction)")


if __name__ == "__main__":
    sys.exit(main())
