#!/usr/bin/env python3
"""Extract the marker code tables of the reference into a compact binary blob.

The code tables are *data* (the published ARUCO / AprilTag / ARTag / ARToolKit+ / Chilitags
codebooks), read from the reference's `src/dictionaries.rs:5-19` (constants) and
`src/dictionaries.rs:30-113` (name -> {num_bits, tau, table} map).  Nothing else is taken from
that file.  Output: `aruco3_b200/data/dictionaries.bin`

    char     magic[8]  = "A3DICT01"
    uint32   n_entries
    uint32   n_codes_total
    entry[n_entries]:
        char   name[24]   (NUL padded, upper case, the map key)
        uint8  num_bits
        uint8  tau_table  (0 = "compute as min pairwise distance", dictionaries.rs:124)
        uint16 reserved
        uint32 n_codes
        uint32 first_code (index into the code array)
        uint32 reserved2
    uint64   codes[n_codes_total]   (little endian)

Run here (the reference is not present on the GPU box):  python tools/extract_dictionaries.py
"""
import re
import struct
import sys
from pathlib import Path

REF = Path("/root/reference/src/dictionaries.rs")
OUT = Path(__file__).resolve().parent.parent / "aruco3_b200" / "data" / "dictionaries.bin"


def main() -> int:
    text = REF.read_text()
    tables = {}
    aliases = {}
    for m in re.finditer(r"const\s+(\w+)\s*:\s*&'static\s*\[u64\]\s*=\s*(&\[[^\]]*\]|\w+)\s*;", text):
        name, body = m.group(1), m.group(2)
        if body.startswith("&["):
            tables[name] = [int(tok, 16) for tok in re.findall(r"0x[0-9a-fA-F]+", body)]
        else:
            aliases[name] = body
    entries = []
    # strip the commented-out block so a disabled entry is not picked up
    live = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for m in re.finditer(r'"(\w+)"\s*=>\s*ARDictionary\s*\{\s*num_bits:\s*(\d+),\s*tau:\s*(\d+),\s*code_list:\s*(\w+)\s*,?\s*\}', live):
        key, nb, tau, tbl = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4)
        tbl = aliases.get(tbl, tbl)
        entries.append((key, nb, tau, tbl))
    entries.sort()
    order, first = [], {}
    for _, _, _, tbl in entries:
        if tbl not in first:
            first[tbl] = sum(len(tables[t]) for t in order)
            order.append(tbl)
    codes = [c for t in order for c in tables[t]]
    blob = bytearray(b"A3DICT01")
    blob += struct.pack("<II", len(entries), len(codes))
    for key, nb, tau, tbl in entries:
        blob += struct.pack("<24sBBHIII", key.encode(), nb, tau, 0, len(tables[tbl]), first[tbl], 0)
    blob += struct.pack("<%dQ" % len(codes), *codes)
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_bytes(bytes(blob))
    for key, nb, tau, tbl in entries:
        print(f"{key:18s} bits={nb:2d} tau={tau:2d} n={len(tables[tbl]):5d} first={first[tbl]}")
    print(f"wrote {OUT} ({len(blob)} bytes, {len(codes)} codes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
