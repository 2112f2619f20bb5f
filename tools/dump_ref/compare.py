#!/usr/bin/env python3
"""Diff the dumps of the real reference (tools/dump_ref/src/main.rs) against the oracle's golden vectors
(tests/golden/detect_golden.json): grey bytes, candidates, markers (id, hamming distance, observed code, corners) and
which projections failed.  All equal => the oracle (and with it the CUDA path, which tests hold bit-exact to the oracle)
is pinned to the reference on these frames.  usage: python tools/dump_ref/compare.py DUMPDIR"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]


def main(dump_dir):
    golden = json.loads((ROOT / "tests" / "golden" / "detect_golden.json").read_text())
    bad = 0
    for case in golden["cases"]:
        stems = sorted(Path(dump_dir).glob(f"{case['config']}_{case['frame']}_*.json"))
        if not stems:
            print("missing dump for", case["config"], case["frame"])
            bad += 1
            continue
        ref = json.loads(stems[0].read_text())
        grey = stems[0].with_suffix(".grey").read_bytes()
        checks = {
            "grey": hashlib.sha256(grey).hexdigest() == case["grey_sha256"],
            "candidates": ref["candidates"] == case["candidates"],
            # golden marker rows: [candidate, id, rotation, hamming, code, corners...]; the reference exposes no candidate / rotation
            "markers": ref["markers"] == [[m[1], m[3], m[4]] + m[5:] for m in case["markers"]],
        }
        ok = all(checks.values())
        bad += not ok
        print(case["config"], case["frame"], "OK" if ok else f"DIFFERS: {[k for k, v in checks.items() if not v]}")
    print("all equal: the oracle is pinned on these frames" if not bad else f"{bad} case(s) differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
