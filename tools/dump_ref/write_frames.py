#!/usr/bin/env python3
"""Write the seeded frames of tests/golden/detect_golden.json as raw RGB files for tools/dump_ref (the real reference).
usage: python tools/dump_ref/write_frames.py OUTDIR"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from aruco3_b200 import synth  # noqa: E402


def main(out):
    out = Path(out)
    out.mkdir(parents=True, exist_ok=True)
    golden = json.loads((ROOT / "tests" / "golden" / "detect_golden.json").read_text())
    for case in golden["cases"]:
        spec = synth.CONFIGS[case["config"]]
        img, _ = synth.render_frame(spec, case["frame"])
        assert hashlib.sha256(img.tobytes()).hexdigest() == case["rgb_sha256"], "synth.py no longer reproduces the golden frames"
        h, w = img.shape[:2]
        name = f"{case['config']}_{case['frame']}_{spec.dictionary}_{spec.min_corner_separation_factor}_{w}x{h}.rgb"
        (out / name).write_bytes(img.tobytes())
        print(name)


if __name__ == "__main__":
    main(sys.argv[1])
