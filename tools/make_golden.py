#!/usr/bin/env python3
"""Write tests/golden/detect_golden.json: per-frame results of the CPU oracle on seeded synthetic frames, plus the ids
an independent detector (OpenCV's cv2.aruco, DICT_ARUCO_ORIGINAL == the reference's `ARUCO` table) finds on the same
frames.  The reference itself (Rust) cannot run in this image, so these are regression + cross-check vectors, not
upstream outputs: see oracle/a3ref.h "PARITY STATUS".  Run here:  python tools/make_golden.py"""
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

from aruco3_b200 import synth  # noqa: E402
from oracle import a3ref_py  # noqa: E402

CASES = [("C1", 6), ("C1n", 2), ("C3", 1), ("C5", 1), ("C2a", 1)]


def main():
    try:
        import cv2
        cvdet = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_ARUCO_ORIGINAL), cv2.aruco.DetectorParameters())
    except Exception:
        cvdet = None
    out = {"generator": "tools/make_golden.py", "oracle": "oracle/a3ref.c", "cases": []}
    for name, n in CASES:
        spec = synth.CONFIGS[name]
        cfg = a3ref_py.default_config(min_corner_separation_factor=spec.min_corner_separation_factor)
        for f in range(n):
            img, truth = synth.render_frame(spec, f)
            r = a3ref_py.detect(img, spec.dictionary, cfg)
            case = {"config": name, "frame": f, "rgb_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
                    "grey_sha256": hashlib.sha256(r.grey.tobytes()).hexdigest(),
                    "mask_sha256": hashlib.sha256(r.mask.tobytes()).hexdigest(),
                    "n_contours": int(r.stats["n_contours"]), "n_contour_points": int(r.stats["n_contour_points"]),
                    "candidates": r.candidates.tolist(), "otsu": r.otsu.tolist(), "has_codes": r.has_codes.tolist(),
                    "markers": [[m["candidate"], m["id"], m["rotation"], m["hamming_distance"], m["code"]] + m["corners"] for m in r.markers],
                    "truth_ids": sorted(int(t.id) for t in truth)}
            if cvdet is not None and spec.dictionary == "ARUCO" and not spec.pure_noise:
                _, ids, _ = cvdet.detectMarkers(img)
                case["cv2_ids"] = sorted(int(i) for i in (ids.ravel() if ids is not None else []))
            out["cases"].append(case)
            print(name, f, "markers", len(r.markers), "truth", len(truth), "cv2", len(case.get("cv2_ids", [])))
    path = ROOT / "tests" / "golden" / "detect_golden.json"
    path.write_text(json.dumps(out, separators=(",", ":")))
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
