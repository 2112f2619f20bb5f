"""Golden vectors (tests/golden/detect_golden.json, written by tools/make_golden.py).

The reference's own tests pin nothing on the pixel path and the Rust crate cannot be built here, so the vectors are
(a) the oracle's outputs, frozen, and (b) what two independent sources say about the same frames: the renderer's
ground truth and OpenCV's ArUco detector (DICT_ARUCO_ORIGINAL is the reference's `ARUCO` table).  The CPU test pins
the oracle; the GPU test pins the CUDA path to the same file without running the oracle at all."""
import hashlib
import json
from collections import Counter
from pathlib import Path

import numpy as np
import pytest

GOLDEN = json.loads((Path(__file__).resolve().parent / "golden" / "detect_golden.json").read_text())


def _render(case):
    from aruco3_b200 import synth
    spec = synth.CONFIGS[case["config"]]
    img, _ = synth.render_frame(spec, case["frame"])
    assert hashlib.sha256(img.tobytes()).hexdigest() == case["rgb_sha256"], "the synthetic frame generator changed"
    return spec, img


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: f"{c['config']}-{c['frame']}")
def test_oracle_reproduces_golden(oracle, case):
    spec, img = _render(case)
    r = oracle.detect(img, spec.dictionary, oracle.default_config(min_corner_separation_factor=spec.min_corner_separation_factor))
    assert hashlib.sha256(r.grey.tobytes()).hexdigest() == case["grey_sha256"]
    assert hashlib.sha256(r.mask.tobytes()).hexdigest() == case["mask_sha256"]
    assert (r.stats["n_contours"], r.stats["n_contour_points"]) == (case["n_contours"], case["n_contour_points"])
    assert r.candidates.tolist() == case["candidates"]
    assert r.otsu.tolist() == case["otsu"] and r.has_codes.tolist() == case["has_codes"]
    got = [[m["candidate"], m["id"], m["rotation"], m["hamming_distance"], m["code"]] + m["corners"] for m in r.markers]
    assert got == case["markers"]


@pytest.mark.parametrize("case", [c for c in GOLDEN["cases"] if "cv2_ids" in c], ids=lambda c: f"{c['config']}-{c['frame']}")
def test_golden_agrees_with_independent_sources(case):
    """ids: never an id that neither the renderer nor OpenCV knows about; at most one marker per frame missed."""
    ours = Counter(m[1] for m in case["markers"])
    truth, cv = Counter(case["truth_ids"]), Counter(case["cv2_ids"])
    assert cv == truth, "OpenCV reads the rendered frame differently from the ground truth"
    assert not (ours - truth), f"ids not in the ground truth: {sorted((ours - truth).elements())}"
    assert sum((truth - ours).values()) <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: f"{c['config']}-{c['frame']}")
def test_cuda_path_reproduces_golden(case):
    import aruco3_b200 as a3
    spec, img = _render(case)
    with a3.Detector(a3.DetectorConfig(min_corner_separation_factor=spec.min_corner_separation_factor), spec.dictionary) as d:
        det = d.detect_batch(img[None], full=True, want_mask=True)[0]
    assert hashlib.sha256(np.ascontiguousarray(det.grey).tobytes()).hexdigest() == case["grey_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(det.mask).tobytes()).hexdigest() == case["mask_sha256"]
    assert [list(sum(c, ())) for c in det.candidates] == case["candidates"]
    assert [dc["otsu"] for dc in det.decodes] == case["otsu"]
    assert [int(dc["has_codes"]) for dc in det.decodes] == case["has_codes"]
    got = [[m.candidate, m.id, m.rotation, m.hamming_distance, m.code] + [v for c in m.corners for v in c] for m in det.markers]
    assert got == case["markers"]
    assert (d.last_stats["n_contours"], d.last_stats["n_contour_points"]) == (case["n_contours"], case["n_contour_points"])
