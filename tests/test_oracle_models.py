"""Cross-checks of the oracle against independent models where the semantics provably coincide:
pure-integer stages against numpy restatements written from the spec, the border set against OpenCV's Suzuki-Abe,
the homography against OpenCV's solver, the triangle resize against Pillow (soft), and the whole path against the
renderer's ground truth including the rotation / corner-order convention."""
import ctypes as C

import numpy as np
import pytest


def test_luma_matches_integer_formula(oracle):
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    c = rgb.astype(np.uint32)
    want = ((2126 * c[..., 0] + 7152 * c[..., 1] + 722 * c[..., 2]) // 10000).astype(np.uint8)
    assert np.array_equal(oracle.to_luma8(rgb), want)
    rgba = np.dstack([rgb, rng.integers(0, 256, size=(37, 53, 1), dtype=np.uint8)])
    assert np.array_equal(oracle.to_luma8(rgba), want)            # alpha ignored
    assert np.array_equal(oracle.to_luma8(want), want)            # Luma8 passes through
    px = np.array([[[255, 255, 255], [0, 255, 0], [255, 0, 0], [0, 0, 255]]], np.uint8)
    assert oracle.to_luma8(px).tolist() == [[255, 182, 54, 18]]
    # camera byte order (examples/webcam_kamera.rs:38-52 swizzle, then into_luma8)
    assert np.array_equal(oracle.to_luma8(np.ascontiguousarray(rgb[..., ::-1]), order="bgr"), want)
    bgra = np.dstack([rgb[..., ::-1], rgba[..., 3:]])
    assert np.array_equal(oracle.to_luma8(np.ascontiguousarray(bgra), order="bgr"), want)


@pytest.mark.parametrize("w,h,r", [(40, 30, 7), (9, 5, 7), (15, 15, 7), (64, 3, 3), (1, 1, 7), (33, 47, 1), (50, 20, 12)])
def test_adaptive_threshold_matches_window_model(oracle, w, h, r):
    """out = 255 iff pix >= floor(sum(window) / area(window)), window clipped to the image."""
    rng = np.random.default_rng(w * h + r)
    for kind in range(3):
        g = rng.integers(0, 256, size=(h, w), dtype=np.uint8) if kind == 0 else \
            (np.full((h, w), 128, np.uint8) if kind == 1 else (120 + rng.integers(0, 3, size=(h, w))).astype(np.uint8))
        want = np.zeros_like(g)
        for y in range(h):
            for x in range(w):
                win = g[max(0, y - r):min(h, y + r + 1), max(0, x - r):min(w, x + r + 1)].astype(np.uint32)
                want[y, x] = 255 if g[y, x] >= win.sum() // win.size else 0
        assert np.array_equal(oracle.adaptive_threshold(g, r), want)
    assert (oracle.adaptive_threshold(np.full((20, 20), 77, np.uint8), 7) == 255).all()  # flat regions are white


def test_otsu_matches_independent_model(oracle):
    rng = np.random.default_rng(5)
    L = oracle.lib()
    for _ in range(30):
        img = np.clip(rng.normal(rng.integers(40, 120), 15, (49, 49)) * (rng.random((49, 49)) < 0.5) +
                      rng.normal(rng.integers(140, 230), 20, (49, 49)) * (rng.random((49, 49)) < 0.5), 0, 255).astype(np.uint8)
        hist = np.bincount(img.ravel(), minlength=256).astype(np.float64)
        t = np.arange(256, dtype=np.float64)
        bw, bs = np.cumsum(hist), np.cumsum(t * hist)
        fw, fs = bw[-1] - bw, bs[-1] - bs
        with np.errstate(divide="ignore", invalid="ignore"):
            var = np.where((bw > 0) & (fw > 0), bw * fw * (bs / bw - fs / fw) ** 2, -1.0)
        best, level = 0.0, 0
        for i in range(256):
            if var[i] > best:
                best, level = var[i], i
        assert L.a3ref_otsu_level(np.ascontiguousarray(img).ctypes.data, 49, 49) == level
    assert L.a3ref_otsu_level(np.full((49, 49), 9, np.uint8).ctypes.data, 49, 49) == 0  # uniform patch -> 0


def test_border_set_matches_opencv(oracle):
    """imageproc's find_contours and OpenCV's findContours are both Suzuki-Abe: away from the image frame they must
    produce the same borders (as point sets) — outer and hole borders alike."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for trial in range(12):
        m = np.zeros((90, 120), np.uint8)
        for _ in range(10):
            x0, y0 = rng.integers(4, 100), rng.integers(4, 70)
            m[y0:y0 + rng.integers(2, 30), x0:x0 + rng.integers(2, 30)] = 255
        for _ in range(5):
            x0, y0 = rng.integers(4, 100), rng.integers(4, 70)
            m[y0:y0 + rng.integers(1, 10), x0:x0 + rng.integers(1, 10)] = 0
        m[rng.random(m.shape) < 0.02] = 255
        m[:2] = m[-2:] = 0
        m[:, :2] = m[:, -2:] = 0
        ours, outer = oracle.find_contours(m)
        theirs, _ = cv2.findContours(m, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        a = sorted(tuple(sorted(set(map(tuple, c.tolist())))) for c in ours)
        b = sorted(tuple(sorted(set(map(tuple, c.reshape(-1, 2).tolist())))) for c in theirs)
        assert a == b, f"trial {trial}: {len(a)} vs {len(b)} borders"
        for c, is_outer in zip(ours, outer):  # an outer border starts at its raster-first point (discovery order)
            pts = c.tolist()
            if is_outer:
                assert pts[0] == min(pts, key=lambda p: (p[1], p[0]))


def test_projection_matches_opencv(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    L = oracle.lib()
    to = np.array([0, 0, 49, 0, 49, 49, 0, 49], np.float32)
    for _ in range(20):
        c = rng.uniform(100, 500, 2)
        quad = (c + rng.uniform(-1, 1, (4, 2)) * 8 + np.array([[-40, -40], [40, -40], [40, 40], [-40, 40]])).astype(np.float32)
        fwd, inv = np.zeros(9, np.float32), np.zeros(9, np.float32)
        cls = C.c_int()
        src = np.ascontiguousarray(quad.ravel())
        assert L.a3ref_projection_from_control_points(src.ctypes.data, to.ctypes.data, fwd.ctypes.data, inv.ctypes.data, C.byref(cls))
        h = cv2.getPerspectiveTransform(quad, to.reshape(4, 2))
        assert np.allclose(fwd.reshape(3, 3), h / h[2, 2], rtol=2e-4, atol=2e-4)
        assert np.allclose(inv.reshape(3, 3) @ fwd.reshape(3, 3) / (inv.reshape(3, 3) @ fwd.reshape(3, 3))[2, 2], np.eye(3), atol=1e-3)


def test_triangle_resize_close_to_pillow(oracle):
    """image::imageops::resize(Triangle) and Pillow's BILINEAR reduce use the same support-scaled triangle filter
    (Pillow in fixed point): within +-1 grey level for 49 -> 6 / 7 / 8 / 10 (soft check)."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(4)
    L = oracle.lib()
    for ms in (6, 7, 8, 10):
        for _ in range(5):
            src = ((rng.random((49, 49)) < 0.5) * 255).astype(np.uint8)
            out = np.zeros((ms, ms), np.uint8)
            L.a3ref_resize_triangle(src.ctypes.data, 49, 49, ms, ms, out.ctypes.data)
            ref = np.asarray(Image.fromarray(src).resize((ms, ms), Image.BILINEAR))
            assert np.abs(out.astype(int) - ref.astype(int)).max() <= 1


@pytest.mark.parametrize("quarter_turns", [0, 1, 2, 3])
def test_rotation_and_corner_convention(oracle, quarter_turns):
    """corners[0] is the marker's own top-left, winding TL, TR, BR, BL (screen-clockwise); a marker turned by
    k quarter-turns clockwise on screen decodes with rotation index k."""
    from aruco3_b200 import dictionaries, synth
    table = dictionaries.table("ARUCO")
    img = np.empty((480, 640, 3), np.uint8)
    img[:] = (204, 200, 192)
    base = np.array([[-50.0, -50.0], [50.0, -50.0], [50.0, 50.0], [-50.0, 50.0]])  # TL TR BR BL
    ang = np.pi / 2 * quarter_turns
    rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
    quad = base @ rot.T + np.array([320.0, 240.0])
    synth.draw_marker(img, synth.marker_cells(table, 77), quad, (250, 248, 240), (24, 28, 36))
    r = oracle.detect(img, "ARUCO")
    assert [m["id"] for m in r.markers] == [77] and r.markers[0]["hamming_distance"] == 0
    c = np.array(r.markers[0]["corners"], float).reshape(4, 2)
    assert np.abs(c - quad).max() <= 2.5, (c.tolist(), quad.tolist())
    assert r.markers[0]["rotation"] == quarter_turns


def test_ground_truth_on_clean_frames(oracle):
    """The oracle on BASELINE.json's 640x480 config: ~92 % of the rendered markers come back, none with a wrong id."""
    from collections import Counter
    from aruco3_b200 import synth
    found = wanted = 0
    for f in range(12):
        img, truth = synth.render_frame(synth.CONFIGS["C1"], f)
        have, want = Counter(m["id"] for m in oracle.detect(img, "ARUCO").markers), Counter(t.id for t in truth)
        assert not (have - want)
        found += sum((have & want).values())
        wanted += sum(want.values())
    assert found >= 0.85 * wanted


def test_wide_formats_into_luma8_model(oracle):
    """oracle/a3ref.c's into_luma8 for LumaA8 / Luma16 / LumaA16 / Rgb16 / Rgba16 against an independent numpy model of the
    image-crate rules it restates (recalled, unpinned upstream), and the crate's own claim about its
    u16 -> u8 rule: (c + 128) / 257 == round(c * 255 / 65535) for every c."""
    c = np.arange(65536, dtype=np.int64)
    assert np.array_equal((c + 128) // 257, np.floor(c * 255 / 65535 + 0.5).astype(np.int64))
    rng = np.random.default_rng(3)
    h, w = 37, 53
    rgba16 = rng.integers(0, 65536, size=(h, w, 4), dtype=np.uint16)
    rgba16[0, :8] = [[0, 0, 0, 0], [65535] * 4, [65535, 0, 0, 9], [0, 65535, 0, 9], [0, 0, 65535, 9], [127, 128, 129, 0], [385, 386, 384, 0], [32767, 32768, 32769, 1]]
    x = rgba16.astype(np.uint64)
    l16 = (2126 * x[..., 0] + 7152 * x[..., 1] + 722 * x[..., 2]) // 10000
    want_rgb = ((l16 + 128) // 257).astype(np.uint8)
    assert np.array_equal(oracle.to_luma8(rgba16), want_rgb)
    assert np.array_equal(oracle.to_luma8(np.ascontiguousarray(rgba16[..., :3])), want_rgb)
    want_l = ((x[..., 0] + 128) // 257).astype(np.uint8)
    assert np.array_equal(oracle.to_luma8(np.ascontiguousarray(rgba16[..., 0])), want_l)
    assert np.array_equal(oracle.to_luma8(np.ascontiguousarray(rgba16[..., :2])), want_l)
    la8 = rng.integers(0, 256, size=(h, w, 2), dtype=np.uint8)
    assert np.array_equal(oracle.to_luma8(la8), la8[..., 0])
    # a detect() on a 16-bit image equals detect() on its Luma8 conversion
    from aruco3_b200 import synth
    rgb, _ = synth.render_frame(synth.CONFIGS["C1"], 1)
    rgb16 = rgb.astype(np.uint16) * 257
    a, b = oracle.detect(rgb16, "ARUCO"), oracle.detect(oracle.to_luma8(rgb16), "ARUCO")
    assert np.array_equal(a.grey, b.grey) and a.markers == b.markers and len(a.markers) >= 3
