"""Host stage of the product (a3_quads_from_mask: border following, RDP, hull, edge test, clockwise, discard —
/root/reference/src/aruco.rs:64-69) against the oracle's restatement of imageproc's find_contours & co.  CPU only."""
import numpy as np
import pytest


def _same(a3, oracle, mask, cfg=None, ocfg=None):
    got = a3.quads_from_mask(mask, cfg)
    want = oracle.candidates_from_mask(mask, ocfg)
    assert got.shape == want.shape and (got == want).all(), f"{got.tolist()} vs {want.tolist()}"
    return got


@pytest.fixture(scope="module")
def a3():
    import aruco3_b200
    return aruco3_b200


@pytest.mark.parametrize("name,frames", [("C1", 4), ("C1n", 2), ("C3", 2), ("C3n", 1), ("C2a", 1), ("C5", 1)])
def test_synthetic_masks(a3, oracle, name, frames):
    from aruco3_b200 import synth
    spec = synth.CONFIGS[name]
    cfg = a3.DetectorConfig(min_corner_separation_factor=spec.min_corner_separation_factor)
    ocfg = oracle.default_config(min_corner_separation_factor=spec.min_corner_separation_factor)
    for f in range(frames):
        img, _ = synth.render_frame(spec, f)
        mask = oracle.adaptive_threshold(oracle.to_luma8(img), 7)
        q = _same(a3, oracle, mask, cfg, ocfg)
        if name in ("C1", "C3"):
            assert len(q) >= 4


@pytest.mark.parametrize("w,h", [(1, 1), (2, 2), (3, 1), (1, 7), (31, 5), (32, 32), (33, 17), (64, 64), (65, 40), (200, 120), (97, 203)])
def test_random_blob_masks(a3, oracle, w, h):
    """Random blobs at several densities: thin structures, touching borders, components on every image edge."""
    rng = np.random.default_rng(w * 1000 + h)
    cfg = a3.DetectorConfig(min_side_length_factor=0.02, min_corner_separation_factor=0.01)
    ocfg = oracle.default_config(min_side_length_factor=0.02, min_corner_separation_factor=0.01)
    for density in (0.05, 0.3, 0.5, 0.7, 0.95):
        for _ in range(6):
            mask = ((rng.random((h, w)) < density) * 255).astype(np.uint8)
            _same(a3, oracle, mask, cfg, ocfg)
    # blocky content produces real quads
    for _ in range(6):
        mask = np.zeros((h, w), np.uint8)
        for _ in range(8):
            x0, y0 = rng.integers(0, max(w, 1)), rng.integers(0, max(h, 1))
            mask[y0:y0 + rng.integers(1, 40), x0:x0 + rng.integers(1, 40)] = 255
        for _ in range(4):
            x0, y0 = rng.integers(0, max(w, 1)), rng.integers(0, max(h, 1))
            mask[y0:y0 + rng.integers(1, 12), x0:x0 + rng.integers(1, 12)] = 0
        _same(a3, oracle, mask, cfg, ocfg)


def test_degenerate_masks(a3, oracle):
    for mask in (np.zeros((40, 50), np.uint8), np.full((40, 50), 255, np.uint8), np.eye(40, 50, dtype=np.uint8) * 255):
        _same(a3, oracle, mask)
    m = np.zeros((60, 80), np.uint8)
    m[10:50, 10:70] = 255          # one rectangle: outer border only
    q = _same(a3, oracle, m, a3.DetectorConfig(min_side_length_factor=0.05), oracle.default_config(min_side_length_factor=0.05))
    assert q.tolist() == [[10, 10, 69, 10, 69, 49, 10, 49]]   # hull order: top-left first, screen-clockwise
    m[20:40, 20:60] = 0            # ring: outer + hole border
    _same(a3, oracle, m, a3.DetectorConfig(min_side_length_factor=0.05), oracle.default_config(min_side_length_factor=0.05))
    m[:, 0] = 255                  # column 0 never starts an outer border
    m[:, -1] = 255
    _same(a3, oracle, m)


def test_contour_counts_match(a3, oracle):
    """The stage's counters (parity probes): same number of borders and border points as find_contours."""
    import ctypes as C
    from aruco3_b200 import _ffi, synth
    img, _ = synth.render_frame(synth.CONFIGS["C1n"], 3)
    mask = oracle.adaptive_threshold(oracle.to_luma8(img), 7)
    contours, _ = oracle.find_contours(mask)
    cfg = a3.DetectorConfig().to_c()
    st = _ffi.A3Stats()
    n = C.c_uint32()
    quads = np.zeros((256, 8), np.uint32)
    _ffi.check(_ffi.lib().a3_quads_from_mask(C.byref(cfg), mask.ctypes.data, 640, 480, quads.ctypes.data, 256, C.byref(n), C.byref(st)))
    assert st.n_contours == len(contours) and st.n_contour_points == sum(len(c) for c in contours)
