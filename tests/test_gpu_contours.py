"""Kernel K3 (find_contours + quad filters on the device, csrc/k3_contours.cu) against the oracle and the host stage.

K3 decides every border start from the image alone (tools/contour_parallel_proto.py); frames where that formulation
could differ from the sequential reference are flagged by the kernel and redone by the host stage, so the test checks
(a) unflagged frames: identical quads, contour count and point count, (b) the flag is raised exactly for the documented
reason often enough to matter on frame-touching content and never on the benchmark workloads, (c) the end-to-end path
gives identical results in both contour modes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def a3():
    import aruco3_b200
    return aruco3_b200


def _oracle_frame(oracle, mask, ocfg):
    contours, _ = oracle.find_contours(mask)
    return oracle.candidates_from_mask(mask, ocfg), len(contours), sum(len(c) for c in contours)


def _check(a3, oracle, masks, cfg, ocfg, allow_flags=True):
    with a3.Detector(cfg) as d:
        quads, flags, contours, points = d.quads_from_masks_device(masks)
    nflag = 0
    for f in range(len(masks)):
        if flags[f]:
            nflag += 1
            assert allow_flags, f"frame {f} was flagged ({flags[f]})"
            continue
        want, nc, npnt = _oracle_frame(oracle, masks[f], ocfg)
        assert quads[f].tolist() == want.tolist(), f"frame {f}: quads"
        assert (int(contours[f]), int(points[f])) == (nc, npnt), f"frame {f}: contour statistics"
    return nflag


@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (31, 7), (32, 32), (33, 17), (64, 48), (97, 131), (200, 120)])
def test_random_masks(a3, oracle, w, h):
    rng = np.random.default_rng(w * 977 + h)
    cfg = a3.DetectorConfig(min_side_length_factor=0.02, min_corner_separation_factor=0.01)
    ocfg = oracle.default_config(min_side_length_factor=0.02, min_corner_separation_factor=0.01)
    masks = []
    for density in (0.05, 0.3, 0.5, 0.7, 0.95):
        for k in range(8):
            m = ((rng.random((h, w)) < density) * 255).astype(np.uint8)
            if k % 2 == 0 and w > 2 and h > 2:  # nothing touches the frame: K3 may not flag these
                m[0] = m[-1] = 0
                m[:, 0] = m[:, -1] = 0
            masks.append(m)
    for _ in range(10):  # blocky content produces real quads
        m = np.zeros((h, w), np.uint8)
        for _ in range(8):
            x0, y0 = rng.integers(0, w), rng.integers(0, h)
            m[y0:y0 + rng.integers(1, 40), x0:x0 + rng.integers(1, 40)] = 255
        for _ in range(4):
            x0, y0 = rng.integers(0, w), rng.integers(0, h)
            m[y0:y0 + rng.integers(1, 12), x0:x0 + rng.integers(1, 12)] = 0
        masks.append(m)
    masks = np.stack(masks)
    _check(a3, oracle, masks, cfg, ocfg)
    inner = np.stack([m for i, m in enumerate(masks[:40]) if i % 2 == 0]) if w > 2 and h > 2 else None
    if inner is not None:
        assert _check(a3, oracle, inner, cfg, ocfg, allow_flags=False) == 0


@pytest.mark.parametrize("name,frames", [("C1", 4), ("C1n", 2), ("C3", 2), ("C3n", 1), ("C2a", 1), ("C5", 1)])
def test_benchmark_masks_are_not_flagged(a3, oracle, name, frames):
    from aruco3_b200 import synth
    spec = synth.CONFIGS[name]
    cfg = a3.DetectorConfig(min_corner_separation_factor=spec.min_corner_separation_factor)
    ocfg = oracle.default_config(min_corner_separation_factor=spec.min_corner_separation_factor)
    masks = np.stack([oracle.adaptive_threshold(oracle.to_luma8(synth.render_frame(spec, f)[0]), 7) for f in range(frames)])
    nflag = _check(a3, oracle, masks, cfg, ocfg)
    if name in ("C1", "C3", "C5"):
        assert nflag == 0


def test_flag_for_barred_start(a3, oracle):
    """A white area whose top row spans the frame (raster-first pixel in column 0) with a dark blob on the left edge: the
    reference starts that border from a west crack below its top — the documented reason for a host redo."""
    m = np.full((60, 80), 255, np.uint8)
    m[20:30, 0:12] = 0
    with a3.Detector() as d:
        _, flags, _, _ = d.quads_from_masks_device(m)
    assert flags[0] & 1


@pytest.mark.parametrize("name,frames", [("C1", 5), ("C1n", 3), ("C3", 2), ("C3n", 1), ("C2a", 1), ("C5", 1)])
def test_detect_is_identical_in_both_modes(a3, name, frames):
    from aruco3_b200 import synth
    spec = synth.CONFIGS[name]
    imgs, _ = synth.render_batch(spec, frames)
    cfg = a3.DetectorConfig(min_corner_separation_factor=spec.min_corner_separation_factor)
    res = {}
    for mode in ("host", "device"):
        with a3.Detector(cfg, spec.dictionary, contours=mode) as d:
            dets = d.detect_batch(imgs, full=True)
            res[mode] = ([(x.candidates, [(m.id, m.rotation, m.hamming_distance, m.code, m.corners, m.candidate) for m in x.markers]) for x in dets],
                         {k: d.last_stats[k] for k in ("n_contours", "n_contour_points", "n_candidates_before_discard", "n_candidates", "n_markers")})
    assert res["host"] == res["device"]


def test_flagged_frames_are_redone_by_the_host_stage(a3, oracle):
    """End to end on frames that K3 flags: a grey-level frame whose mask has the barred-start shape."""
    img = np.full((3, 96, 128), 200, np.uint8)
    img[:, 30:50, 0:20] = 20          # dark blob on the left edge
    img[1, 10:40, 60:100] = 30        # plus a dark rectangle
    with a3.Detector(contours="device") as d:
        dev = d.detect_batch(img, full=True)
        assert d.last_stats["host_fallback_frames"] >= 1
    with a3.Detector(contours="host") as d:
        host = d.detect_batch(img, full=True)
    for f in range(3):
        assert dev[f].candidates == host[f].candidates
        assert dev[f].candidates == [[tuple(q[2 * k:2 * k + 2]) for k in range(4)] for q in oracle.detect(img[f]).candidates.tolist()]


def test_many_small_masks_in_one_batch(a3, oracle):
    """Stress: 1500 random 48x40 masks of every density in ONE K3 call (thin structures, frame contact, nested holes);
    every unflagged frame must equal the sequential oracle, and the flag rate must stay what the barred-start rule explains
    (only frames with foreground in column 0 can be flagged)."""
    rng = np.random.default_rng(2024)
    n, h, w = 1500, 40, 48
    dens = rng.choice([0.1, 0.3, 0.45, 0.55, 0.7, 0.9, 0.98], size=n)
    masks = ((rng.random((n, h, w)) < dens[:, None, None]) * 255).astype(np.uint8)
    masks[::5, :, 0] = 0  # a fifth of the frames have an empty first column: those can never be flagged
    cfg = a3.DetectorConfig(min_side_length_factor=0.05, min_corner_separation_factor=0.02)
    ocfg = oracle.default_config(min_side_length_factor=0.05, min_corner_separation_factor=0.02)
    with a3.Detector(cfg) as d:
        quads, flags, contours, points = d.quads_from_masks_device(masks, quad_capacity=256)
    assert not flags[::5].any()
    assert flags.astype(bool).mean() < 0.5
    for f in range(n):
        if flags[f]:
            assert flags[f] == 1
            continue
        want, nc, npnt = _oracle_frame(oracle, masks[f], ocfg)
        assert quads[f].tolist() == want.tolist() and (int(contours[f]), int(points[f])) == (nc, npnt), f"frame {f}"


def test_wide_frames_take_the_64_bit_distance_path(a3, oracle):
    """Coordinates of 2^14 and more: k3_rdp's point-to-chord numerators no longer fit 32 bits (the fast loop is only taken
    below that), and long thin quads make the products large."""
    rng = np.random.default_rng(77)
    h, w = 40, 40000
    cfg = a3.DetectorConfig(min_side_length_factor=0.1, min_corner_separation_factor=0.05)
    ocfg = oracle.default_config(min_side_length_factor=0.1, min_corner_separation_factor=0.05)
    masks = []
    for k in range(4):
        m = np.zeros((h, w), np.uint8)
        for _ in range(12):  # long thin rectangles and slanted bars far from the origin
            x0, y0 = int(rng.integers(1, w - 9000)), int(rng.integers(1, h - 12))
            ln, th = int(rng.integers(50, 8000)), int(rng.integers(2, 10))
            m[y0:y0 + th, x0:x0 + ln] = 255
            if k % 2:
                for r in range(th):  # shear: one pixel per row
                    m[y0 + r, x0:x0 + r] = 0
        m[:, 0] = m[:, -1] = 0
        m[0] = m[-1] = 0
        noise = (rng.random((h, 2000)) < 0.5) * 255
        m[:, 30000:32000] = noise.astype(np.uint8)
        m[0, 30000:32000] = m[-1, 30000:32000] = 0
        masks.append(m)
    nflag = _check(a3, oracle, np.stack(masks), cfg, ocfg)
    assert nflag == 0


def test_more_than_4096_frames_take_the_radix_sort(a3, oracle):
    """k3_order ranks the long borders frame by frame for calls of up to 4096 frames; larger calls (and frames with more than
    1024 long borders) order them with the radix sort.  4500 small masks in one call, every unflagged one equal to the oracle."""
    rng = np.random.default_rng(4500)
    n, h, w = 4500, 24, 32
    dens = rng.choice([0.2, 0.5, 0.8], size=n)
    masks = ((rng.random((n, h, w)) < dens[:, None, None]) * 255).astype(np.uint8)
    masks[:, :, 0] = 0  # no foreground in column 0: nothing can be flagged
    cfg = a3.DetectorConfig(min_side_length_factor=0.05, min_corner_separation_factor=0.02)
    ocfg = oracle.default_config(min_side_length_factor=0.05, min_corner_separation_factor=0.02)
    with a3.Detector(cfg) as d:
        quads, flags, contours, points = d.quads_from_masks_device(masks, quad_capacity=64)
    assert not flags.any()
    for f in range(0, n, 3):
        want, nc, npnt = _oracle_frame(oracle, masks[f], ocfg)
        assert quads[f].tolist() == want.tolist() and (int(contours[f]), int(points[f])) == (nc, npnt), f"frame {f}"


@pytest.mark.parametrize("route", ["relays", "relays-short-jumps", "pairs"])
def test_both_walk_routes(a3, oracle, route, monkeypatch):
    """Long borders are walked from relay cracks (k3_segments / k3_cycles; calls of up to 16 frames by default) or by lane pairs from
    their start candidate (k3_walkers; batches).  A3_K3_RELAY_MAX_FRAMES, read at every call, forces either route: both must give the
    oracle's quads and contour statistics on random, blocky, striped and noise content, single frames and batches alike."""
    monkeypatch.setenv("A3_K3_RELAY_MAX_FRAMES", "0" if route == "pairs" else "1000000")
    if route == "relays-short-jumps":  # k3_jumps sums 32 segments per jump; 2 makes the cycles of these small masks take many jumps
        monkeypatch.setenv("A3_K3_JUMP", "2")
    rng = np.random.default_rng(4242)
    cfg = a3.DetectorConfig(min_side_length_factor=0.02, min_corner_separation_factor=0.01)
    ocfg = oracle.default_config(min_side_length_factor=0.02, min_corner_separation_factor=0.01)
    for (w, h) in [(33, 17), (97, 131), (200, 120), (640, 70)]:
        masks = []
        for density in (0.1, 0.5, 0.9):
            m = ((rng.random((h, w)) < density) * 255).astype(np.uint8)
            m[0] = m[-1] = 0
            m[:, 0] = m[:, -1] = 0
            masks.append(m)
        for _ in range(6):  # blocks, holes, long horizontal and vertical bars: borders that cross many relay rows, or none
            m = np.zeros((h, w), np.uint8)
            for _ in range(6):
                x0, y0 = rng.integers(1, w - 1), rng.integers(1, h - 1)
                m[y0:min(h - 1, y0 + rng.integers(1, 60)), x0:min(w - 1, x0 + rng.integers(1, 90))] = 255
            for _ in range(3):
                x0, y0 = rng.integers(1, w - 1), rng.integers(1, h - 1)
                m[y0:min(h - 1, y0 + rng.integers(1, 14)), x0:min(w - 1, x0 + rng.integers(1, 14))] = 0
            m[h // 2, 1:w - 1] = 255          # a one-pixel line across the frame
            m[1:h - 1, w // 3] = 255
            masks.append(m)
        masks = np.stack(masks)
        assert _check(a3, oracle, masks, cfg, ocfg, allow_flags=False) == 0          # one call with all frames
        for f in (0, 4, len(masks) - 1):                                              # and frame by frame
            assert _check(a3, oracle, masks[f:f + 1], cfg, ocfg, allow_flags=False) == 0


def test_relay_route_on_detection_frames(a3, oracle, monkeypatch):
    """The whole path on rendered frames with relays forced on for a batch and off for single frames: same markers as the oracle."""
    from aruco3_b200 import synth
    frames, _ = synth.render_batch("C1", 6)
    want = [[(m["id"], m["rotation"], m["hamming_distance"], m["code"], m["corners"]) for m in oracle.detect(f, "ARUCO").markers] for f in frames]
    for env, batches in (("1000000", [frames]), ("0", [frames[i:i + 1] for i in range(3)])):
        monkeypatch.setenv("A3_K3_RELAY_MAX_FRAMES", env)
        with a3.Detector() as d:
            k = 0
            for b in batches:
                for _ in range(2):  # second call: the one-shot route with speculated list sizes
                    got = d.detect_batch(b)
                assert [[(m.id, m.rotation, m.hamming_distance, m.code, [v for c in m.corners for v in c]) for m in x.markers] for x in got] == want[k:k + len(b)]
                k += len(b)


def test_relay_list_overflow_goes_to_the_host_stage(a3, oracle, monkeypatch):
    """A relay list too small for the call (A3_K3_RELAY_CAP, a test hook) is an overflow like that of any other K3 list: every frame
    of the call is flagged and redone by the host stage, and the answer stays the oracle's."""
    from aruco3_b200 import synth
    frames, _ = synth.render_batch("C1", 2)
    want = [[(m["id"], m["rotation"], m["hamming_distance"], m["code"], m["corners"]) for m in oracle.detect(f, "ARUCO").markers] for f in frames]
    monkeypatch.setenv("A3_K3_RELAY_CAP", "8")
    with a3.Detector() as d:
        for _ in range(2):
            got = d.detect_batch(frames)
            assert [[(m.id, m.rotation, m.hamming_distance, m.code, [v for c in m.corners for v in c]) for m in x.markers] for x in got] == want
        masks = np.stack([d.gray_threshold(f[None])[1][0] for f in frames])
        _, flags, _, _ = d.quads_from_masks_device(masks)
        assert all(int(f) & 8 for f in flags)  # bit 3: a work list overflowed
    monkeypatch.delenv("A3_K3_RELAY_CAP")
    with a3.Detector() as d:
        _, flags, _, _ = d.quads_from_masks_device(masks)
        assert not any(int(f) for f in flags)
