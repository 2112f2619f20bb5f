"""The one-shot route of a3_detect_batch (device contour stage, one K3 launch per call): from the second call of a geometry
on, K3's second half, the gather of its quads, K2 and K4 are queued without a host synchronisation and sized from the
previous call.  Results must be identical whichever route a call takes — sizes held, K3's lists outgrew the speculation,
the quad list outgrew it, or a frame was flagged for the host stage."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def a3():
    import aruco3_b200
    from aruco3_b200 import _ffi
    if _ffi.lib().a3_device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests have no fallback")
    return aruco3_b200


def _summary(dets):
    return [([list(sum(c, ())) for c in x.candidates],
             [(m.candidate, m.id, m.rotation, m.hamming_distance, m.code, m.corners) for m in x.markers],
             [None if h is None else h.tobytes() for h in (x.homographies or [])]) for x in dets]


def _fresh(a3, frames, **kw):
    with a3.Detector(**kw) as d:
        out = _summary(d.detect_batch(frames, full=True))
        assert d.last_stats["one_shot"] == 0  # no history yet
    return out


def test_repeat_calls_take_the_one_shot_route(a3, oracle):
    from aruco3_b200 import synth
    frames, _ = synth.render_batch("C1", 6)
    want = _fresh(a3, frames)
    for f in range(len(frames)):  # and that is the oracle's answer
        ref = oracle.detect(frames[f], "ARUCO")
        assert want[f][0] == ref.candidates.tolist()
        assert [m[1:5] for m in want[f][1]] == [(m["id"], m["rotation"], m["hamming_distance"], m["code"]) for m in ref.markers]
    with a3.Detector() as d:
        d.set_pose(40.0)
        first = d.detect_batch(frames, full=True)
        assert (d.last_stats["one_shot"], d.last_stats["one_shot_retry"]) == (0, 0)
        for _ in range(3):
            again = d.detect_batch(frames, full=True)
            assert (d.last_stats["one_shot"], d.last_stats["one_shot_retry"]) == (1, 0)
            assert _summary(again) == want
            assert [[(p.error, p.rotation.tobytes(), p.translation.tobytes()) for m in x.markers for p in m.poses] for x in again] == \
                   [[(p.error, p.rotation.tobytes(), p.translation.tobytes()) for m in x.markers for p in m.poses] for x in first]
        assert d.last_stats["decode_kernel_launches"] == 1 and d.last_stats["pose_kernel_launches"] == 1
        assert d.last_stats["n_candidates"] == sum(len(x[0]) for x in want)


def test_k3_lists_outgrow_the_speculation(a3):
    """Flat frames (no border at all), marker frames and pure-noise frames (tens of thousands of long borders) alternate
    in one detector: going up the speculation fails and the call is finished the ordinary way, going down it holds."""
    from aruco3_b200 import synth
    clean, _ = synth.render_batch("C1", 5)
    noise = np.random.default_rng(9).integers(0, 256, size=clean.shape, dtype=np.uint8)
    flat = np.full_like(clean, 200)
    want = {"clean": _fresh(a3, clean), "noise": _fresh(a3, noise), "flat": [([], [], [])] * 5}
    inputs = {"clean": clean, "noise": noise, "flat": flat}
    routes = []
    with a3.Detector() as d:
        for name in ("flat", "flat", "noise", "clean", "noise", "noise", "flat", "clean"):
            assert _summary(d.detect_batch(inputs[name], full=True)) == want[name], name
            routes.append((name, d.last_stats["one_shot"], d.last_stats["one_shot_retry"]))
    assert routes[0][1:] == (0, 0) and all(r[1] + r[2] == 1 for r in routes[1:]), routes
    assert routes[1][1:] == (1, 0), routes   # flat after flat
    assert routes[2][1:] == (0, 1), routes   # noise after flat cannot fit
    assert routes[3][1:] == (1, 0), routes   # clean after noise does
    assert routes[4][1:] == (0, 1), routes   # noise after clean cannot
    assert routes[5][1:] == (1, 0), routes   # noise after noise
    assert routes[6][1:] == (1, 0), routes   # flat after noise


def test_quad_list_outgrows_the_speculation(a3):
    """Many borders but few quads first (noise), then the decode-stress scene with 200+ quads in the same geometry."""
    from aruco3_b200 import synth
    spec = synth.CONFIGS["C5"]
    stress, _ = synth.render_batch(spec, 2)  # two frames: more quads than the 256 of headroom over the noise frames' few
    noise = np.random.default_rng(5).integers(0, 256, size=stress.shape, dtype=np.uint8)
    cfg = a3.DetectorConfig(min_corner_separation_factor=spec.min_corner_separation_factor)
    want = _fresh(a3, stress, config=cfg, dictionary=spec.dictionary)
    assert len(want[0][0]) + len(want[1][0]) > 300
    with a3.Detector(cfg, spec.dictionary) as d:
        d.detect_batch(noise)
        d.detect_batch(noise)
        assert d.last_stats["one_shot"] == 1
        assert _summary(d.detect_batch(stress, full=True)) == want
        assert (d.last_stats["one_shot"], d.last_stats["one_shot_retry"]) == (0, 1)
        assert _summary(d.detect_batch(stress, full=True)) == want
        assert (d.last_stats["one_shot"], d.last_stats["one_shot_retry"]) == (1, 0)


def test_flagged_frames_leave_the_one_shot_route(a3):
    img = np.full((3, 96, 128), 200, np.uint8)
    img[:, 30:50, 0:20] = 20          # dark blob on the left edge: K3 flags the frame (barred start)
    img[1, 10:40, 60:100] = 30
    want = _fresh(a3, img)
    with a3.Detector() as d:
        for call in range(3):
            assert _summary(d.detect_batch(img, full=True)) == want
            assert d.last_stats["host_fallback_frames"] >= 1 and d.last_stats["one_shot"] == 0
            if call:
                assert d.last_stats["one_shot_retry"] == 1


def test_batches_of_changing_size(a3):
    """A change of batch size drops the history: the first call of a geometry is never one-shot, the second is."""
    from aruco3_b200 import synth
    frames, _ = synth.render_batch("C1", 9)
    want = _fresh(a3, frames)
    routes = []
    with a3.Detector() as d:
        for n in (9, 9, 4, 4, 9, 9):
            assert _summary(d.detect_batch(frames[:n], full=True)) == want[:n]
            routes.append(d.last_stats["one_shot"])
    assert routes == [0, 1, 0, 1, 0, 1]


def _markers_only(dets):
    return [[(m.candidate, m.id, m.rotation, m.hamming_distance, m.code, m.corners) for m in x.markers] for x in dets]


def test_markers_only_calls_assemble_on_the_device(a3):
    """Without per-candidate outputs the one-shot route also builds the a3_marker records (and their poses) on the device;
    they must equal the host-assembled ones of a call that asks for everything, in order, frame offsets included."""
    from aruco3_b200 import synth
    frames, _ = synth.render_batch("C1", 7)
    frames[3] = 200  # a frame without markers in the middle: its offset range is empty
    with a3.Detector() as d:
        d.set_pose(40.0)
        full = d.detect_batch(frames, full=True)
    want = _markers_only(full)
    want_poses = [[(p.error, p.rotation.tobytes(), p.translation.tobytes()) for m in x.markers for p in m.poses] for x in full]
    assert want[3] == [] and sum(len(x) for x in want) > 20
    with a3.Detector() as d:
        d.set_pose(40.0)
        for call in range(3):
            got = d.detect_batch(frames)
            assert d.last_stats["one_shot"] == (1 if call else 0)
            assert _markers_only(got) == want
            assert [[(p.error, p.rotation.tobytes(), p.translation.tobytes()) for m in x.markers for p in m.poses] for x in got] == want_poses
        # a marker buffer that is too small: the count comes back, the second attempt of the wrapper succeeds
        got = d.detect_batch(frames, marker_capacity=5)
        assert _markers_only(got) == want
    with a3.Detector() as d:  # and without the pose step
        for call in range(2):
            assert _markers_only(d.detect_batch(frames)) == want


def test_sharded_detector_in_one_process(a3):
    """aruco3_b200.sharding.ShardedDetector: three shards (all on device 0 here), one host thread each, results in frame
    order equal to a single detector's — first call and repeat calls (each shard then runs its own one-shot route)."""
    from aruco3_b200 import synth
    from aruco3_b200.sharding import ShardedDetector
    frames, _ = synth.render_batch("C1", 8)
    with a3.Detector() as d:
        want = _markers_only(d.detect_batch(frames))
    with ShardedDetector(devices=(0, 0, 0)) as sd:
        for _ in range(3):
            assert _markers_only(sd.detect_batch(frames)) == want
        assert _summary(sd.detect_batch(frames[:2], full=True)) == _fresh(a3, frames[:2])  # fewer frames than shards


def test_more_than_400k_candidates_in_one_call(a3, oracle):
    """Marker assembly on the device must not depend on how many of its CTAs are resident at once (it used to chain them
    with a busy-wait): 3400 frames x 121 dark squares = 411 400 candidates in ONE call, 402 chunks of 1024 candidates against
    296 resident CTAs.  With APRILTAG_36H11 an all-black interior is accepted as id 191, so every candidate
    becomes a marker; the device-assembled records of the second call must equal the host-assembled ones of the first and
    the oracle's for the (repeated) frame."""
    import ctypes as C
    from aruco3_b200 import _ffi
    n, side, per = 3400, 256, 121
    one = np.full((side, side), 200, np.uint8)
    for cy in range(11):
        for cx in range(11):
            one[10 + 21 * cy: 22 + 21 * cy, 10 + 21 * cx: 22 + 21 * cx] = 20
    ocfg = oracle.default_config()
    ocfg.min_corner_separation_factor = 0.03
    ref = oracle.detect(one, "APRILTAG_36H11", ocfg)
    assert len(ref.markers) == per and all(m["id"] == 191 for m in ref.markers)
    torch = pytest.importorskip("torch")  # plumbing only: resident frames, so that ONE K3 / K2 launch covers the whole call
    frames = torch.from_numpy(one).cuda().unsqueeze(0).repeat(n, 1, 1).contiguous()
    cfg = a3.DetectorConfig(min_corner_separation_factor=0.03)
    L = _ffi.lib()
    cap = (per + 7) * n
    with a3.Detector(cfg, "APRILTAG_36H11") as d:
        results = []
        for call in range(3):
            markers = (_ffi.A3Marker * cap)()
            n_markers = C.c_uint32()
            stats = _ffi.A3Stats()
            _ffi.check(L.a3_detect_batch(d._h, frames.data_ptr(), _ffi.FMT_LUMA8, _ffi.MEM_DEVICE, n, side, side, side, side * side,
                                         C.cast(markers, C.c_void_p), cap, C.byref(n_markers), None, C.byref(stats)))
            assert n_markers.value == per * n and stats.n_candidates == per * n
            assert stats.one_shot == (1 if call else 0), (call, stats.one_shot, stats.one_shot_retry)
            results.append(np.frombuffer(markers, dtype=np.uint8, count=n_markers.value * C.sizeof(_ffi.A3Marker)).copy())
        assert np.array_equal(results[0], results[1]) and np.array_equal(results[0], results[2])
        rec = np.frombuffer(results[1], dtype=np.dtype([("id", "<u8"), ("code", "<u8"), ("corners", "<u4", 8), ("frame", "<u4"),
                                                         ("candidate", "<u4"), ("hd", "u1"), ("rot", "u1"), ("pad", "u1", 6)]))
        assert np.array_equal(rec["frame"], np.repeat(np.arange(n, dtype=np.uint32), per))
        want_corners = np.array([m["corners"] for m in ref.markers], np.uint32)
        for f in (0, 1, n // 2, n - 1):
            assert np.array_equal(rec["corners"][per * f: per * (f + 1)], want_corners)
            assert rec["id"][per * f: per * (f + 1)].tolist() == [191] * per
            assert rec["rot"][per * f: per * (f + 1)].tolist() == [m["rotation"] for m in ref.markers]
