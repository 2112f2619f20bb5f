"""Pose step on the device (kernel K4) through the C ABI.

Bar: the reference's own golden vectors (src/pose.rs:457-598) within the reference's own tolerances, and BIT-EXACT
against the oracle (same f32 operation order, no contraction) on random quads and on the markers of detected frames.
"""
import numpy as np
import pytest

from tests import test_oracle_pose as gold

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def a3():
    import aruco3_b200
    from aruco3_b200 import _ffi
    if _ffi.lib().a3_device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests have no fallback")
    return aruco3_b200


@pytest.fixture(scope="module")
def det(a3):
    d = a3.Detector(dictionary="ARUCO", device=0)
    yield d
    d.close()


def _same(p, q):
    """MarkerPose (product) == Pose (oracle), bit for bit; a NaN (degenerate quad) must be a NaN on both sides — its
    sign / payload bits are not arithmetic (x86 produces the negative quiet NaN, the GPU the canonical one)."""
    e, r, t = q.as_tuple()
    got = np.concatenate([[np.float32(p.error)], p.rotation.ravel(), p.translation]).astype(np.float32)
    want = np.concatenate([[e], r.ravel(), t]).astype(np.float32)
    nan = np.isnan(want)
    return bool(np.array_equal(np.isnan(got), nan) and got[~nan].tobytes() == want[~nan].tobytes())


def test_canonical_solve_golden(a3, det):  # src/pose.rs:476-512 through the public entry point
    pa, pb = a3.pose.solve_with_normalized_points(det, gold.SQUARE_PTS, 11.0)
    exp_a = np.array([[1.0, 0, 0, 11.0], [0, -1.0, 0, 11.0], [0, 0, -1.0, 55.0]])
    exp_b = np.array([[0.9259259259259256, 0.07407407407407443, -0.3703703703703712, 10.79629629629629],
                      [-0.0740740740740744, -0.9259259259259256, -0.3703703703703713, 10.79629629629629],
                      [-0.3703703703703712, 0.3703703703703713, -0.8518518518518512, 54.99999999999999]])
    # solve_with_normalized_points orders by error; the frontal pose reprojects exactly
    assert pa.error <= pb.error
    for p, e in ((pa, exp_a), (pb, exp_b)):
        assert np.abs(p.rotation - e[:, :3]).sum() < 1e-5
        assert np.abs(p.translation - e[:, 3]).sum() < 1e-4


def test_e2e_pose_golden(a3, det):  # src/pose.rs:514-552
    pa, pb = a3.pose.solve_with_undistorted_points(det, [(90, 89), (95, 150), (80, 170), (75, 90)], 17.0, (1000, 1000))
    for p, e in ((pa, gold.E2E_A), (pb, gold.E2E_B)):
        assert np.abs(p.rotation - np.array(e["r"])).sum() < 2e-5
        assert np.abs(p.translation - np.array(e["t"])).sum() < 0.0005


def test_e2e_pose2_golden(a3, det):  # src/pose.rs:554-598
    pts = [(-0.090, -0.089), (-0.095, -0.150), (-0.080, -0.170), (-0.075, -0.090)]
    pa, pb = a3.pose.solve_with_normalized_points(det, pts, 19.0)
    for p, e in ((pa, gold.E2E2_A), (pb, gold.E2E2_B)):
        assert np.abs(p.rotation - np.array(e["r"])).max() <= 1e-5
        assert np.abs(p.translation - np.array(e["t"])).max() <= 1e-3


def _random_quads(rng, n, w, h):
    """Roughly square, clockwise quads with perspective jitter, plus a share of degenerate ones (collinear, repeated)."""
    c = np.stack([rng.uniform(0.1 * w, 0.9 * w, n), rng.uniform(0.1 * h, 0.9 * h, n)], axis=1)
    s = rng.uniform(6, 0.2 * min(w, h), n)
    a = rng.uniform(0, 2 * np.pi, n)
    base = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], float)
    rot = np.stack([np.stack([np.cos(a), -np.sin(a)], -1), np.stack([np.sin(a), np.cos(a)], -1)], -2)
    q = c[:, None, :] + s[:, None, None] * np.einsum("nij,kj->nki", rot, base) + rng.uniform(-0.15, 0.15, (n, 4, 2)) * s[:, None, None]
    q = np.clip(np.rint(q), 0, [w - 1, h - 1]).astype(np.uint32)
    q[::97, 2] = q[::97, 1]         # repeated corner
    q[5::101, :, 1] = q[5::101, :1, 1]  # all on one row
    return q


def test_random_quads_bit_exact_with_oracle(a3, det, oracle):
    rng = np.random.default_rng(11)
    n, w, h = 4000, 1920, 1080
    quads = _random_quads(rng, n, w, h)
    k = a3.CameraIntrinsics.new_from_fov_horizontal(1.2, 10.0, w, h)
    ko = oracle.intrinsics_from_fov_horizontal(1.2, 10.0, w, h)
    with np.errstate(all="ignore"):
        b1, a1 = a3.pose.solve_with_undistorted_points(det, quads, 40.0, (w, h))
        b2, a2 = a3.pose.solve_with_intrinsics(det, quads, 25.0, k)
        norm = (quads.astype(np.float32) / np.float32([w, h])).astype(np.float32)
        b3, a3_ = a3.pose.solve_with_normalized_points(det, norm, 40.0)
    finite = 0
    for i in range(n):
        ob, oa = oracle.solve_with_undistorted_points(quads[i], 40.0, (w, h))
        assert _same(b1[i], ob) and _same(a1[i], oa), f"undistorted quad {i}: {quads[i].tolist()}"
        ob, oa = oracle.solve_with_intrinsics(quads[i], 25.0, ko)
        assert _same(b2[i], ob) and _same(a2[i], oa), f"intrinsics quad {i}: {quads[i].tolist()}"
        ob, oa = oracle.solve_with_normalized_points(norm[i], 40.0)
        assert _same(b3[i], ob) and _same(a3_[i], oa), f"normalized quad {i}"
        assert _same(b3[i], oracle.solve_with_undistorted_points(quads[i], 40.0, (w, h))[0])  # same points, same result
        if np.isfinite(b1[i].error):
            finite += 1
            assert b1[i].error <= a1[i].error
    assert finite > 0.9 * n


def test_empty_and_single(a3, det):
    b, a = a3.pose.solve_with_normalized_points(det, np.zeros((0, 4, 2), np.float32), 10.0)
    assert b == [] and a == []
    one = a3.pose.solve_with_normalized_points(det, gold.SQUARE_PTS, 11.0)
    many_b, many_a = a3.pose.solve_with_normalized_points(det, np.array([gold.SQUARE_PTS] * 3, np.float32), 11.0)
    assert all(m.rotation.tobytes() == one[0].rotation.tobytes() for m in many_b)
    assert all(m.translation.tobytes() == one[1].translation.tobytes() for m in many_a)


@pytest.mark.parametrize("mode", ["undistorted", "intrinsics"])
@pytest.mark.parametrize("contours", ["device", "host"])
def test_pipeline_poses_match_oracle(a3, oracle, mode, contours):
    """detect_batch with a pose mode: every marker carries the pose pair the reference's examples compute right after
    `detect` (examples/webcam_kamera.rs:68, examples/macroquad_detect.rs:150), bit-exact with the oracle; markers
    themselves are unchanged."""
    from aruco3_b200 import synth
    frames, _ = synth.render_batch("C1", 5)
    h, w = frames.shape[1:3]
    k = a3.CameraIntrinsics.new_from_fov_horizontal(1.0, 10.0, w, h)
    ko = oracle.intrinsics_from_fov_horizontal(1.0, 10.0, w, h)
    with a3.Detector(dictionary="ARUCO", contours=contours) as d:
        plain = d.detect_batch(frames)
        d.set_pose(40.0, k if mode == "intrinsics" else None)
        got = d.detect_batch(frames)
        assert d.last_stats["pose_kernel_launches"] >= 1
        d.set_pose(None)
        off = d.detect_batch(frames)
    n = 0
    for f in range(frames.shape[0]):
        assert [(m.id, m.corners, m.rotation) for m in got[f].markers] == [(m.id, m.corners, m.rotation) for m in plain[f].markers]
        assert all(m.poses is None for m in off[f].markers)
        for m in got[f].markers:
            if mode == "intrinsics":
                ob, oa = oracle.solve_with_intrinsics(m.corners, 40.0, ko)
            else:
                ob, oa = oracle.solve_with_undistorted_points(m.corners, 40.0, (w, h))
            assert _same(m.poses[0], ob) and _same(m.poses[1], oa), f"frame {f} marker {m.id}"
            n += 1
    assert n >= 10


def test_pose_recovers_rendered_geometry(a3):
    """Property at BASELINE size (1080p frames): a fronto-parallel square of side s px seen by a camera of focal f px lies
    at depth f * size / s; the best pose's translation.z must say so (within the corner quantisation), its rotation must be
    orthonormal and its reprojection error small."""
    from aruco3_b200 import synth
    frames, truth = synth.render_batch("C3", 4)
    h, w = frames.shape[1:3]
    f_px = 1500.0
    k = a3.CameraIntrinsics(w, h, f_px, f_px)
    with a3.Detector(dictionary="ARUCO") as d:
        d.set_pose(50.0, k)
        got = d.detect_batch(frames)
    checked = 0
    for f in range(frames.shape[0]):
        for m in got[f].markers:
            best, alt = m.poses
            if not np.isfinite(best.error):
                continue
            c = np.array(m.corners, float)
            side = np.mean([np.linalg.norm(c[i] - c[(i + 1) % 4]) for i in range(4)])
            r = best.rotation.astype(float)
            assert np.abs(r @ r.T - np.eye(3)).max() < 1e-3 and abs(np.linalg.det(r) - 1) < 1e-3
            assert best.error <= alt.error and best.error < 0.05
            assert abs(best.translation[2] - f_px * 50.0 / side) < 0.12 * f_px * 50.0 / side
            centre = c.mean(axis=0)
            assert abs(best.translation[0] / best.translation[2] - (centre[0] - w / 2) / f_px) < 0.02
            checked += 1
    assert checked >= 40
