"""The reference's own pose tests (src/pose.rs:379-598), restated against the oracle: these golden vectors PIN the
pose step.  Tolerances are the reference's."""
import numpy as np
import pytest

from oracle import a3ref_py as o

SQUARE_PTS = [(0.1, 0.1), (0.3, 0.1), (0.3, 0.3), (0.1, 0.3)]

# src/pose.rs:523-540 (test_e2e_pose) and :575-592 (test_e2e_pose2)
E2E_A = dict(t=[20.32196265994096, 29.69316666108512, 238.3658341694123],
             r=[[0.07313995850727262, 0.2953796077825095, 0.9525762089070907],
                [0.9973210134149258, -0.02055233410014844, -0.07020254813082821],
                [-0.001158736630905738, 0.9551588814795613, -0.2960914866390682]])
E2E_B = dict(t=[19.85146615649354, 29.20013946746331, 234.3277337340188],
             r=[[0.05174977302896467, 0.1311239186581316, -0.9900143832021767],
                [0.9667844474723887, -0.2550432732960733, 0.01675592050389792],
                [-0.2502994069448807, -0.957997623536802, -0.1399669967559523]])
E2E2_A = dict(t=[-22.712781796404, -33.18648038591866, 266.408873483460],
              r=[[-0.07313995850727262, -0.2953796077825095, -0.9525762089070907],
                 [-0.9973210134149258, 0.02055233410014844, 0.07020254813082821],
                 [-0.001158736630905738, 0.9551588814795613, -0.2960914866390682]])
E2E2_B = dict(t=[-22.18693276313984, -32.6354499930472, 261.8957024086092],
              r=[[-0.05174977302896467, -0.1311239186581316, 0.9900143832021767],
                 [-0.9667844474723887, 0.2550432732960733, -0.01675592050389792],
                 [-0.2502994069448807, -0.957997623536802, -0.1399669967559523]])


def test_marker_transforms():  # src/pose.rs:379-392
    pose = o.Pose()
    o._pose_lib().a3ref_pose_default(pose)
    assert pose.error == np.float32(1e31) and list(pose.rotation) == [1, 0, 0, 0, 1, 0, 0, 0, 1]
    pose.translation[:] = [1.0, 2.0, 3.0]
    pose.rotation[:] = [0, 0, 1, 0, 1, 0, 1, 0, 0]
    out = o.pose_apply(pose, [(0, 0, 0), (7, 11, 13)])
    assert out.tolist() == [[1, 2, 3], [14, 13, 10]]


def test_marker_identity_random():  # src/pose.rs:394-439 (seeded here)
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(100):
        pose = o.Pose()
        r1 = np.array([1 + rng.random(), 1 + rng.random(), 0.0], np.float32)
        r2 = np.array([0.0, 1.1 + rng.random(), 1 + rng.random()], np.float32)
        r1 /= np.linalg.norm(r1); r2 /= np.linalg.norm(r2)
        r3 = np.cross(r1, r2); r3 /= np.linalg.norm(r3)
        for _ in range(10):
            r2 = np.cross(r1, r3)
            r1 = np.cross(r3, r2)
        pose.rotation[:] = np.stack([r1, r2, r3], axis=1).astype(np.float32).ravel().tolist()
        pose.translation[:] = rng.random(3).astype(np.float32).tolist()
        pts = rng.random((100, 3)).astype(np.float32)
        back = o.pose_apply(pose, o.pose_apply(pose, pts), inverse=True)
        worst = max(worst, float(np.abs(back - pts).sum(axis=1).max()))
    assert worst <= 1e-5


def test_gen_marker_square():  # src/pose.rs:441-455
    assert o.marker_square(11.0).tolist() == [[-5.5, 5.5, 0], [5.5, 5.5, 0], [5.5, -5.5, 0], [-5.5, -5.5, 0]]


def test_homography_solve():  # src/pose.rs:457-474
    expected = np.array([[0.01818181818181819, 0.0, 0.2],
                         [9.856383386231859e-19, -0.01818181818181819, 0.2000000000000001],
                         [1.577021341797097e-17, -1.577021341797097e-17, 1.0]])
    h = o.homography_from_marker_square(11.0, SQUARE_PTS)
    assert np.abs(h - expected).sum() < 1e-5


def test_canonical_solve():  # src/pose.rs:476-512
    pa, pb = o.solve_canonical_form(11.0, SQUARE_PTS)
    exp_a = np.array([[1.0, -2.775557561562891e-17, 1.02695629777827e-15, 10.99999999999999],
                      [7.632783294297951e-17, -1.0, 1.02695629777827e-15, 11.0],
                      [1.02695629777827e-15, -9.992007221626409e-16, -1.0, 54.99999999999996]])
    exp_b = np.array([[0.9259259259259256, 0.07407407407407443, -0.3703703703703712, 10.79629629629629],
                      [-0.0740740740740744, -0.9259259259259256, -0.3703703703703713, 10.79629629629629],
                      [-0.3703703703703712, 0.3703703703703713, -0.8518518518518512, 54.99999999999999]])
    for p, e in ((pa, exp_a), (pb, exp_b)):
        _, r, t = p.as_tuple()
        assert np.abs(r - e[:, :3]).sum() < 1e-5
        assert np.abs(t - e[:, 3]).sum() < 1e-4


def test_e2e_pose():  # src/pose.rs:514-552
    pa, pb = o.solve_with_undistorted_points([(90, 89), (95, 150), (80, 170), (75, 90)], 17.0, (1000, 1000))
    for p, e in ((pa, E2E_A), (pb, E2E_B)):
        _, r, t = p.as_tuple()
        assert np.abs(r - np.array(e["r"])).sum() < 2e-5
        assert np.abs(t - np.array(e["t"])).sum() < 0.0005
    assert pa.error <= pb.error


def test_e2e_pose2():  # src/pose.rs:554-598
    pts = [(-0.090, -0.089), (-0.095, -0.150), (-0.080, -0.170), (-0.075, -0.090)]
    expected_h = np.array([[0.0001197249881460392, -0.00193812233285917, -0.08585585585585585],
                           [-0.003084400189663352, -0.00115457562825984, -0.1225675675675677],
                           [-0.004504504504504568, 0.01351351351351346, 1.0]])
    assert np.abs(o.homography_from_marker_square(19.0, pts) - expected_h).max() <= 1e-5
    pa, pb = o.solve_with_normalized_points(pts, 19.0)
    for p, e in ((pa, E2E2_A), (pb, E2E2_B)):
        _, r, t = p.as_tuple()
        assert np.abs(r - np.array(e["r"])).max() <= 1e-5
        assert np.abs(t - np.array(e["t"])).max() <= 1e-3


def test_solve_with_intrinsics_matches_manual_unproject():  # src/pose.rs:52-55 + src/pinhole.rs:88-93
    k = o.intrinsics_new(1000, 1000, 1000.0, 1000.0, 0.0, 0.0)
    corners = [(90, 89), (95, 150), (80, 170), (75, 90)]
    a1, b1 = o.solve_with_intrinsics(corners, 17.0, k)
    a2, b2 = o.solve_with_undistorted_points(corners, 17.0, (1000, 1000))
    for p, q in ((a1, a2), (b1, b2)):
        assert bytes(p) == bytes(q)


def test_intrinsics_constructors():  # src/pinhole.rs:26-60
    k = o.intrinsics_new(640, 480, 1.0, 1.0)
    assert (k.principal_x, k.principal_y) == (320.0, 240.0)
    k = o.intrinsics_new(640, 480, 2.0, 3.0, 11.0, 12.0)
    assert (k.focal_x, k.focal_y, k.principal_x, k.principal_y) == (2.0, 3.0, 11.0, 12.0)
    k = o.intrinsics_from_fov_horizontal(np.pi / 2, 36.0, 1920, 1080)
    assert k.focal_x == pytest.approx(18.0, rel=1e-6)
    vfov = np.float32(np.pi / 2) / (np.float32(1920) / np.float32(1080))
    assert k.focal_y == pytest.approx(float((36.0 / (1920 / 1080) * 0.5) / np.tan(vfov * 0.5)), rel=1e-5)
    assert (k.principal_x, k.principal_y) == (960.0, 540.0)
