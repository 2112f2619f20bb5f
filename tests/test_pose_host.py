"""Host-only parts of the pose API (constructors, one-point conversions, pose transforms) against the reference's own
tests and the oracle; the solvers themselves need the device (tests/test_gpu_pose.py)."""
import ctypes as C

import numpy as np
import pytest

import aruco3_b200 as a3
from aruco3_b200 import _ffi


def test_marker_transforms():  # /root/reference/src/pose.rs:379-392
    pose = a3.MarkerPose()
    assert pose.error == np.float32(1e31) and np.array_equal(pose.rotation, np.eye(3)) and not pose.translation.any()
    pose.translation[:] = [1.0, 2.0, 3.0]
    pose.rotation = np.array([[0, 0, 1], [0, 1, 0], [1, 0, 0]], np.float32)
    assert pose.apply_transform_to_points([(0, 0, 0), (7, 11, 13)]) == [(1, 2, 3), (14, 13, 10)]
    assert pose.apply_inverse_transform_to_points([(14, 13, 10)]) == [(7, 11, 13)]


def test_pose_default_matches_library():
    p = _ffi.A3Pose()
    _ffi.lib().a3_pose_default(C.byref(p))
    assert p.error == np.float32(1e31) and list(p.rotation) == [1, 0, 0, 0, 1, 0, 0, 0, 1] and list(p.translation) == [0, 0, 0]


def test_transforms_bit_exact_with_oracle(oracle):
    rng = np.random.default_rng(7)
    for _ in range(50):
        q, _r = np.linalg.qr(rng.standard_normal((3, 3)))
        pose = a3.MarkerPose(0.0, q.astype(np.float32), rng.standard_normal(3).astype(np.float32) * 100)
        po = oracle.Pose()
        po.rotation[:] = pose.rotation.ravel().tolist()
        po.translation[:] = pose.translation.tolist()
        pts = (rng.standard_normal((40, 3)) * 50).astype(np.float32)
        for inverse in (False, True):
            want = oracle.pose_apply(po, pts, inverse)
            got = np.array(pose._apply(pts, inverse), np.float32)
            assert got.tobytes() == want.tobytes()
        back = np.array(pose.apply_inverse_transform_to_points(pose.apply_transform_to_points(pts)), np.float32)
        assert np.abs(back - pts).sum(axis=1).max() < 1e-3  # src/pose.rs:394-439 at these magnitudes


def test_camera_intrinsics_match_oracle(oracle):  # /root/reference/src/pinhole.rs:26-94
    k = a3.CameraIntrinsics(640, 480, 1.0, 1.0)
    assert (k.image_width, k.image_height, k.principal_x, k.principal_y) == (640, 480, 320.0, 240.0)
    rng = np.random.default_rng(3)
    for _ in range(20):
        w, h = int(rng.integers(16, 4096)), int(rng.integers(16, 4096))
        fov, sw = float(rng.uniform(0.3, 2.5)), float(rng.uniform(1, 40))
        k = a3.CameraIntrinsics.new_from_fov_horizontal(fov, sw, w, h)
        ko = oracle.intrinsics_from_fov_horizontal(fov, sw, w, h)
        assert bytes(k._c) == bytes(ko)
        k2 = a3.CameraIntrinsics(w, h, k.focal_x, k.focal_y, float(rng.uniform(0, w)), None)
        ko2 = oracle.intrinsics_new(w, h, k.focal_x, k.focal_y, k2.principal_x, None)
        assert bytes(k2._c) == bytes(ko2) and k2.principal_y == np.float32(h) / np.float32(2)
        x, y, z = (float(v) for v in rng.standard_normal(3).astype(np.float32))
        out3, out2 = (C.c_float * 3)(), (C.c_float * 2)()
        L = oracle._pose_lib()
        L.a3ref_project(C.byref(ko), x, y, z, out3)
        assert k.project(x, y, z) == tuple(out3)
        ok = L.a3ref_project_culled(C.byref(ko), x, y, z, out2)
        assert k.project_culled(x, y, z) == (tuple(out2) if ok else None)
        L.a3ref_unproject(C.byref(ko), x * 100, y * 100, out2)
        assert k.unproject(x * 100, y * 100) == tuple(out2)
    assert k.project_culled(1.0, 1.0, 0.0) is None
    m = a3.CameraIntrinsics(640, 480, 2.0, 3.0).to_matrix3x4()
    assert m.shape == (3, 4) and m[0, 0] == 2 and m[1, 1] == 3 and m[0, 2] == 320 and m[1, 2] == 240 and m[2, 2] == 1


def test_solvers_fail_loudly_without_a_device():
    if _ffi.lib().a3_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(a3.A3Error):
        a3.Detector()
