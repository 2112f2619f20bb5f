"""ctypes view of oracle/liba3ref.so — TEST INFRASTRUCTURE ONLY (see oracle/a3ref.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
FMT_RGB8, FMT_RGBA8, FMT_LUMA8, FMT_BGR8, FMT_BGRA8 = 0, 1, 2, 3, 4
FMT_LUMAA8, FMT_LUMA16, FMT_LUMAA16, FMT_RGB16, FMT_RGBA16 = 5, 6, 7, 8, 9


class Config(C.Structure):
    _fields_ = [("threshold_window", C.c_uint32), ("contour_simplification_epsilon", C.c_double),
                ("min_side_length_factor", C.c_float), ("min_corner_separation_factor", C.c_float),
                ("homography_sample_size", C.c_uint32), ("filter_high_bit_errors", C.c_uint8)]


class Dictionary(C.Structure):
    _fields_ = [("num_bits", C.c_uint8), ("tau", C.c_uint8), ("n_codes", C.c_uint32),
                ("codes", C.POINTER(C.c_uint64))]


class Marker(C.Structure):
    _fields_ = [("id", C.c_uint64), ("code", C.c_uint64), ("corners", C.c_uint32 * 8), ("candidate", C.c_uint32),
                ("hamming_distance", C.c_uint8), ("rotation", C.c_uint8), ("pad", C.c_uint8 * 2)]


class Contours(C.Structure):
    _fields_ = [("n_contours", C.c_uint32), ("n_points", C.c_uint32), ("offsets", C.POINTER(C.c_uint32)),
                ("points", C.POINTER(C.c_uint32)), ("is_outer", C.POINTER(C.c_uint8))]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("n_contours", "n_contour_points", "reject_point_count", "reject_convexity",
                                          "reject_edge_length", "n_candidates_before_discard", "n_candidates",
                                          "n_border_pass", "n_markers")] + \
               [(n, C.c_double) for n in ("ms_gray", "ms_threshold", "ms_contours", "ms_quads", "ms_warp", "ms_decode",
                                          "ms_total")]


class Detection(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("grey", C.POINTER(C.c_uint8)),
                ("mask", C.POINTER(C.c_uint8)), ("n_candidates", C.c_uint32), ("candidates", C.POINTER(C.c_uint32)),
                ("patch_size", C.c_uint32), ("homographies", C.POINTER(C.c_uint8)),
                ("homography_ok", C.POINTER(C.c_uint8)), ("otsu", C.POINTER(C.c_uint8)),
                ("has_codes", C.POINTER(C.c_uint8)), ("codes", C.POINTER(C.c_uint64)), ("n_markers", C.c_uint32),
                ("markers", C.POINTER(Marker)), ("stats", Stats)]


_lib = None


def build() -> Path:
    subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
    return HERE / "liba3ref.so"


def lib():
    global _lib
    if _lib is None:
        so = HERE / "liba3ref.so"
        if not so.exists():
            build()
        L = C.CDLL(str(so))
        L.a3ref_dictionary_name.restype = C.c_char_p
        L.a3ref_hamming_distance.restype = C.c_uint8
        L.a3ref_hamming_distance.argtypes = [C.c_uint64, C.c_uint64]
        L.a3ref_calculate_tau.restype = C.c_uint8
        L.a3ref_mark_size.restype = C.c_uint8
        L.a3ref_find_nearest.argtypes = [C.POINTER(Dictionary), C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)]
        L.a3ref_try_find_nearest.argtypes = L.a3ref_find_nearest.argtypes
        L.a3ref_make_binary_image.argtypes = [C.POINTER(Dictionary), C.c_uint64, C.c_void_p, C.c_uint32]
        L.a3ref_make_binary_image.restype = C.c_uint32
        L.a3ref_to_luma8.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_size_t, C.c_void_p]
        L.a3ref_adaptive_threshold.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.a3ref_find_contours.restype = C.POINTER(Contours)
        L.a3ref_find_contours.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.a3ref_contours_free.argtypes = [C.POINTER(Contours)]
        L.a3ref_approximate_polygon_dp.restype = C.c_size_t
        L.a3ref_approximate_polygon_dp.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_void_p]
        L.a3ref_convex_hull.restype = C.c_size_t
        L.a3ref_convex_hull.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.a3ref_contours_to_candidates.restype = C.c_uint32
        L.a3ref_contours_to_candidates.argtypes = [C.POINTER(Contours), C.c_uint32, C.c_double,
                                                   C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(Stats)]
        L.a3ref_enforce_clockwise_corners.argtypes = [C.c_void_p, C.c_uint32]
        L.a3ref_discard_too_near.restype = C.c_uint32
        L.a3ref_discard_too_near.argtypes = [C.c_void_p, C.c_uint32, C.c_float]
        L.a3ref_perimeter.restype = C.c_float
        L.a3ref_perimeter.argtypes = [C.c_void_p]
        L.a3ref_projection_from_control_points.argtypes = [C.c_void_p] * 4 + [C.POINTER(C.c_int)]
        L.a3ref_extract_homography.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        L.a3ref_otsu_level.restype = C.c_uint8
        L.a3ref_otsu_level.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.a3ref_resize_triangle.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.a3ref_homography_to_code_permutations.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint8, C.c_void_p,
                                                            C.c_void_p, C.c_void_p]
        L.a3ref_rotate_bit_matrix.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.a3ref_default_config.argtypes = [C.POINTER(Config)]
        L.a3ref_detect.restype = C.POINTER(Detection)
        L.a3ref_detect.argtypes = [C.POINTER(Config), C.POINTER(Dictionary), C.c_void_p, C.c_int, C.c_uint32,
                                   C.c_uint32, C.c_size_t]
        L.a3ref_detection_free.argtypes = [C.POINTER(Detection)]
        L.a3ref_detect_many.restype = C.c_uint64
        L.a3ref_detect_many.argtypes = [C.POINTER(Config), C.POINTER(Dictionary), C.c_void_p, C.c_int, C.c_uint32,
                                        C.c_uint32, C.c_uint32, C.c_size_t, C.c_size_t, C.c_uint32, C.POINTER(Stats)]
        _lib = L
    return _lib


def default_config(**overrides) -> Config:
    cfg = Config()
    lib().a3ref_default_config(C.byref(cfg))
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def dictionary(name: str) -> Dictionary:
    d = Dictionary()
    if lib().a3ref_dictionary_by_name(name.encode(), C.byref(d)) != 0:
        raise KeyError(name)
    return d


def find_nearest(d: Dictionary, bits: int):
    idx, dist = C.c_uint64(), C.c_uint8()
    lib().a3ref_find_nearest(C.byref(d), bits, C.byref(idx), C.byref(dist))
    return idx.value, dist.value


def try_find_nearest(d: Dictionary, bits: int):
    idx, dist = C.c_uint64(), C.c_uint8()
    ok = lib().a3ref_try_find_nearest(C.byref(d), bits, C.byref(idx), C.byref(dist))
    return (idx.value, dist.value) if ok else None


def _fmt_of(img: np.ndarray, order: str = "rgb") -> int:
    if img.dtype == np.uint16:  # Luma16 [H,W], LumaA16 [H,W,2], Rgb16 [H,W,3], Rgba16 [H,W,4]
        return FMT_LUMA16 if img.ndim == 2 else {2: FMT_LUMAA16, 3: FMT_RGB16, 4: FMT_RGBA16}[img.shape[2]]
    if img.ndim == 3 and img.shape[2] == 2:
        return FMT_LUMAA8
    if img.ndim == 2:
        return FMT_LUMA8
    return {("rgb", 3): FMT_RGB8, ("rgb", 4): FMT_RGBA8, ("bgr", 3): FMT_BGR8, ("bgr", 4): FMT_BGRA8}[(order, img.shape[2])]


def to_luma8(img: np.ndarray, order: str = "rgb") -> np.ndarray:
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    out = np.empty((h, w), np.uint8)
    lib().a3ref_to_luma8(img.ctypes.data, _fmt_of(img, order), w, h, img.strides[0], out.ctypes.data)
    return out


def adaptive_threshold(grey: np.ndarray, radius: int = 7) -> np.ndarray:
    grey = np.ascontiguousarray(grey)
    out = np.empty_like(grey)
    lib().a3ref_adaptive_threshold(grey.ctypes.data, grey.shape[1], grey.shape[0], radius, out.ctypes.data)
    return out


def find_contours(mask: np.ndarray):
    """-> list of int arrays [n,2] (x,y), list of is_outer flags."""
    mask = np.ascontiguousarray(mask)
    c = lib().a3ref_find_contours(mask.ctypes.data, mask.shape[1], mask.shape[0])
    cc = c.contents
    offs = np.ctypeslib.as_array(cc.offsets, (cc.n_contours + 1,)).copy()
    pts = np.ctypeslib.as_array(cc.points, (cc.n_points, 2)).copy() if cc.n_points else np.zeros((0, 2), np.uint32)
    outer = np.ctypeslib.as_array(cc.is_outer, (max(cc.n_contours, 1),))[:cc.n_contours].copy() if cc.n_contours else np.zeros(0, np.uint8)
    lib().a3ref_contours_free(c)
    return [pts[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)], outer


def candidates_from_mask(mask: np.ndarray, cfg: Config | None = None):
    """find_contours -> contours_to_candidates -> enforce_clockwise -> discard_too_near; -> uint32 [n,8]."""
    cfg = cfg or default_config()
    mask = np.ascontiguousarray(mask)
    h, w = mask.shape
    mn = min(w, h)
    min_edge = int(np.float32(mn) * np.float32(cfg.min_side_length_factor))
    min_sep = float(np.float32(mn) * np.float32(cfg.min_corner_separation_factor))
    c = lib().a3ref_find_contours(mask.ctypes.data, w, h)
    quads = C.POINTER(C.c_uint32)()
    st = Stats()
    n = lib().a3ref_contours_to_candidates(c, min_edge, cfg.contour_simplification_epsilon, C.byref(quads), C.byref(st))
    lib().a3ref_contours_free(c)
    lib().a3ref_enforce_clockwise_corners(quads, n)
    n = lib().a3ref_discard_too_near(quads, n, min_sep)
    out = np.ctypeslib.as_array(quads, (max(n, 1), 8))[:n].copy()
    C.CDLL(None).free(quads)
    return out


class Result:
    """Python copy of a3ref_detection (all stages)."""

    def __init__(self, det: Detection):
        w, h, n, ps = det.width, det.height, det.n_candidates, det.patch_size
        arr = np.ctypeslib.as_array
        self.grey = arr(det.grey, (h, w)).copy()
        self.mask = arr(det.mask, (h, w)).copy()
        self.candidates = arr(det.candidates, (max(n, 1), 8))[:n].copy()
        self.homographies = arr(det.homographies, (max(n, 1), ps, ps))[:n].copy()
        self.homography_ok = arr(det.homography_ok, (max(n, 1),))[:n].copy()
        self.otsu = arr(det.otsu, (max(n, 1),))[:n].copy()
        self.has_codes = arr(det.has_codes, (max(n, 1),))[:n].copy()
        self.codes = arr(det.codes, (max(n, 1), 4))[:n].copy()
        self.markers = [dict(id=int(m.id), code=int(m.code), corners=[int(v) for v in m.corners],
                             candidate=int(m.candidate), hamming_distance=int(m.hamming_distance),
                             rotation=int(m.rotation)) for m in (det.markers[i] for i in range(det.n_markers))]
        self.stats = {f: getattr(det.stats, f) for f, _ in Stats._fields_}


def detect(img: np.ndarray, dict_name: str = "ARUCO", cfg: Config | None = None) -> Result:
    cfg = cfg or default_config()
    d = dictionary(dict_name)
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    p = lib().a3ref_detect(C.byref(cfg), C.byref(d), img.ctypes.data, _fmt_of(img), w, h, img.strides[0])
    res = Result(p.contents)
    lib().a3ref_detection_free(p)
    return res


def detect_many(frames: np.ndarray, dict_name: str = "ARUCO", cfg: Config | None = None, threads: int = 1):
    """Frame-parallel CPU baseline; -> (total markers, summed stats dict)."""
    cfg = cfg or default_config()
    d = dictionary(dict_name)
    frames = np.ascontiguousarray(frames)
    n, h, w = frames.shape[:3]
    fmt = FMT_LUMA8 if frames.ndim == 3 else {3: FMT_RGB8, 4: FMT_RGBA8}[frames.shape[3]]
    st = Stats()
    total = lib().a3ref_detect_many(C.byref(cfg), C.byref(d), frames.ctypes.data, fmt, n, w, h, frames.strides[1],
                                    frames.strides[0], threads, C.byref(st))
    return int(total), {f: getattr(st, f) for f, _ in Stats._fields_}


# ---- pose step (oracle/a3ref_pose.c; src/pose.rs, src/pinhole.rs) ----------------------------------------------------
class Pose(C.Structure):
    _fields_ = [("error", C.c_float), ("rotation", C.c_float * 9), ("translation", C.c_float * 3)]

    def as_tuple(self):
        return (np.float32(self.error), np.array(self.rotation, np.float32).reshape(3, 3),
                np.array(self.translation, np.float32))


class Intrinsics(C.Structure):
    _fields_ = [("image_width", C.c_uint32), ("image_height", C.c_uint32), ("focal_x", C.c_float),
                ("focal_y", C.c_float), ("principal_x", C.c_float), ("principal_y", C.c_float)]


_pose_ready = False


def _pose_lib():
    global _pose_ready
    L = lib()
    if not _pose_ready:
        fp, PP = C.POINTER(C.c_float), C.POINTER(Pose)
        L.a3ref_pose_default.argtypes = [PP]
        L.a3ref_make_marker_square.argtypes = [C.c_float, fp]
        L.a3ref_homography_from_marker_square.argtypes = [C.c_float, fp, fp]
        L.a3ref_find_rotation_to_z.argtypes = [fp, fp]
        L.a3ref_compute_rotations.argtypes = [fp, C.c_float, C.c_float, fp, fp]
        L.a3ref_compute_translation.argtypes = [fp, fp, fp, fp]
        L.a3ref_reprojection_error.restype = C.c_float
        L.a3ref_reprojection_error.argtypes = [PP, fp, fp]
        L.a3ref_solve_canonical_form.argtypes = [fp, fp, fp, PP, PP]
        L.a3ref_solve_with_normalized_points.argtypes = [fp, C.c_float, PP, PP]
        L.a3ref_solve_with_undistorted_points.argtypes = [C.POINTER(C.c_uint32), C.c_float, C.c_uint32, C.c_uint32, PP, PP]
        L.a3ref_solve_with_intrinsics.argtypes = [C.POINTER(C.c_uint32), C.c_float, C.POINTER(Intrinsics), PP, PP]
        L.a3ref_pose_apply.argtypes = [PP, fp, C.c_size_t, C.c_int, fp]
        L.a3ref_intrinsics_new.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, fp, fp, C.POINTER(Intrinsics)]
        L.a3ref_intrinsics_from_fov_horizontal.argtypes = [C.c_float, C.c_float, C.c_uint32, C.c_uint32,
                                                           C.POINTER(Intrinsics)]
        L.a3ref_project.argtypes = [C.POINTER(Intrinsics), C.c_float, C.c_float, C.c_float, fp]
        L.a3ref_project_culled.argtypes = [C.POINTER(Intrinsics), C.c_float, C.c_float, C.c_float, fp]
        L.a3ref_unproject.argtypes = [C.POINTER(Intrinsics), C.c_float, C.c_float, fp]
        _pose_ready = True
    return L


def _f(values):
    a = np.ascontiguousarray(values, np.float32).ravel()
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def marker_square(size: float) -> np.ndarray:
    out, po = _f(np.zeros(12))
    _pose_lib().a3ref_make_marker_square(size, po)
    return out.reshape(4, 3)


def homography_from_marker_square(size: float, pts) -> np.ndarray:
    p, pp = _f(pts)
    out, po = _f(np.zeros(9))
    _pose_lib().a3ref_homography_from_marker_square(size, pp, po)
    return out.reshape(3, 3)


def solve_canonical_form(size: float, pts):
    sq, psq = _f(marker_square(size))
    p, pp = _f(pts)
    h, ph = _f(homography_from_marker_square(size, pts))
    a, b = Pose(), Pose()
    _pose_lib().a3ref_solve_canonical_form(psq, pp, ph, C.byref(a), C.byref(b))
    return a, b


def solve_with_normalized_points(pts, size: float):
    p, pp = _f(pts)
    a, b = Pose(), Pose()
    _pose_lib().a3ref_solve_with_normalized_points(pp, size, C.byref(a), C.byref(b))
    return a, b


def solve_with_undistorted_points(corners, size: float, image_size):
    c = np.ascontiguousarray(corners, np.uint32).ravel()
    a, b = Pose(), Pose()
    _pose_lib().a3ref_solve_with_undistorted_points(c.ctypes.data_as(C.POINTER(C.c_uint32)), size, image_size[0],
                                                    image_size[1], C.byref(a), C.byref(b))
    return a, b


def solve_with_intrinsics(corners, size: float, k: Intrinsics):
    c = np.ascontiguousarray(corners, np.uint32).ravel()
    a, b = Pose(), Pose()
    _pose_lib().a3ref_solve_with_intrinsics(c.ctypes.data_as(C.POINTER(C.c_uint32)), size, C.byref(k), C.byref(a),
                                            C.byref(b))
    return a, b


def pose_apply(pose: Pose, pts, inverse: bool = False) -> np.ndarray:
    p, pp = _f(pts)
    out, po = _f(np.zeros(p.size))
    _pose_lib().a3ref_pose_apply(C.byref(pose), pp, p.size // 3, int(inverse), po)
    return out.reshape(-1, 3)


def intrinsics_new(w, h, fx, fy, px=None, py=None) -> Intrinsics:
    k = Intrinsics()
    cx = C.byref(C.c_float(px)) if px is not None else None
    cy = C.byref(C.c_float(py)) if py is not None else None
    _pose_lib().a3ref_intrinsics_new(w, h, fx, fy, C.cast(cx, C.POINTER(C.c_float)) if cx else None,
                                     C.cast(cy, C.POINTER(C.c_float)) if cy else None, C.byref(k))
    return k


def intrinsics_from_fov_horizontal(hfov, sensor_w, rx, ry) -> Intrinsics:
    k = Intrinsics()
    _pose_lib().a3ref_intrinsics_from_fov_horizontal(hfov, sensor_w, rx, ry, C.byref(k))
    return k
