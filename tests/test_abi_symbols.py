"""The C-ABI library loads on a machine without a GPU and exports every symbol include/aruco3_b200.h declares;
struct layouts of the ctypes binding match the header (no compute calls here)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "aruco3_b200.h"


def _declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(a3_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    from aruco3_b200 import _ffi
    L = _ffi.lib()
    names = _declared_functions()
    assert len(names) >= 18, names
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/aruco3_b200.h but not exported"


def test_no_unexpected_cuda_requirement_at_load():
    """Loading and the host-only entry points work without a device; compute entry points fail loudly instead of falling back."""
    from aruco3_b200 import _ffi
    L = _ffi.lib()
    assert b"sm_100a" in L.a3_version()
    cfg = _ffi.A3Config()
    L.a3_config_default(C.byref(cfg))
    assert (cfg.threshold_window, cfg.homography_sample_size, cfg.filter_high_bit_errors) == (7, 49, 1)
    assert abs(cfg.contour_simplification_epsilon - 0.05) < 1e-12
    if L.a3_device_count() == 0:
        d = _ffi.A3Dictionary()
        assert L.a3_dictionary_by_name(b"ARUCO", C.byref(d)) == _ffi.A3_OK
        h = C.c_void_p()
        st = L.a3_detector_create(C.byref(cfg), C.byref(d), 0, C.byref(h))
        assert st == _ffi.A3_ERR_CUDA and b"no CPU fallback" in L.a3_last_error()


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof as the C compiler sees them == the ctypes mirrors."""
    from aruco3_b200 import _ffi
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "aruco3_b200.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(a3_config), sizeof(a3_dictionary), sizeof(a3_marker), sizeof(a3_decode),'
                   ' sizeof(a3_stats), sizeof(a3_outputs), sizeof(a3_k1_tuning), sizeof(a3_pose), sizeof(a3_camera_intrinsics));'
                   'printf("%zu %zu %zu %zu\\n", offsetof(a3_marker, frame), offsetof(a3_decode, has_codes), offsetof(a3_stats, ms_h2d),'
                   ' offsetof(a3_outputs, frame_marker_offsets)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(v) for v in out]
    want = [C.sizeof(t) for t in (_ffi.A3Config, _ffi.A3Dictionary, _ffi.A3Marker, _ffi.A3Decode, _ffi.A3Stats, _ffi.A3Outputs, _ffi.A3K1Tuning,
                                  _ffi.A3Pose, _ffi.A3CameraIntrinsics)]
    want += [_ffi.A3Marker.frame.offset, _ffi.A3Decode.has_codes.offset, _ffi.A3Stats.ms_h2d.offset, _ffi.A3Outputs.frame_marker_offsets.offset]
    assert sizes == want


def test_dictionary_functions_match_reference_kats():
    """The product's own dictionary entry points against the reference's unit tests
    (/root/reference/src/lib.rs:28-40, /root/reference/src/dictionaries.rs:239-281)."""
    import aruco3_b200 as a3
    assert a3.hamming_distance(0xFFFFFFFF, 0) == 32 and a3.hamming_distance(0, 0xFFFFFFFFFFFFFFFF) == 64
    assert all(a3.hamming_distance(i, i) == 0 for i in range(255))
    d = a3.ARDictionary.new_from_named_dict("aruco_default")
    assert d.tau == 3 and d.get_mark_size() == 7
    assert d.find_nearest(0x1084210) == (0, 0) and d.find_nearest(0x1084209) == (2, 0) and d.find_nearest(0x1084217) == (1, 0)
    assert d.find_nearest(0b00000001_00001000_01000010_10001001) == (2, 1)
    assert d.try_find_nearest(0b01100001_00001000_01000010_00001001)[0] == 2
    assert d.try_find_nearest(0b11111111_0000100_01000010_00001001) is None
    assert set(a3.ARDictionary.get_dictionary_names()) >= {"ARUCO", "APRILTAG_36H11", "CHILITAGS", "ARTAG"}
    with pytest.raises(a3.A3Error):
        a3.ARDictionary.new_from_named_dict("nope")
    width, bits = d.make_binary_image(5)  # (width, bits) as src/dictionaries.rs:212
    assert width == 7 and len(bits) == 49 and not any(bits[:7])
