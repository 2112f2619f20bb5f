"""include/aruco3_b200.hpp — the C++ host mirror of `Detector { config, dictionary }.detect(img)` — compiled with g++,
run on the GPU and compared with the oracle (markers, candidates, grey) on a seeded frame."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

PROGRAM = r'''
#include <cstdio>
#include <vector>
#include "aruco3_b200.hpp"
int main(int argc, char **argv) {
    const uint32_t w = 640, h = 480;
    std::vector<uint8_t> rgb((size_t)w * h * 3);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(rgb.data(), 1, rgb.size(), f) != rgb.size()) return 2;
    fclose(f);
    try {
        aruco3::ARDictionary::new_from_named_dict("no such dictionary");
        return 3;
    } catch (const aruco3::Error &e) {
        if (e.status != A3_ERR_UNKNOWN_DICTIONARY) return 4;
    }
    aruco3::Detector detector(aruco3::DetectorConfig(), aruco3::ARDictionary::new_from_named_dict("ARUCO"));
    aruco3::Detection d = detector.detect(rgb.data(), w, h);
    unsigned long long sum = 0;
    for (uint8_t v : d.grey.data) sum += v;
    printf("grey %u %u %llu\n", d.grey.width, d.grey.height, sum);
    for (auto &c : d.candidates) printf("cand %u %u %u %u %u %u %u %u\n", c[0].first, c[0].second, c[1].first, c[1].second, c[2].first, c[2].second, c[3].first, c[3].second);
    for (auto &p : d.homographies) printf("patch %u\n", p.width);
    for (auto &m : d.markers) {  // examples/webcam_kamera.rs:68
        auto poses = aruco3::pose::solve_with_undistorted_points(detector, m.corners, 40.0f, {w, h});
        printf("pose %.9g %.9g %.9g %.9g %.9g\n", poses.first.error, poses.first.translation[0], poses.first.translation[1],
               poses.first.translation[2], poses.second.translation[2]);
    }
    aruco3::MarkerPose mp;
    mp.translation = {1.0f, 2.0f, 3.0f};
    mp.rotation = {0, 0, 1, 0, 1, 0, 1, 0, 0};
    auto moved = mp.apply_transform_to_points({{7.0f, 11.0f, 13.0f}});  // src/pose.rs:379-392
    if (moved[0][0] != 14.0f || moved[0][1] != 13.0f || moved[0][2] != 10.0f) return 5;
    if (aruco3::CameraIntrinsics(640, 480, 1.0f, 1.0f).principal_x != 320.0f) return 6;
    for (auto &m : d.markers)
        printf("marker %zu %llu %u %u %u %u %u %u %u %u %u\n", m.id, (unsigned long long)m.code, m.hamming_distance, m.corners[0].first, m.corners[0].second,
               m.corners[1].first, m.corners[1].second, m.corners[2].first, m.corners[2].second, m.corners[3].first, m.corners[3].second);
    return 0;
}
'''


def test_header_compiles_without_gpu(tmp_path):
    src = tmp_path / "mirror.cpp"
    src.write_text(PROGRAM)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / "mirror.o")], check=True)


@pytest.mark.gpu
def test_cpp_detector_matches_oracle(oracle, tmp_path):
    from aruco3_b200 import _ffi, synth
    _ffi.lib()
    img, _ = synth.render_frame(synth.CONFIGS["C1"], 0)
    (tmp_path / "frame.rgb").write_bytes(img.tobytes())
    src, exe = tmp_path / "mirror.cpp", tmp_path / "mirror"
    src.write_text(PROGRAM)
    lib_dir = ROOT / "aruco3_b200"
    subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "include"), str(src), "-o", str(exe), f"-L{lib_dir}", "-laruco3_b200",
                    f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe), str(tmp_path / "frame.rgb")], check=True, capture_output=True, text=True).stdout.splitlines()
    ref = oracle.detect(img, "ARUCO")
    assert out[0] == f"grey 640 480 {int(ref.grey.astype(np.uint64).sum())}"
    cands = [[int(v) for v in ln.split()[1:]] for ln in out if ln.startswith("cand")]
    assert cands == ref.candidates.tolist()
    assert [int(ln.split()[1]) for ln in out if ln.startswith("patch")] == [49 if ok else 1 for ok in ref.homography_ok]
    markers = [[int(v) for v in ln.split()[1:]] for ln in out if ln.startswith("marker")]
    assert markers == [[m["id"], m["code"], m["hamming_distance"]] + m["corners"] for m in ref.markers]
    poses = [[np.float32(v) for v in ln.split()[1:]] for ln in out if ln.startswith("pose")]
    assert len(poses) == len(ref.markers) > 0
    for got, m in zip(poses, ref.markers):
        best, alt = oracle.solve_with_undistorted_points(m["corners"], 40.0, (640, 480))
        assert got == [np.float32(best.error)] + [np.float32(v) for v in best.translation] + [np.float32(alt.translation[2])]


SHARDED = r'''
#include <cstdio>
#include <vector>
#include "aruco3_b200.hpp"
int main(int argc, char **argv) {
    const uint32_t w = 640, h = 480, n = 7;
    std::vector<uint8_t> rgb((size_t)n * w * h * 3);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(rgb.data(), 1, rgb.size(), f) != rgb.size()) return 2;
    fclose(f);
    if (aruco3::ShardedDetector::shard_range(7, 0, 3) != std::make_pair(0u, 3u) || aruco3::ShardedDetector::shard_range(7, 2, 3) != std::make_pair(5u, 7u)) return 3;
    const aruco3::ARDictionary dict = aruco3::ARDictionary::new_from_named_dict("ARUCO");
    aruco3::Detector one(aruco3::DetectorConfig(), dict);
    aruco3::ShardedDetector many(aruco3::DetectorConfig(), dict, {0, 0, 0});  // three shards (on the one device a test box has)
    for (int round = 0; round < 2; round++) {
        const auto a = one.detect_batch(rgb.data(), n, w, h), b = many.detect_batch(rgb.data(), n, w, h);
        if (a.size() != n || b.size() != n) return 4;
        for (uint32_t i = 0; i < n; i++) {
            if (a[i].markers.size() != b[i].markers.size()) return 5;
            for (size_t k = 0; k < a[i].markers.size(); k++)
                if (a[i].markers[k].id != b[i].markers[k].id || a[i].markers[k].corners != b[i].markers[k].corners || a[i].markers[k].code != b[i].markers[k].code) return 6;
            printf("frame %u markers %zu\n", i, b[i].markers.size());
        }
    }
    return 0;
}
'''


@pytest.mark.gpu
def test_cpp_sharded_detector_equals_single(tmp_path):
    """aruco3::ShardedDetector (contiguous frame blocks, one Detector and host thread per shard, no collective) returns what
    one Detector returns for the whole batch, frame for frame."""
    from aruco3_b200 import _ffi, synth
    _ffi.lib()
    frames, _ = synth.render_batch("C1", 7)
    (tmp_path / "frames.rgb").write_bytes(frames.tobytes())
    src, exe = tmp_path / "sharded.cpp", tmp_path / "sharded"
    src.write_text(SHARDED)
    lib_dir = ROOT / "aruco3_b200"
    subprocess.run(["g++", "-std=c++17", "-pthread", "-I", str(ROOT / "include"), str(src), "-o", str(exe), f"-L{lib_dir}", "-laruco3_b200",
                    f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe), str(tmp_path / "frames.rgb")], check=True, capture_output=True, text=True).stdout.splitlines()
    assert len(out) == 14 and sum(int(ln.split()[-1]) for ln in out[:7]) > 20


def test_sharded_header_compiles_without_gpu(tmp_path):
    src = tmp_path / "sharded.cpp"
    src.write_text(SHARDED)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-pthread", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / "sharded.o")], check=True)


def test_cpp_shard_range_equals_python(tmp_path):
    """aruco3::ShardedDetector::shard_range is the rule of aruco3_b200.sharding.shard_range (no GPU needed: the function is
    static and the program never creates a detector)."""
    from aruco3_b200.sharding import shard_range
    src, exe = tmp_path / "ranges.cpp", tmp_path / "ranges"
    src.write_text(r'''
#include <cstdio>
#include "aruco3_b200.hpp"
int main() {
    for (unsigned n = 0; n <= 40; n++)
        for (unsigned w = 1; w <= 9; w++)
            for (unsigned r = 0; r < w; r++) {
                auto p = aruco3::ShardedDetector::shard_range(n, r, w);
                printf("%u %u %u %u %u\n", n, w, r, p.first, p.second);
            }
    return 0;
}
''')
    lib_dir = ROOT / "aruco3_b200"
    subprocess.run(["g++", "-std=c++17", "-pthread", "-I", str(ROOT / "include"), str(src), "-o", str(exe), f"-L{lib_dir}", "-laruco3_b200",
                    f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    vals = list(map(int, out))
    assert len(vals) == 5 * 41 * 45
    for i in range(0, len(vals), 5):
        n, w, r, lo, hi = vals[i:i + 5]
        assert (lo, hi) == shard_range(n, r, w), (n, w, r)


PLAIN = r'''
#include <cstdio>
#include <vector>
#include "aruco3_b200.hpp"
// The reference's usage pattern (benches/detect_markers.rs:17-25, examples/webcam_kamera.rs:14-17, 56): a Detector built
// with a struct literal, `detect` once per frame.  100 frames through one PlainDetector, a second literal with the same
// fields in between, pageable std::vector input: exactly one a3_detector_create, and from the second call on the one-shot
// route (the handle that comes back from the cache is warm).
int main(int argc, char **argv) {
    const uint32_t w = 640, h = 480;
    std::vector<uint8_t> rgb((size_t)w * h * 3);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(rgb.data(), 1, rgb.size(), f) != rgb.size()) return 2;
    fclose(f);
    const unsigned long long before = a3_detector_create_count();
    const aruco3::PlainDetector detector{aruco3::DetectorConfig(), aruco3::ARDictionary::new_from_named_dict("ARUCO")};
    size_t markers0 = 0;
    unsigned warm = 0;
    for (int i = 0; i < 100; i++) {
        a3_stats st{};
        const aruco3::PlainDetector again{aruco3::DetectorConfig(), aruco3::ARDictionary::new_from_named_dict("aruco")};
        const auto dets = (i % 10 == 9 ? again : detector).detect_batch(rgb.data(), 1, w, h, aruco3::PixelFormat::Rgb8, /*full=*/i % 2 == 0, &st);
        if (i == 0) markers0 = dets[0].markers.size();
        if (dets[0].markers.size() != markers0) return 3;
        warm += st.one_shot;
    }
    aruco3::DetectorConfig other;
    other.min_corner_separation_factor = 0.05f;  // a different config is a different handle
    const aruco3::PlainDetector third{other, aruco3::ARDictionary::new_from_named_dict("ARUCO")};
    third.detect(rgb.data(), w, h);
    printf("created %llu warm %u markers %zu\n", a3_detector_create_count() - before, warm, markers0);
    a3_detector_cache_clear();
    return 0;
}
'''


def test_plain_header_compiles_without_gpu(tmp_path):
    src = tmp_path / "plain.cpp"
    src.write_text(PLAIN)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / "plain.o")], check=True)


@pytest.mark.gpu
def test_cpp_plain_detector_creates_one_handle_for_100_detects(oracle, tmp_path):
    from aruco3_b200 import _ffi, synth
    _ffi.lib()
    img, _ = synth.render_frame(synth.CONFIGS["C1"], 3)
    (tmp_path / "frame.rgb").write_bytes(img.tobytes())
    src, exe = tmp_path / "plain.cpp", tmp_path / "plain"
    src.write_text(PLAIN)
    lib_dir = ROOT / "aruco3_b200"
    subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "include"), str(src), "-o", str(exe), f"-L{lib_dir}", "-laruco3_b200",
                    f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe), str(tmp_path / "frame.rgb")], check=True, capture_output=True, text=True).stdout.split()
    ref = oracle.detect(img, "ARUCO")
    assert out == ["created", "2", "warm", "99", "markers", str(len(ref.markers))], out


@pytest.mark.gpu
def test_handle_cache_through_ctypes():
    """a3_detector_acquire / a3_detector_release: same key -> same handle back, different key or a handle still leased -> a new one."""
    import ctypes as C
    from aruco3_b200 import _ffi
    L = _ffi.lib()
    L.a3_detector_cache_clear()
    cfg, d = _ffi.A3Config(), _ffi.A3Dictionary()
    L.a3_config_default(C.byref(cfg))
    _ffi.check(L.a3_dictionary_by_name(b"ARUCO", C.byref(d)))
    n0 = L.a3_detector_create_count()
    h1, h2, h3 = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _ffi.check(L.a3_detector_acquire(C.byref(cfg), C.byref(d), 0, C.byref(h1)))
    _ffi.check(L.a3_detector_acquire(C.byref(cfg), C.byref(d), 0, C.byref(h2)))  # h1 is leased: a second handle
    assert h1.value != h2.value and L.a3_detector_create_count() == n0 + 2
    L.a3_detector_release(h1)
    _ffi.check(L.a3_detector_acquire(C.byref(cfg), C.byref(d), 0, C.byref(h3)))
    assert h3.value == h1.value and L.a3_detector_create_count() == n0 + 2
    cfg.threshold_window = 5
    h4 = C.c_void_p()
    _ffi.check(L.a3_detector_acquire(C.byref(cfg), C.byref(d), 0, C.byref(h4)))
    assert h4.value not in (h1.value, h2.value) and L.a3_detector_create_count() == n0 + 3
    for h in (h2, h3, h4):
        L.a3_detector_release(h)
    L.a3_detector_cache_clear()
