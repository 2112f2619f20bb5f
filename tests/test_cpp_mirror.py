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
