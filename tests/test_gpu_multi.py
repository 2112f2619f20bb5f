"""Frame-batch sharding on REAL distinct GPUs (the driver's own GPU test box has one GPU, where these skip —
run with `gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py -m gpu`): every device returns the oracle's answer for
its block of frames, in-process (`ShardedDetector(devices=range(N))`, C++ `aruco3::ShardedDetector`) and the results do
not depend on which device a frame lands on."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def ndev():
    from aruco3_b200 import _ffi
    n = _ffi.lib().a3_device_count()
    if n < 1:
        pytest.fail("no CUDA device visible: the gpu tests have no fallback")
    if n < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus N)")
    return n


def _markers(dets):
    return [[(m.candidate, m.id, m.rotation, m.hamming_distance, m.code, m.corners) for m in x.markers] for x in dets]


def test_sharded_detector_on_distinct_gpus_matches_oracle(ndev, oracle):
    import aruco3_b200 as a3
    from aruco3_b200 import synth
    from aruco3_b200.sharding import ShardedDetector, shard_range
    n = 5 * ndev + 3  # uneven blocks
    frames, _ = synth.render_batch("C1", n)
    want = []
    for f in range(n):
        ref = oracle.detect(frames[f], "ARUCO")
        want.append([(m["candidate"], m["id"], m["rotation"], m["hamming_distance"], m["code"],
                      [(m["corners"][2 * k], m["corners"][2 * k + 1]) for k in range(4)]) for m in ref.markers])
    assert sum(len(x) for x in want) > 4 * n
    with ShardedDetector(devices=tuple(range(ndev))) as sd:
        assert [s.device for s in sd.shards] == list(range(ndev))
        for _ in range(3):  # first call, then each shard's one-shot route
            assert _markers(sd.detect_batch(frames)) == want
        full = sd.detect_batch(frames, full=True)
    for f in range(n):
        ref = oracle.detect(frames[f], "ARUCO")
        assert np.array_equal(full[f].grey, ref.grey) and [list(sum(c, ())) for c in full[f].candidates] == ref.candidates.tolist()
    # every single device gives the same answer for the whole batch
    for dev in range(ndev):
        with a3.Detector(device=dev) as d:
            assert _markers(d.detect_batch(frames)) == want, f"device {dev}"
    assert shard_range(n, ndev - 1, ndev)[1] == n


def test_4k_batch_sharded_over_all_gpus(ndev, oracle):
    """BASELINE.json configs[3] in small: a 4K batch cut into contiguous blocks over every GPU of the box."""
    from aruco3_b200 import synth
    from aruco3_b200.sharding import ShardedDetector
    n = 2 * ndev + 1
    frames, _ = synth.render_batch("C4", n)
    with ShardedDetector(devices=tuple(range(ndev))) as sd:
        got = _markers(sd.detect_batch(frames))
    for f in range(n):
        ref = oracle.detect(frames[f], "ARUCO")
        assert [m[1:] for m in got[f]] == [(m["id"], m["rotation"], m["hamming_distance"], m["code"],
                                            [(m["corners"][2 * k], m["corners"][2 * k + 1]) for k in range(4)]) for m in ref.markers], f
        assert len(got[f]) >= 15


SHARDED = r'''
#include <cstdio>
#include <vector>
#include "aruco3_b200.hpp"
int main(int argc, char **argv) {
    const uint32_t w = 640, h = 480, n = (uint32_t)atoi(argv[2]);
    const int ndev = a3_device_count();
    std::vector<uint8_t> rgb((size_t)n * w * h * 3);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(rgb.data(), 1, rgb.size(), f) != rgb.size()) return 2;
    fclose(f);
    const aruco3::ARDictionary dict = aruco3::ARDictionary::new_from_named_dict("ARUCO");
    std::vector<int> devices;
    for (int d = 0; d < ndev; d++) devices.push_back(d);
    aruco3::ShardedDetector many(aruco3::DetectorConfig(), dict, devices);
    for (int round = 0; round < 2; round++) {
        const auto b = many.detect_batch(rgb.data(), n, w, h);
        if (b.size() != n) return 4;
        for (uint32_t i = 0; i < n; i++)
            for (auto &m : b[i].markers)
                printf("%d %u %zu %llu %u %u %u %u %u %u %u %u %u\n", round, i, m.id, (unsigned long long)m.code, m.hamming_distance, m.corners[0].first,
                       m.corners[0].second, m.corners[1].first, m.corners[1].second, m.corners[2].first, m.corners[2].second, m.corners[3].first, m.corners[3].second);
    }
    printf("devices %d\n", ndev);
    return 0;
}
'''


def test_cpp_sharded_detector_on_distinct_gpus(ndev, oracle, tmp_path):
    from aruco3_b200 import synth
    n = 3 * ndev + 1
    frames, _ = synth.render_batch("C1", n)
    (tmp_path / "frames.rgb").write_bytes(frames.tobytes())
    src, exe = tmp_path / "sharded.cpp", tmp_path / "sharded"
    src.write_text(SHARDED)
    lib_dir = ROOT / "aruco3_b200"
    subprocess.run(["g++", "-std=c++17", "-pthread", "-I", str(ROOT / "include"), str(src), "-o", str(exe), f"-L{lib_dir}", "-laruco3_b200",
                    f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe), str(tmp_path / "frames.rgb"), str(n)], check=True, capture_output=True, text=True).stdout.splitlines()
    assert out[-1] == f"devices {ndev}"
    want = []
    for f in range(n):
        for m in oracle.detect(frames[f], "ARUCO").markers:
            want.append([f, m["id"], m["code"], m["hamming_distance"]] + m["corners"])
    rows = [[int(v) for v in ln.split()] for ln in out[:-1]]
    assert [r[1:] for r in rows if r[0] == 0] == want and [r[1:] for r in rows if r[0] == 1] == want
