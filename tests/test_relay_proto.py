"""The relay scheme of K3 (k3_segments / k3_cycles in aruco3_b200/csrc/k3_contours.cu) as its pure-Python specification
(tools/relay_proto.py): relays = visits owning a west / east crack on every R-th row; one walker per relay, the relays of a
border linked into a cycle, the border's start = the smallest candidate key on the cycle.  On random masks every border that owns a
relay crack must come out with the length, the start and the point order of a plain trace from its raster-first candidate."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
import relay_proto  # noqa: E402


def test_relay_borders_equal_plain_traces():
    rng = np.random.default_rng(11)
    relay_borders = 0
    for t in range(120):
        h, w = int(rng.integers(3, 36)), int(rng.integers(3, 64))
        m = rng.random((h, w)) < rng.uniform(0.25, 0.75)
        if t % 3 == 0:  # thick shapes: long borders crossing several relay rows
            m = np.zeros((h, w), bool)
            for _ in range(4):
                x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
                m[y0:y0 + int(rng.integers(2, h)), x0:x0 + int(rng.integers(2, w))] ^= True
        for R, J in ((1, 2), (4, 1), (4, 3), (16, 32)):  # J: segments summed up per jump (k3_jumps)
            a, _ = relay_proto.check(m, R, J)
            relay_borders += a
    assert relay_borders > 1000


def test_relay_visit_owning_both_cracks_is_named_by_the_west_one():
    # a one-pixel-wide vertical line: every visit going down owns the west AND the east crack of its pixel
    m = np.zeros((9, 5), bool)
    m[1:8, 2] = True
    for R in (1, 2, 4):
        got, left = relay_proto.check(m, R)
        assert got == 1 and left == 0
