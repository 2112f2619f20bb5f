"""Known-answer tests of the reference, restated against the oracle.

Each test names the reference test it restates; these are the only results the reference's own
tests pin on the detection path.
"""
import ctypes as C

import numpy as np


def test_hamming_distance(oracle):
    """/root/reference/src/lib.rs:28-40"""
    h = oracle.lib().a3ref_hamming_distance
    for i in range(255):
        assert h(i, i) == 0
    assert h(0xFFFFFFFF, 0x0) == 32
    assert h(0x0, 0xFFFFFFFFFFFFFFFF) == 64
    assert h(0b10000000_00000000_00000000_00000000, 0b01000000_00000000_00000000_00000000) == 2


def test_tau_sanity(oracle):
    """/root/reference/src/dictionaries.rs:239-243"""
    assert oracle.dictionary("ARUCO_DEFAULT").tau == 3


def test_find_nearest_aruco_default(oracle):
    """/root/reference/src/dictionaries.rs:245-269"""
    d = oracle.dictionary("ARUCO_DEFAULT")
    assert oracle.find_nearest(d, 0x1084210) == (0, 0)
    assert oracle.find_nearest(d, 0x1084209) == (2, 0)
    assert oracle.find_nearest(d, 0b00000001_00001000_01000010_00001001) == (2, 0)
    assert oracle.find_nearest(d, 0b00000001_00001000_01000010_10001001) == (2, 1)
    assert oracle.find_nearest(d, 0x1084217) == (1, 0)


def test_try_find_nearest_aruco_default(oracle):
    """/root/reference/src/dictionaries.rs:271-281 (the second literal has a 7-digit group: 31 bits)"""
    d = oracle.dictionary("ARUCO_DEFAULT")
    m = oracle.try_find_nearest(d, 0b01100001_00001000_01000010_00001001)
    assert m is not None and m[0] == 2
    assert oracle.try_find_nearest(d, 0b11111111_0000100_01000010_00001001) is None


def test_enforce_clockwise(oracle):
    """/root/reference/src/aruco.rs:400-412"""
    q = np.array([[0, 0, 0, 1, 1, 1, 1, 0], [0, 0, 1, 0, 1, 1, 0, 1]], dtype=np.uint32)
    oracle.lib().a3ref_enforce_clockwise_corners(q.ctypes.data, 2)
    assert (q[0] == q[1]).all()


def _rot(oracle, m):
    m = np.ascontiguousarray(m, dtype=np.uint8)
    out = np.empty((m.shape[1], m.shape[0]), np.uint8)
    oracle.lib().a3ref_rotate_bit_matrix(m.ctypes.data, m.shape[0], m.shape[1], out.ctypes.data)
    return out.tolist()


def test_bit_rotate(oracle):
    """/root/reference/src/aruco.rs:414-444"""
    assert _rot(oracle, [[1, 1, 1], [1, 0, 0], [0, 1, 0]]) == [[1, 0, 0], [1, 0, 1], [1, 1, 0]]
    assert _rot(oracle, [[1, 1, 1, 1], [1, 1, 1, 0], [1, 1, 0, 0], [1, 0, 0, 0]]) == \
        [[1, 0, 0, 0], [1, 1, 0, 0], [1, 1, 1, 0], [1, 1, 1, 1]]


def test_drop_too_near(oracle):
    """/root/reference/src/aruco.rs:446-459"""
    q = np.array([[0, 0, 10, 0, 10, 10, 0, 10], [1, 0, 10, 0, 10, 10, 0, 10], [0, 0, 10, 2, 10, 10, 0, 10],
                  [0, 0, 10, 0, 10, 10, 3, 10]], dtype=np.uint32)
    assert oracle.lib().a3ref_discard_too_near(q.ctypes.data, 4, 10.0) == 1


def test_dictionary_inventory(oracle):
    """SURVEY.md §8a-6 (counts extracted from /root/reference/src/dictionaries.rs:5-19, map :30-113)."""
    expect = {"ARUCO": (1023, 25, 7, 3), "ARUCO_DEFAULT": (1023, 25, 7, 3), "ARUCO_MIP_16H3": (250, 16, 6, 3),
              "ARUCO_MIP_25H7": (100, 25, 7, 7), "ARUCO_MIP_36H12": (250, 36, 8, 12), "APRILTAG_16H5": (30, 16, 6, 5),
              "APRILTAG_25H7": (242, 25, 7, 7), "APRILTAG_25H9": (35, 25, 7, 9), "APRILTAG_36H9": (5329, 36, 8, 9),
              "APRILTAG_36H10": (2320, 36, 8, 10), "APRILTAG_36H11": (587, 36, 8, 11), "ARTAG": (1024, 36, 8, 0),
              "ARTOOLKITPLUS": (512, 36, 8, 4), "ARTOOLKITPLUSBCH": (4096, 36, 8, 9), "CHILITAGS": (1024, 64, 10, 5)}
    L = oracle.lib()
    names = {L.a3ref_dictionary_name(i).decode() for i in range(L.a3ref_dictionary_count())}
    assert names == set(expect)
    for name, (n, bits, ms, tau) in expect.items():
        d = oracle.dictionary(name.lower())  # new_from_named_dict upper-cases (dictionaries.rs:141)
        assert (d.n_codes, d.num_bits, L.a3ref_mark_size(C.byref(d)), d.tau) == (n, bits, ms, tau), name


def test_python_tables_match_oracle(oracle):
    from aruco3_b200 import dictionaries
    for name in dictionaries.dictionary_names():
        t, d = dictionaries.table(name), oracle.dictionary(name)
        assert (t.num_bits, t.mark_size, len(t.codes), t.tau) == (d.num_bits, oracle.lib().a3ref_mark_size(C.byref(d)), d.n_codes, d.tau)
        assert (np.ctypeslib.as_array(d.codes, (d.n_codes,)) == t.codes).all()


def test_aruco_table_is_opencv_original():
    """Independent anchor: `ARUCO` == OpenCV DICT_ARUCO_ORIGINAL, row-major, first interior cell = MSB."""
    cv2 = __import__("pytest").importorskip("cv2")
    from aruco3_b200 import dictionaries
    t = dictionaries.table("ARUCO")
    d = cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_ARUCO_ORIGINAL)
    for i in (0, 1, 2, 77, 500, 1022):
        img = cv2.aruco.generateImageMarker(d, i, 7, borderBits=1)
        bits = (img[1:6, 1:6] > 127).astype(np.uint64).ravel()
        code = 0
        for b in bits:
            code = (code << 1) | int(b)
        assert code == int(t.codes[i])


def test_make_binary_image_quirk_q8(oracle):
    """dictionaries.rs:212-232 writes LSB first while the decoder reads MSB first."""
    d = oracle.dictionary("ARUCO")
    buf = np.zeros(128, np.uint8)
    w = oracle.lib().a3ref_make_binary_image(C.byref(d), 5, buf.ctypes.data, 128)
    assert w == 7
    grid = buf[:49].reshape(7, 7)
    assert not grid[0].any() and not grid[-1].any() and not grid[:, 0].any() and not grid[:, -1].any()
    code = int(np.ctypeslib.as_array(d.codes, (d.n_codes,))[5])
    inner = grid[1:6, 1:6].ravel()
    assert [int(b) for b in inner] == [(code >> i) & 1 for i in range(25)]
