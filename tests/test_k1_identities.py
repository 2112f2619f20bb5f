"""Integer identities the pixel kernels rely on, checked exhaustively on the CPU (numpy).  They are what makes the
CUDA arithmetic bit-exact with `(2126 R + 7152 G + 722 B) / 10000` and `pix >= floor(S / cnt)`."""
import numpy as np


def test_luma_dp4a_split_and_magic_division():
    """k1_strips.cu: v = dp4a(px, [78,240,210,0]) + 256 * dp4a(px, [8,27,2,0]);  grey = umulhi(v, ceil(2^40/1e4)) >> 8."""
    assert (8 * 256 + 78, 27 * 256 + 240, 2 * 256 + 210) == (2126, 7152, 722)
    m = -(-(1 << 40) // 10000)
    assert m == 109951163
    v = np.arange(0, 2126 * 255 + 7152 * 255 + 722 * 255 + 1, dtype=np.uint64)
    hi = (v * np.uint64(m)) >> np.uint64(32)
    assert (hi >> np.uint64(8) == v // np.uint64(10000)).all()
    assert int(hi.max()) < 65536  # bytes 2 and 3 of the high word are zero: the PRMT packing relies on it


def test_luma_fixed_point_24bit_all_rgb():
    """k1_threshold.cu (generic kernel): grey = (kWr R + kWg G + kWb B + kBias) >> 24 for all 2^24 colours."""
    kwr, kwg, kwb, kbias = 3566836, 11999065, 1211315, 1678
    assert kwr + kwg + kwb == 1 << 24
    g, b = np.meshgrid(np.arange(256, dtype=np.uint64), np.arange(256, dtype=np.uint64), indexing="ij")
    for r in range(256):
        want = (2126 * r + 7152 * g + 722 * b) // 10000
        got = (kwr * r + kwg * g + kwb * b + kbias) >> np.uint64(24)
        assert int((kwr * 255 + kwg * 255 + kwb * 255 + kbias)) < 1 << 32
        assert (want == got).all(), r


def test_threshold_without_division():
    """pix >= S // cnt  <=>  S < (pix + 1) * cnt  <=>  S - 256 cnt + (255 - pix) cnt < 0, for every window area and pixel."""
    rng = np.random.default_rng(1)
    for cnt in range(64, 226):
        s = rng.integers(0, 255 * cnt + 1, size=4000)
        pix = rng.integers(0, 256, size=4000)
        a = pix >= s // cnt
        b = s < (pix + 1) * cnt
        c = (s - 256 * cnt + (255 - pix) * cnt) < 0
        assert (a == b).all() and (a == c).all()
    # the packed u16 column sums never overflow: 15 rows of 255, and 15 x 15 windows fit 16 bits
    assert 15 * 255 < 1 << 16 and 225 * 255 < 1 << 16


def test_mask_nibble_expansion():
    """expand4: 4 mask bits -> 4 bytes of 0 / 255."""
    for nib in range(16):
        v = (((nib * 0x00204081) & 0xFFFFFFFF) & 0x01010101) * 0xFF & 0xFFFFFFFF
        assert [(v >> (8 * j)) & 0xFF for j in range(4)] == [255 * ((nib >> j) & 1) for j in range(4)]
