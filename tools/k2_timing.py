#!/usr/bin/env python3
"""Decode-kernel (K2) time on the resident 256 x 1080p headline workload, from a3_stats (CUDA events around the launch).
usage: python tools/k2_timing.py [batch]"""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from aruco3_b200 import Detector, _ffi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base, _ = synth.render_batch("C3", min(n, 32))
frames = torch.from_numpy(base).repeat((n + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:n].contiguous().cuda()
h, w = frames.shape[1:3]
with Detector(dictionary="ARUCO") as det:
    markers = (_ffi.A3Marker * (64 * n))()
    nm, st = C.c_uint32(), _ffi.A3Stats()
    t = []
    for it in range(8):
        _ffi.check(_ffi.lib().a3_detect_batch(det._h, frames.data_ptr(), _ffi.FMT_RGB8, _ffi.MEM_DEVICE, n, w, h, w * 3, w * h * 3,
                                              C.cast(markers, C.c_void_p), 64 * n, C.byref(nm), None, C.byref(st)))
        t.append((st.ms_decode_kernel, st.ms_contour_kernels, st.ms_pixel_kernel, st.ms_total))
    print("decode / contour / pixel / total ms (last 5 calls):", [tuple(round(v, 3) for v in x) for x in t[-5:]], "candidates", st.n_candidates, "markers", nm.value)
