#!/usr/bin/env python3
"""Host timestamps and device spans of the phases of a host-input a3_detect_batch call over 256 x 1080p frames (A3_TRACE=1:
the library prints them on stderr): where the time after the last byte of the H2D stream goes.
usage: python tools/trace_e2e.py"""
import ctypes as C
import os
import sys
import time
from pathlib import Path

os.environ["A3_TRACE"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from aruco3_b200 import Detector, _ffi, synth  # noqa: E402

n, h, w = 256, 1080, 1920
pinned = torch.empty((n, h, w, 3), dtype=torch.uint8, pin_memory=True)
synth.render_batch("C3", n, 0, out=pinned.numpy())
with Detector(dictionary="ARUCO") as det:
    markers = (_ffi.A3Marker * (64 * n))()
    nm, st = C.c_uint32(), _ffi.A3Stats()
    for it in range(5):
        t0 = time.perf_counter()
        _ffi.check(_ffi.lib().a3_detect_batch(det._h, pinned.data_ptr(), _ffi.FMT_RGB8, _ffi.MEM_HOST, n, w, h, w * 3, w * h * 3,
                                              C.cast(markers, C.c_void_p), 64 * n, C.byref(nm), None, C.byref(st)))
        print(f"wall {1e3 * (time.perf_counter() - t0):.3f} ms total {st.ms_total:.3f} h2d {st.ms_h2d:.3f}", file=sys.stderr, flush=True)
