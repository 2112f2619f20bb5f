import ctypes as C, os, sys, time
sys.path.insert(0, "/root/repo")
os.environ["A3_TRACE"] = "1"
import numpy as np, torch
from aruco3_b200 import Detector, _ffi, synth
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
