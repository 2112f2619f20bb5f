#!/usr/bin/env python3
"""Single-frame latency of a3_detect_batch on the two BASELINE.json configs[1] workloads (a marker frame and the reference
bench's uniform-noise frame, 1920x1080 RGB from pinned host memory), with the library's per-phase K3 times
(A3_K3_TIMING=1) and stage statistics.  usage: python tools/latency_probe.py [noise|marker]"""
import ctypes as C
import os
import sys
import time
from pathlib import Path

os.environ["A3_K3_TIMING"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aruco3_b200 import Detector, _ffi, synth  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "noise"
    h, w = 1080, 1920
    frame = torch.empty((1, h, w, 3), dtype=torch.uint8, pin_memory=True)
    if which == "noise":
        frame.numpy()[:] = np.random.default_rng(0xA3C0DE00 + 2000).integers(0, 256, size=(1, h, w, 3), dtype=np.uint8)
    else:
        frame.numpy()[:] = synth.render_batch("C3", 1)[0]
    with Detector(dictionary="ARUCO") as det:
        markers = (_ffi.A3Marker * 4096)()
        nm, st = C.c_uint32(), _ffi.A3Stats()
        for it in range(6):
            print(f"--- call {it}", file=sys.stderr, flush=True)
            t0 = time.perf_counter()
            _ffi.check(_ffi.lib().a3_detect_batch(det._h, frame.data_ptr(), _ffi.FMT_RGB8, _ffi.MEM_HOST, 1, w, h, w * 3, w * h * 3,
                                                  C.cast(markers, C.c_void_p), 4096, C.byref(nm), None, C.byref(st)))
            print(f"wall {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr, flush=True)
        print({k: round(v, 3) if isinstance(v, float) else v for k, v in st.as_dict().items()})


if __name__ == "__main__":
    main()
