#!/usr/bin/env python3
"""Phase times of the device contour stage (K3) on the resident 256 x 1080p headline workload: runs a3_detect_batch with
A3_K3_TIMING=1 (the library prints CUDA-event times per phase on stderr).  usage: python tools/k3_timing.py [batch]"""
import ctypes as C
import os
import sys
from pathlib import Path

os.environ["A3_K3_TIMING"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from aruco3_b200 import Detector, _ffi, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    base, _ = synth.render_batch("C3", min(n, 32))
    frames = torch.from_numpy(base).repeat((n + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:n].contiguous().cuda()
    h, w = frames.shape[1:3]
    with Detector(dictionary="ARUCO") as det:
        markers = (_ffi.A3Marker * (64 * n))()
        nm, st = C.c_uint32(), _ffi.A3Stats()
        for it in range(4):
            print(f"--- call {it}", file=sys.stderr, flush=True)
            _ffi.check(_ffi.lib().a3_detect_batch(det._h, frames.data_ptr(), _ffi.FMT_RGB8, _ffi.MEM_DEVICE, n, w, h, w * 3, w * h * 3,
                                                  C.cast(markers, C.c_void_p), 64 * n, C.byref(nm), None, C.byref(st)))
        print({k: round(v, 3) if isinstance(v, float) else v for k, v in st.as_dict().items()})


if __name__ == "__main__":
    main()
