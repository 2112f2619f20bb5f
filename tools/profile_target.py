#!/usr/bin/env python3
"""What `ncu` profiles for profiles/: a few resident 256 x 1080p batches through a3_detect_batch with the pose step on
(K1 strips, K3's kernels, K2, K4).  usage: ncu ... python tools/profile_target.py [batch] [calls]"""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from aruco3_b200 import Detector, _ffi, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    base, _ = synth.render_batch("C3", min(n, 32))
    frames = torch.from_numpy(base).repeat((n + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:n].contiguous().cuda()
    h, w = frames.shape[1:3]
    with Detector(dictionary="ARUCO") as det:
        det.set_pose(40.0)
        cap = 64 * n
        markers = (_ffi.A3Marker * cap)()
        poses = (_ffi.A3Pose * (2 * cap))()
        outs = _ffi.A3Outputs()
        outs.marker_poses = C.cast(poses, C.c_void_p).value
        nm, st = C.c_uint32(), _ffi.A3Stats()
        for _ in range(calls):
            _ffi.check(_ffi.lib().a3_detect_batch(det._h, frames.data_ptr(), _ffi.FMT_RGB8, _ffi.MEM_DEVICE, n, w, h, w * 3, w * h * 3,
                                                  C.cast(markers, C.c_void_p), cap, C.byref(nm), C.byref(outs), C.byref(st)))
        print({k: round(v, 3) if isinstance(v, float) else v for k, v in st.as_dict().items()})


if __name__ == "__main__":
    main()
