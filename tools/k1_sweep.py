#!/usr/bin/env python3
"""Sweep the launch-shape knobs of the pixel kernel (K1) on resident frames; prints one line per setting.

    python tools/k1_sweep.py [--frames 64,256] [--mask]

Every setting is checked for byte equality against the generic kernel before it is timed.
"""
import argparse
import ctypes as C
import itertools
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from aruco3_b200 import Detector, _ffi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", default="64,256")
    ap.add_argument("--mask", action="store_true", help="also write the byte mask (5 B/px of traffic)")
    ap.add_argument("--segs", default="0")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    L = _ffi.lib()
    w, h = 1920, 1080
    wpr = (w + 31) // 32
    nmax = max(int(v) for v in args.frames.split(","))
    base, _ = synth.render_batch("C3", 8)
    src = torch.from_numpy(base).cuda().repeat((nmax + 7) // 8, 1, 1, 1)[:nmax].contiguous()
    src += torch.randint(0, 3, src.shape, dtype=torch.uint8, device="cuda")  # decorrelate the copies a little
    grey = torch.empty((nmax, h, w), dtype=torch.uint8, device="cuda")
    mask = torch.empty((nmax, h, w), dtype=torch.uint8, device="cuda") if args.mask else None
    bits = torch.empty((nmax, h, wpr), dtype=torch.int32, device="cuda")
    det = Detector()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(n):
        _ffi.check(L.a3_gray_threshold_batch(det._h, src.data_ptr(), _ffi.FMT_RGB8, _ffi.MEM_DEVICE, n, w, h, w * 3, w * h * 3,
                                             grey.data_ptr(), mask.data_ptr() if args.mask else None, bits.data_ptr(), stream))

    def tune(**kw):
        t = _ffi.A3K1Tuning(**kw)
        _ffi.check(L.a3_detector_set_k1_tuning(det._h, C.byref(t)))

    # reference outputs from the generic kernel
    tune(force_generic=1)
    run(nmax)
    torch.cuda.synchronize()
    ref_grey, ref_bits = grey.clone(), bits.clone()
    ref_mask = mask.clone() if args.mask else None
    bpp_moved = 3 + 1 + 0.125 + (1 if args.mask else 0)
    settings = [dict(force_generic=1)] + [dict(seg_rows=g) for g in [int(v) for v in args.segs.split(",")]]
    for n in [int(v) for v in args.frames.split(",")]:
        for kw in settings:
            tune(**kw)
            grey.zero_(); bits.zero_()
            if args.mask:
                mask.zero_()
            run(n)
            torch.cuda.synchronize()
            ok = bool(torch.equal(grey[:n], ref_grey[:n]) and torch.equal(bits[:n], ref_bits[:n]) and
                      (not args.mask or torch.equal(mask[:n], ref_mask[:n])))
            for _ in range(3):
                run(n)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
            ev[0].record()
            for i in range(args.reps):
                run(n)
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.reps))[args.reps // 2]
            print(json.dumps({"frames": n, **kw, "exact": ok, "ms": round(ms, 4), "fps": round(n / ms * 1e3),
                              "GBps_moved": round(bpp_moved * w * h * n / ms / 1e6, 1),
                              "GBps_algorithmic_5Bpp": round(5 * w * h * n / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
