#!/usr/bin/env python3
"""Benchmark of the detection hot path (BASELINE.json: frames/sec at 1080p, batch 256, on 1/2/4/8 B200;
threshold-kernel HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the C restatement of the reference (oracle/)

One step = one pass of `Detector::detect` (src/aruco.rs:52-121) over one batch of 256 synthetic 1920x1080 RGB
frames with 20 ARUCO markers each (BASELINE.json configs[2], aruco3_b200/synth.py "C3") PER GPU: frames are
independent, so N GPUs run N shards with no collective (weak scaling; torch.distributed is used only for the
barrier and the max-over-ranks of the timing).
  value  frames/s with the RGB frames already resident in HBM (a3_detect_batch, A3_MEM_DEVICE)
  e2e    frames/s through the same call with pinned HOST frames: H2D of the frames and D2H of the mask bits,
         decode records and markers are inside the timed region
  roofline  K1 (fused gray + adaptive threshold), 5 algorithmic bytes per pixel, timed with CUDA events on the
         library's own stream inside the timed region (a3_stats.ms_pixel_kernel)
The reference (Rust) cannot be built in this image; `--impl reference` and `cpu_baseline` time oracle/a3ref.c
("port"), frame-parallel over the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WORKLOAD = "C3"            # 1920x1080 RGB, 20 ARUCO markers per frame, noise 0
BATCH = 256
ALGO_BYTES_PER_PIXEL = 5   # 3 B RGB read + 1 B grey written + 1 B mask written


def bind_near_gpu(index):
    """One process per GPU: run this rank (and so first-touch its pinned frames) on the cores NVML reports as local to the
    GPU, so that the H2D stream of every rank reads host memory of its own socket.  Returns the number of cores, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index), (ncpu + 63) // 64)
        cpus = {i * 64 + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def rank_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, threading.Event(), [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self.stop_flag.wait(0.002)

    def result(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def render(n, first_index, out):
    from aruco3_b200 import synth
    synth.render_batch(WORKLOAD, n, first_index, out=out)


def cpu_reference_run(frames: np.ndarray, threads: int, steps: int, warmup: int):
    """oracle/a3ref.c, frame-parallel: the only place bench.py executes oracle/ (as the CPU arm, never on the product path)."""
    from oracle import a3ref_py
    for _ in range(warmup):
        a3ref_py.detect_many(frames[:max(1, min(len(frames), threads))], "ARUCO", threads=threads)
    t0 = time.perf_counter()
    markers = 0
    for _ in range(steps):
        m, st = a3ref_py.detect_many(frames, "ARUCO", threads=threads)
        markers += m
    dt = time.perf_counter() - t0
    return len(frames) * steps / dt, dt / steps, markers // max(steps, 1), st


def run_reference(args):
    rank, _, world = rank_info()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = min(BATCH, max(32, 2 * cores))
    frames = np.empty((sample, 1080, 1920, 3), np.uint8)
    render(sample, 0, frames)
    fps, s_per_step, markers, st = cpu_reference_run(frames, cores, max(1, args.steps), min(args.warmup, 1))
    line = {"impl": "reference", "metric": "frames_per_sec_1080p_batch256", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32 integer + f32/f64 decode",
            "data": "synthetic",
            "config": {"workload": f"{WORKLOAD}: 1920x1080 RGB8, 20 ARUCO markers/frame, noise 0 (BASELINE.json configs[2])",
                       "frames_per_step": sample, "host_threads": cores},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} frames of the {BATCH}-frame batch per step, frame-parallel over {cores} threads "
                                       f"(oracle/a3ref.c; the Rust reference cannot be built in this image)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "markers_per_step": markers, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="frames per GPU per step (the metric is quoted at 256)")
    ap.add_argument("--host-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk", type=int, default=0, help="frames per pipeline chunk (0 = library default)")
    ap.add_argument("--contours", default="device", choices=["device", "host"], help="where find_contours + quad filters run")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    from aruco3_b200 import Detector, _ffi

    rank, local_rank, world = rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = bind_near_gpu(local_rank) if world > 1 else None  # before the pinned buffers are allocated and touched
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cores = os.cpu_count() or 1
    host_threads = args.host_threads or max(1, cores // max(world, 1))
    n, h, w = args.batch, 1080, 1920

    # ---- synthetic frames: pinned host copy (e2e arm) and a resident device copy (value arm) ----
    from aruco3_b200.sharding import shard_range
    lo, hi = shard_range(n * world, rank, world)  # this rank's contiguous block of the global batch (weak scaling: 256 per GPU)
    assert hi - lo == n
    pinned = torch.empty((n, h, w, 3), dtype=torch.uint8, pin_memory=True)
    render(n, lo, pinned.numpy())
    resident = pinned.cuda(non_blocking=False)
    torch.cuda.synchronize()

    det = Detector(dictionary="ARUCO", device=local_rank, host_threads=host_threads, contours=args.contours)
    L = _ffi.lib()
    if args.chunk:
        tune = _ffi.A3K1Tuning(chunk_frames=args.chunk)
        _ffi.check(L.a3_detector_set_k1_tuning(det._h, C.byref(tune)))
    cap = 64 * n
    markers = (_ffi.A3Marker * cap)()
    n_markers = C.c_uint32()
    stats = _ffi.A3Stats()

    def step(ptr, mem, want_stats=True):
        _ffi.check(L.a3_detect_batch(det._h, ptr, _ffi.FMT_RGB8, mem, n, w, h, w * 3, w * h * 3, C.cast(markers, C.c_void_p),
                                     cap, C.byref(n_markers), None, C.byref(stats) if want_stats else None))
        return stats.as_dict() if want_stats else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(ptr, mem, steps, warmup, stats_in_loop=True):
        """stats_in_loop=False: the timed calls pass no a3_stats (the library then skips its ~100 CUDA-event queries per chunked
        call); the stage statistics come from `steps` instrumented calls after the timed region instead."""
        for _ in range(warmup):
            step(ptr, mem)
        acc = {}
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            s = step(ptr, mem, stats_in_loop)
            for k, v in (s or {}).items():
                acc[k] = acc.get(k, 0) + v
        e1.record()
        barrier()
        clocks = sampler.result()
        if not stats_in_loop:
            for _ in range(steps):
                for k, v in step(ptr, mem).items():
                    acc[k] = acc.get(k, 0) + v
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, acc, clocks

    ms_dev, acc_dev, clocks = timed(resident.data_ptr(), _ffi.MEM_DEVICE, args.steps, args.warmup)
    ms_e2e, acc_e2e, clocks_e2e = timed(pinned.data_ptr(), _ffi.MEM_HOST, args.steps, 1, stats_in_loop=False)

    # ---- BASELINE.json configs[1] beside the headline: one 1080p frame from host memory per call, wall-clock latency ----
    def single_frame_latency(frame_tensor, reps=40):
        one = (_ffi.A3Marker * 4096)()
        lat = []
        for it in range(reps + 5):
            t0 = time.perf_counter()
            _ffi.check(L.a3_detect_batch(det._h, frame_tensor.data_ptr(), _ffi.FMT_RGB8, _ffi.MEM_HOST, 1, w, h, w * 3, w * h * 3,
                                         C.cast(one, C.c_void_p), 4096, C.byref(n_markers), None, None))
            if it >= 5:
                lat.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(lat))

    latency = None
    if rank == 0:
        noise = torch.empty((1, h, w, 3), dtype=torch.uint8, pin_memory=True)
        noise.numpy()[:] = np.random.default_rng(0xA3C0DE00 + 2000).integers(0, 256, size=(1, h, w, 3), dtype=np.uint8)
        latency = {"unit": "ms per detect() call, host frame in, markers out (median of 40)",
                   "marker_frame_1080p": single_frame_latency(pinned[:1]),
                   "noise_frame_1080p_reference_bench_workload": single_frame_latency(noise)}

    # ---- K1 alone on the resident frames (isolated figure; the roofline entry uses the in-pipeline time) ----
    wpr = (w + 31) // 32
    d_grey = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    d_bits = torch.empty((n, h, wpr), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    def k1_alone():
        _ffi.check(L.a3_gray_threshold_batch(det._h, resident.data_ptr(), _ffi.FMT_RGB8, _ffi.MEM_DEVICE, n, w, h, w * 3, w * h * 3,
                                             d_grey.data_ptr(), None, d_bits.data_ptr(), C.c_void_p(stream)))
    for _ in range(3):
        k1_alone()
    torch.cuda.synchronize()
    reps = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        k1_alone()
        ev[i + 1].record()
    torch.cuda.synchronize()
    k1_ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))[reps // 2]

    peak, peak_src = measured_peaks()
    launches = int(acc_dev["pixel_kernel_launches"])
    k1_avg_ms = acc_dev["ms_pixel_kernel"] / max(launches, 1)
    frames_per_launch = n * args.steps / max(launches, 1)
    bytes_per_launch = ALGO_BYTES_PER_PIXEL * w * h * frames_per_launch
    achieved = bytes_per_launch / (k1_avg_ms * 1e-3) / 1e9 if k1_avg_ms > 0 else 0.0
    # what the in-pipeline launch really moves: RGB in, grey + 1-bit mask out (the byte mask is only written on request)
    moved = (3 + 1 + 0.125) * w * h * frames_per_launch / (k1_avg_ms * 1e-3) / 1e9 if k1_avg_ms > 0 else 0.0
    # the isolated launch writes grey + bits (no byte mask): 3 + 1 + 1/8 bytes per pixel
    iso_bytes = (3 + 1 + 0.125) * w * h * n
    iso_gbs = iso_bytes / (k1_ms * 1e-3) / 1e9

    # DRAM bytes per K1 launch from the committed ncu --set full capture (profiles/k1_traffic.json), scaled to this
    # launch's frame count; null when the file is absent
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "k1_traffic.json"
    if tf.exists():
        t = json.loads(tf.read_text())
        traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) * (bytes_per_launch / (5.0 * w * h)) / t["frames_per_launch"]
        traffic_src = t["source"]

    total_frames = n * world * args.steps
    value = total_frames / (ms_dev * 1e-3)
    e2e_value = total_frames / (ms_e2e * 1e-3)
    # device -> host per step: the decode records, plus the mask bits (host contour stage) or K3's quads (first 64 per frame)
    # and five per-frame counters (device contour stage)
    d2h_front = n * h * wpr * 4 if args.contours == "host" else n * (64 * 32 + 4 * 4 + 8)
    d2h = int(d2h_front + acc_e2e["n_candidates"] / args.steps * C.sizeof(_ffi.A3Decode))
    line = {
        "metric": "frames_per_sec_1080p_batch256", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/u32 integer pixels, f32/f64 decode", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: 1920x1080 RGB8 x {n} frames per GPU, 20 ARUCO markers/frame, noise 0 (BASELINE.json configs[2])",
                   "frames_per_gpu_per_step": n, "host_threads_per_rank": host_threads, "host_cores": cores,
                   "rank_cpu_affinity_cores": numa,
                   "l2": f"inputs larger than L2 ({n * h * w * 3 / 1e6:.0f} MB of RGB per step per GPU, never re-read)",
                   "parallelism": f"frame-batch sharding x{world}, no collective"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": n * h * w * 3, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        # ours per step: K1, K2, K3's eight kernels (candidates, walk_short, walkers, flag_all, order, emit, rdp, finalize) and, on
        # the one-shot route, the size check of the speculative K3 finish, the two kernels that gather its quads for K2 and the
        # marker assembly
        "gpu_launches": int(acc_dev["pixel_kernel_launches"] + acc_dev["decode_kernel_launches"] + acc_dev["pose_kernel_launches"]
                            + 8 * acc_dev["contour_kernel_launches"] + 4 * acc_dev["one_shot"]),
        "one_shot_route_steps": int(acc_dev["one_shot"]),
        "contour_stage": args.contours, "host_fallback_frames_per_step": acc_dev["host_fallback_frames"] / args.steps,
        "roofline": {"bound": "hbm", "kernel": "k1_strips_kernel<RGB8> (fused into_luma8 + adaptive_threshold, TMA tensor tiles)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "frac_of_nominal_8TBps": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_src,
                     "achieved_moved": moved, "frac_moved": moved / peak,
                     "note": "achieved = SURVEY 8d algorithmic 5 B/px (3 RGB + 1 grey + 1 mask) / K1 time; the pipeline writes the mask "
                             "1 bit/px, so it moves 4.125 B/px: achieved_moved",
                     "bytes_per_launch": bytes_per_launch, "avg_launch_ms": k1_avg_ms, "launches": launches,
                     "isolated": {"gbs": iso_gbs, "ms": k1_ms, "bytes": iso_bytes, "frac": iso_gbs / peak,
                                  "fps": n / (k1_ms * 1e-3), "note": "K1 alone over the 256 resident frames, grey + 1-bit mask outputs"}},
        "stages_ms_per_step": {k: acc_dev[k] / args.steps for k in ("ms_h2d", "ms_pixel_kernel", "ms_contour_kernels", "ms_mask_d2h", "ms_host_quads",
                                                                     "ms_host_cpu", "ms_decode_kernel", "ms_total")},
        # from instrumented calls after the timed region (the timed e2e calls pass no a3_stats)
        "stages_ms_per_step_e2e": {k: acc_e2e[k] / args.steps for k in ("ms_h2d", "ms_pixel_kernel", "ms_contour_kernels", "ms_mask_d2h", "ms_host_quads",
                                                                         "ms_host_cpu", "ms_decode_kernel", "ms_total")},
        "counts_per_step": {k: acc_dev[k] / args.steps for k in ("n_contours", "n_contour_points", "n_candidates", "n_markers")},
        "clocks": clocks, "clocks_e2e": clocks_e2e,
        "single_frame_latency": latency,
    }
    # the latency-bound stages are reported as units per second of their own kernel time, not as roofline fractions
    if acc_dev["ms_decode_kernel"] > 0:
        line["k2_decode"] = {"candidates_per_s": acc_dev["n_candidates"] / (acc_dev["ms_decode_kernel"] * 1e-3), "unit": "candidates/s",
                             "ms_per_step": acc_dev["ms_decode_kernel"] / args.steps, "candidates_per_step": acc_dev["n_candidates"] / args.steps}
    if acc_dev["ms_contour_kernels"] > 0:
        line["k3_contours"] = {"border_points_per_s": acc_dev["n_contour_points"] / (acc_dev["ms_contour_kernels"] * 1e-3), "unit": "border points/s",
                               "ms_per_step": acc_dev["ms_contour_kernels"] / args.steps}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = min(n, max(32, 2 * cores))
        fps, s_per_step, mk, st = cpu_reference_run(pinned.numpy()[:sample], cores, 1, 1)
        fps1, _, _, st1 = cpu_reference_run(pinned.numpy()[:4], 1, 1, 0)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"{sample} of the {n} frames, frame-parallel over {cores} threads (oracle/a3ref.c)",
                                "single_thread_fps": fps1,
                                "single_thread_stage_ms_per_frame": {k: st1[k] / 4 for k in ("ms_gray", "ms_threshold", "ms_contours",
                                                                                              "ms_quads", "ms_warp", "ms_decode")}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    det.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
