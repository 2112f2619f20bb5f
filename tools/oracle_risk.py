#!/usr/bin/env python3
"""Quantify the two known risks of the oracle's restatement of imageproc without a Rust toolchain.

R3  `Projection::from_control_points` solves its 8x8 system with nalgebra's f64 SVD; the oracle and kernel K2 use f64
    Gaussian elimination with partial pivoting.  Both are backward stable, so the f64 solutions agree to ~cond * 2^-52, and
    what matters is how often the cast to f32 lands on a different float, and what that does downstream.  Here every quad is
    solved three ways — the oracle's elimination, an independent f64 SVD (LAPACK through numpy, solved the way nalgebra's
    `svd.solve` does: V diag(1/s) U^T b) and, as the arbiter, the SVD in 80-bit long double — and the f32 coefficients, the
    49x49 patch bytes, the four codes and the decoded (id, rotation, distance) are compared.
R4  imageproc's bilinear blend quantises the two horizontal blends to u8 before the vertical blend (restated from 0.25);
    a one-stage blend (truncated, or rounded to nearest) is what other versions / libraries do.  The same comparison.
ulp A forward transform whose coefficients are each moved by one f32 ulp at random: how fragile patch and bits are to ANY
    last-bit difference in the solve.

    python tools/oracle_risk.py [--fuzz 100000] [--out profiles/r02_oracle_risk.json]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

from aruco3_b200 import synth  # noqa: E402
from oracle import a3ref_py  # noqa: E402

HS = 49


def _lib():
    L = a3ref_py.lib()
    L.a3ref_projection_from_control_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    L.a3ref_projection_from_control_points.restype = C.c_int
    L.a3ref_warp_with_transform.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
    L.a3ref_warp_with_transform.restype = C.c_int
    L.a3ref_homography_to_code_permutations.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint8, C.c_void_p, C.c_void_p, C.c_void_p]
    L.a3ref_homography_to_code_permutations.restype = C.c_int
    return L


def systems(quads: np.ndarray):
    """The reference's 8x8 systems for quads [n,8] (u32 corners cast to f32 then f64) -> A [n,8,8], b [n,8]."""
    q = quads.astype(np.float32).astype(np.float64).reshape(-1, 4, 2)
    to = np.array([[0, 0], [HS, 0], [HS, HS], [0, HS]], np.float64)
    n = len(q)
    A = np.zeros((n, 8, 8))
    b = np.zeros((n, 8))
    for k in range(4):
        xf, yf = q[:, k, 0], q[:, k, 1]
        x, y = to[k]
        A[:, 2 * k, 3], A[:, 2 * k, 4], A[:, 2 * k, 5], A[:, 2 * k, 6], A[:, 2 * k, 7] = -xf, -yf, -1.0, y * xf, y * yf
        b[:, 2 * k] = -y
        A[:, 2 * k + 1, 0], A[:, 2 * k + 1, 1], A[:, 2 * k + 1, 2], A[:, 2 * k + 1, 6], A[:, 2 * k + 1, 7] = xf, yf, 1.0, -x * xf, -x * yf
        b[:, 2 * k + 1] = x
    return A, b


def solve_svd(A, b):
    """x = V diag(1/s) U^T b, singular values <= eps dropped (nalgebra svd.solve(b, f64::EPSILON))."""
    U, s, Vt = np.linalg.svd(A)
    utb = np.einsum("nji,nj->ni", U, b)
    inv = np.where(s > np.finfo(np.float64).eps, 1.0 / s, 0.0)
    return np.einsum("nji,nj->ni", Vt, utb * inv)


def solve_exact(A, b):
    """Arbiter: numpy's LU in f64 followed by two steps of iterative refinement with long-double residuals."""
    good = np.abs(np.linalg.det(A)) > 0
    x = np.full(b.shape, np.nan)
    Ag, bg = A[good], b[good]
    xg = np.linalg.solve(Ag, bg[..., None])[..., 0]
    Al, bl = Ag.astype(np.longdouble), bg.astype(np.longdouble)
    for _ in range(2):
        r = bl - np.einsum("nij,nj->ni", Al, xg.astype(np.longdouble))
        xg = (xg.astype(np.longdouble) + np.linalg.solve(Ag, r.astype(np.float64)[..., None])[..., 0].astype(np.longdouble)).astype(np.float64)
    x[good] = xg
    return x


def solve_oracle(quads):
    L = _lib()
    out = np.zeros((len(quads), 9), np.float32)
    ok = np.zeros(len(quads), bool)
    to = np.array([0, 0, HS, 0, HS, HS, 0, HS], np.float32)
    inv = np.zeros(9, np.float32)
    cls = C.c_int()
    for i, q in enumerate(quads):
        fr = q.astype(np.float32)
        ok[i] = bool(L.a3ref_projection_from_control_points(fr.ctypes.data, to.ctypes.data, out[i].ctypes.data, inv.ctypes.data, C.byref(cls)))
    return out, ok


def warp(grey, t9, variant=0):
    L = _lib()
    patch = np.zeros((HS, HS), np.uint8)
    t = np.ascontiguousarray(t9, np.float32)
    ok = L.a3ref_warp_with_transform(grey.ctypes.data, grey.shape[1], grey.shape[0], t.ctypes.data, HS, variant, patch.ctypes.data)
    return patch, bool(ok)


def decode(patch, d, ms, tau):
    """-> (has_codes, codes tuple, (id, rotation, distance) or None with the reference's acceptance rule)."""
    L = _lib()
    codes = (C.c_uint64 * 4)()
    otsu = C.c_uint8()
    red = np.zeros(ms * ms, np.uint8)
    good = L.a3ref_homography_to_code_permutations(patch.ctypes.data, HS, HS, ms, codes, C.byref(otsu), red.ctypes.data)
    if not good:
        return False, None, None
    best = (256, 0, 0)
    for r in range(4):
        idx, dist = a3ref_py.find_nearest(d, int(codes[r]))
        if dist < best[0]:
            best = (dist, r, idx)
    acc = (best[2], best[1], best[0]) if best[0] < tau else None
    return True, tuple(int(c) for c in codes), acc


def t9_from(x8):
    return np.concatenate([x8.astype(np.float32), np.ones((len(x8), 1), np.float32)], axis=1)


def ulp_jitter(t9, rng):
    t = t9.copy()
    for i in range(8):
        step = rng.integers(-1, 2)
        if step:
            t[i] = np.nextafter(t[i], np.float32(np.inf if step > 0 else -np.inf))
    return t


def run(workloads=(("C1", 8), ("C3", 6), ("C5", 2)), n_fuzz=100000, seed=7, patch_sample=1500):
    rng = np.random.default_rng(seed)
    report = {"how": "tools/oracle_risk.py (see its header)", "homography_sample_size": HS, "workloads": {}, "fuzz": {}}
    # ---- the rendered workloads: every candidate the oracle finds ----
    fuzz_base = []
    for name, frames in workloads:
        spec = synth.CONFIGS[name]
        cfg = a3ref_py.default_config()
        cfg.min_corner_separation_factor = spec.min_corner_separation_factor
        d = a3ref_py.dictionary(spec.dictionary)
        ms = int(a3ref_py.lib().a3ref_mark_size(C.byref(d)))
        stats = dict(candidates=0, coeff_sets_differing_svd=0, coeffs_differing_svd=0, coeff_sets_differing_exact=0, patches_differing_svd=0,
                     patch_bytes_differing_svd=0, codes_differing_svd=0, ids_differing_svd=0,
                     patches_differing_onestage_trunc=0, patch_bytes_differing_onestage_trunc=0, codes_differing_onestage_trunc=0,
                     ids_differing_onestage_trunc=0, codes_differing_onestage_round=0, ids_differing_onestage_round=0,
                     patches_differing_ulp=0, codes_differing_ulp=0, ids_differing_ulp=0)
        for f in range(frames):
            img, _ = synth.render_frame(spec, f)
            res = a3ref_py.detect(img, spec.dictionary, cfg)
            quads = res.candidates.astype(np.uint32)
            if not len(quads):
                continue
            A, b = systems(quads)
            t_ge, ok = solve_oracle(quads)
            t_svd, t_ex = t9_from(solve_svd(A, b)), t9_from(solve_exact(A, b))
            for i in range(len(quads)):
                if not ok[i]:
                    continue
                stats["candidates"] += 1
                nd = int((t_ge[i] != t_svd[i]).sum())
                stats["coeffs_differing_svd"] += nd
                stats["coeff_sets_differing_svd"] += nd > 0
                stats["coeff_sets_differing_exact"] += int((t_ge[i] != t_ex[i]).any())
                p0, _ = warp(res.grey, t_ge[i], 0)
                base = decode(p0, d, ms, d.tau)
                for key, t, variant in (("svd", t_svd[i], 0), ("onestage_trunc", t_ge[i], 1), ("onestage_round", t_ge[i], 2),
                                        ("ulp", ulp_jitter(t_ge[i], rng), 0)):
                    if key == "svd" and nd == 0:
                        continue
                    p1, _ = warp(res.grey, t, variant)
                    got = decode(p1, d, ms, d.tau)
                    nb = int((p0 != p1).sum())
                    if f"patches_differing_{key}" in stats:
                        stats[f"patches_differing_{key}"] += nb > 0
                    if f"patch_bytes_differing_{key}" in stats:
                        stats[f"patch_bytes_differing_{key}"] += nb
                    stats[f"codes_differing_{key}"] += got[:2] != base[:2]
                    stats[f"ids_differing_{key}"] += got[2] != base[2]
            if name == "C3":
                fuzz_base.append(quads[ok])
        report["workloads"][name] = {k: int(v) for k, v in stats.items()} | {"frames": frames, "dictionary": spec.dictionary}
    # ---- fuzz: realistic marker quads (C3 candidates) with every corner moved by up to +-12 px ----
    base_quads = np.concatenate(fuzz_base)
    pick = rng.integers(0, len(base_quads), size=n_fuzz)
    fq = (base_quads[pick].astype(np.int64) + rng.integers(-12, 13, size=(n_fuzz, 8))).clip(0, 4000).astype(np.uint32)
    A, b = systems(fq)
    t_ge, ok = solve_oracle(fq)
    t_svd, t_ex = t9_from(solve_svd(A, b)), t9_from(solve_exact(A, b))
    ok &= np.isfinite(t_ex).all(1) & np.isfinite(t_svd).all(1)
    dif_svd = (t_ge != t_svd) & ok[:, None]
    dif_ex = (t_ge != t_ex) & ok[:, None]
    dif_svd_ex = (t_svd != t_ex) & ok[:, None]
    cond = np.linalg.cond(A[ok][:2000])
    fz = {"quads": int(ok.sum()), "coeff_sets_differing_ge_vs_svd": int(dif_svd.any(1).sum()), "coeffs_differing_ge_vs_svd": int(dif_svd.sum()),
          "coeff_sets_differing_ge_vs_exact": int(dif_ex.any(1).sum()), "coeff_sets_differing_svd_vs_exact": int(dif_svd_ex.any(1).sum()),
          "condition_number_median": float(np.median(cond)), "condition_number_max": float(cond.max())}
    # downstream effect on the quads whose coefficients differ (all of them, capped) on a C3 frame's grey
    spec = synth.CONFIGS["C3"]
    img, _ = synth.render_frame(spec, 0)
    grey = a3ref_py.to_luma8(img)
    d = a3ref_py.dictionary("ARUCO")
    ms = int(a3ref_py.lib().a3ref_mark_size(C.byref(d)))
    idx = np.flatnonzero(dif_svd.any(1))[:patch_sample]
    fz.update(patch_checked=len(idx), patches_differing=0, patch_bytes_differing=0, codes_differing=0, ids_differing=0)
    for i in idx:
        p0, _ = warp(grey, t_ge[i])
        p1, _ = warp(grey, t_svd[i])
        nb = int((p0 != p1).sum())
        fz["patches_differing"] += nb > 0
        fz["patch_bytes_differing"] += nb
        a, c = decode(p0, d, ms, d.tau), decode(p1, d, ms, d.tau)
        fz["codes_differing"] += a[:2] != c[:2]
        fz["ids_differing"] += a[2] != c[2]
    report["fuzz"] = fz
    return report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fuzz", type=int, default=100000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rep = run(n_fuzz=args.fuzz)
    text = json.dumps(rep, indent=1)
    print(text)
    if args.out:
        Path(args.out).write_text(text + "\n")


if __name__ == "__main__":
    main()
