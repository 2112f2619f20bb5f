#!/usr/bin/env python3
"""Prototype (CPU, pure Python) of an order-independent formulation of imageproc's find_contours — the spec for a GPU
contour stage.  The sequential algorithm (Suzuki-Abe with imageproc's start guards, restated in
oracle/a3ref.c:a3ref_find_contours) decides border starts from labels written by earlier traces; here every start
candidate decides for itself, from the image alone:

  * a border is a closed chain of cracks (foreground pixel, side with a background 4-neighbour); following it from any
    of its cracks visits the same pixels in the same cyclic order (follow() never reads labels);
  * a start candidate is a west crack (p, W) with p.x > 0 ("outer" rule) or an east crack (p, E) with p.x + 1 < w
    ("hole" rule); a border is traced once, from its raster-first ELIGIBLE candidate; east candidates are always
    eligible, a west candidate is eligible iff no other border through p has started before the scan reaches p;
  * a candidate therefore only has to look along its own border for a raster-earlier candidate (cheap: the pixel above
    is usually one), and only a west candidate on a border whose raster-first pixel sits in column 0 (where the outer
    rule is barred) needs the "other borders through p" check.

check(mask) compares the result with the sequential oracle: same contours, same order, same points.
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

DX = [-1, -1, 0, 1, 1, 1, 0, -1]   # ring order w nw n ne e se s sw (screen clockwise)
DY = [0, -1, -1, -1, 0, 1, 1, 1]
W_, N_, E_, S_ = 0, 2, 4, 6


class Img:
    def __init__(self, mask):
        self.fg = np.asarray(mask) > 0
        self.h, self.w = self.fg.shape

    def on(self, x, y):
        return 0 <= x < self.w and 0 <= y < self.h and bool(self.fg[y, x])


def follow(img, sx, sy, adj):
    """The oracle's trace from (sx, sy) with the zero neighbour in ring direction `adj`: list of (x, y, examined_dirs)
    where examined_dirs = ring directions of the zero neighbours looked at while leaving the pixel."""
    first = None
    for k in range(8):
        d = (adj + k) & 7
        if img.on(sx + DX[d], sy + DY[d]):
            first = d
            break
    if first is None:
        return [(sx, sy, set(range(8)))]
    p1 = (sx + DX[first], sy + DY[first])
    out = []
    p3, front = (sx, sy), first
    while True:
        examined = set()
        d4 = front
        for k in range(1, 8):
            d = (front - k) & 7
            if img.on(p3[0] + DX[d], p3[1] + DY[d]):
                d4 = d
                break
            examined.add(d)
        out.append((p3[0], p3[1], examined))
        p4 = (p3[0] + DX[d4], p3[1] + DY[d4])
        if p4 == (sx, sy) and p3 == p1:
            break
        front = (d4 + 4) & 7
        p3 = p4
    return out


def cracks_of_trace(trace):
    """(x, y, side) for side in W/N/E/S whose zero neighbour was examined from that pixel on this border."""
    s = set()
    for x, y, ex in trace:
        for side in (W_, N_, E_, S_):
            if side in ex:
                s.add((x, y, side))
    return s


def raster(p):
    return (p[1], p[0])


def find_contours_parallel(mask):
    img = Img(mask)
    w, h = img.w, img.h
    traces, border_of = {}, {}

    def border(x, y, side):
        """(trace, candidate cracks in raster order) of the border that owns crack (x, y, side)."""
        key = (x, y, side)
        if key not in border_of:
            tr = follow(img, x, y, side)
            cr = cracks_of_trace(tr)
            cands = sorted([(cy, cx, 0) for (cx, cy, s) in cr if s == W_ and cx > 0] + [(cy, cx, 1) for (cx, cy, s) in cr if s == E_ and cx + 1 < w])
            b = (tr, cr, cands)
            for c in cr:
                border_of[c] = b
            border_of[key] = b
        return border_of[key]

    def eligible(y, x, kind, own_cracks):
        if kind == 1:
            # the hole rule is only reached when the outer rule did not fire at this pixel (if / else if)
            return not (x > 0 and not img.on(x - 1, y) and is_start(x, y, 0))
        # the outer rule wants a pixel no earlier trace has touched: no OTHER border through it started before it
        for side in (N_, E_, S_):
            if img.on(x + DX[side], y + DY[side]) or (x, y, side) in own_cracks:
                continue
            _, cr2, cands2 = border(x, y, side)
            for (cy, cx, k2) in cands2:
                if (cy, cx) >= (y, x):
                    break
                if eligible(cy, cx, k2, cr2):
                    return False  # that border was traced before the scan got here: the pixel is no longer 1
        return True

    memo = {}

    def is_start(x, y, kind):
        key = (x, y, kind)
        if key in memo:
            return memo[key]
        _, cr, cands = border(x, y, W_ if kind == 0 else E_)
        res = False
        for (cy, cx, k) in cands:
            if (cy, cx, k) == (y, x, kind):
                res = eligible(cy, cx, k, cr)
                break
            if eligible(cy, cx, k, cr):
                break
        memo[key] = res
        return res

    out = []
    for y in range(h):
        for x in range(w):
            if not img.fg[y, x]:
                continue
            for kind, side, ok in ((0, W_, x > 0 and not img.on(x - 1, y)), (1, E_, x + 1 < w and not img.on(x + 1, y))):
                if ok and is_start(x, y, kind):
                    tr = border(x, y, side)[0]
                    # the sequential trace starts at this pixel with this adjacent zero: re-trace for the point order
                    out.append(((x, y), kind == 0, [(px, py) for (px, py, _) in follow(img, x, y, side)]))
                    break
    return out


def check(mask):
    from oracle import a3ref_py
    cs, outer = a3ref_py.find_contours(np.ascontiguousarray(mask, dtype=np.uint8))
    want = [((int(c[0][0]), int(c[0][1])), bool(t), [tuple(int(v) for v in p) for p in c]) for c, t in zip(cs, outer)]
    got = find_contours_parallel(mask)
    return want == got, want, got


def main():
    rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    trials = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    bad = 0
    for t in range(trials):
        h, w = int(rng.integers(1, 14)), int(rng.integers(1, 14))
        dens = float(rng.choice([0.2, 0.4, 0.5, 0.6, 0.75, 0.9]))
        m = ((rng.random((h, w)) < dens) * 255).astype(np.uint8)
        if t % 3 == 0 and h > 2 and w > 2:
            m[0, :] = m[-1, :] = 0
            m[:, 0] = m[:, -1] = 0
        ok, want, got = check(m)
        if not ok:
            bad += 1
            if bad <= 3:
                print("MISMATCH trial", t)
                print((m > 0).astype(int))
                print("want", [(s, o) for s, o, _ in want])
                print("got ", [(s, o) for s, o, _ in got])
    print("trials", trials, "mismatches", bad)
    return 1 if bad else 0




# ---------------------------------------------------------------------------------------------------------------
# The GPU formulation (csrc/k3_contours.cu follows this): every start candidate walks its own border and survives iff
# no candidate crack of that border is raster-earlier.  West candidates walk backwards (up a left edge the previous
# pixel is usually an earlier candidate), east candidates forwards.  A frame whose survivors include a west crack that
# is not the raster-first pixel of its border (a border whose natural start is barred by the `x > 0` guard) is
# flagged: only there can eligibility differ from "first candidate", and such frames take the sequential host path.

def step_forward(img, p, front):
    """next border pixel from p when the previous one lies in ring direction `front`; -> (next, dir, examined)"""
    examined = []
    for k in range(1, 8):
        d = (front - k) & 7
        if img.on(p[0] + DX[d], p[1] + DY[d]):
            return (p[0] + DX[d], p[1] + DY[d]), d, examined
        examined.append(d)
    return (p[0] + DX[front], p[1] + DY[front]), front, examined


def step_backward(img, p, succ):
    """previous border pixel of p when the next one lies in ring direction `succ`"""
    examined = []
    for k in range(1, 8):
        d = (succ + k) & 7
        if img.on(p[0] + DX[d], p[1] + DY[d]):
            return (p[0] + DX[d], p[1] + DY[d]), d, examined
        examined.append(d)
    return (p[0] + DX[succ], p[1] + DY[succ]), succ, examined


def survivors_gpu_style(mask):
    """-> ([(start, is_outer)] in raster order, flagged)"""
    img = Img(mask)
    w, h = img.w, img.h
    out, flagged = [], False
    for y in range(h):
        for x in range(w):
            if not img.fg[y, x]:
                continue
            for kind, adj, ok in ((0, W_, x > 0 and not img.on(x - 1, y)), (1, E_, x + 1 < w and not img.on(x + 1, y))):
                if not ok:
                    continue
                me = (y, x, kind)
                # first neighbour clockwise from the zero pixel = the predecessor of (x, y) on the border
                pred_dir, ex0 = None, []
                for k in range(8):
                    d = (adj + k) & 7
                    if img.on(x + DX[d], y + DY[d]):
                        pred_dir = d
                        break
                    ex0.append(d)
                if pred_dir is None:  # isolated pixel: its own border
                    alive = not (kind == 1 and x > 0)   # (p, W) of the same pixel comes first
                    if alive:
                        out.append(((x, y), kind == 0))
                    continue
                alive, first_pixel = True, (y, x)

                def earlier(p, examined):
                    nonlocal first_pixel
                    first_pixel = min(first_pixel, (p[1], p[0]))
                    for d, k2 in ((W_, 0), (E_, 1)):
                        if d in examined and ((k2 == 0 and p[0] > 0) or (k2 == 1 and p[0] + 1 < w)) and (p[1], p[0], k2) < me:
                            return True
                    return False

                if kind == 0:
                    # backwards: the start pixel's own visit examines `ex0` going clockwise from W
                    if earlier((x, y), ex0):
                        alive = False
                    p, succ = (x + DX[pred_dir], y + DY[pred_dir]), (pred_dir + 4) & 7
                    start_state = (p, succ)
                    while alive:
                        q, d, ex = step_backward(img, p, succ)
                        if earlier(p, ex):
                            alive = False
                            break
                        p, succ = q, (d + 4) & 7
                        if (p, succ) == start_state:
                            break
                else:
                    p, front = (x, y), pred_dir
                    start_state = (p, front)
                    while alive:
                        q, d, ex = step_forward(img, p, front)
                        if earlier(p, ex):
                            alive = False
                            break
                        p, front = q, (d + 4) & 7
                        if (p, front) == start_state:
                            break
                if alive:
                    if kind == 0 and first_pixel != (y, x):
                        flagged = True
                    out.append(((x, y), kind == 0))
    return out, flagged


def check_gpu_style(mask):
    from oracle import a3ref_py
    cs, outer = a3ref_py.find_contours(np.ascontiguousarray(mask, dtype=np.uint8))
    want = [((int(c[0][0]), int(c[0][1])), bool(t)) for c, t in zip(cs, outer)]
    got, flagged = survivors_gpu_style(mask)
    return flagged, want == got, want, got


def main_gpu_style():
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    trials = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
    bad = nflag = 0
    for t in range(trials):
        h, w = int(rng.integers(1, 16)), int(rng.integers(1, 16))
        dens = float(rng.choice([0.2, 0.4, 0.5, 0.6, 0.75, 0.9, 0.97]))
        m = ((rng.random((h, w)) < dens) * 255).astype(np.uint8)
        if t % 3 == 0 and h > 2 and w > 2:
            m[0, :] = m[-1, :] = 0
            m[:, 0] = m[:, -1] = 0
        flagged, ok, want, got = check_gpu_style(m)
        nflag += flagged
        if not flagged and not ok:
            bad += 1
            if bad <= 3:
                print("MISMATCH (unflagged) trial", t)
                print((m > 0).astype(int))
                print("want", want)
                print("got ", got)
    print("gpu-style trials", trials, "flagged", nflag, "unflagged mismatches", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main_gpu_style() if len(sys.argv) > 1 and sys.argv[1] == "gpu" else main())
