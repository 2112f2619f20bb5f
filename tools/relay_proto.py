#!/usr/bin/env python3
"""Prototype (CPU, pure Python) of K3's relay walks (k3_contours.cu: k3_segments / k3_cycles): the spec the kernels follow, checked
against a plain trace of every border on random masks.

A border is a cycle of VISITS (pixel, state = ring direction of the previous border pixel); the step function reads only the 3x3
neighbourhood.  Every crack (foreground pixel, side with a background 4-neighbour) is examined by exactly one visit, which the
neighbourhood names: state = the first foreground neighbour clockwise from the crack's side.  RELAYS are the visits that own a
candidate crack (west with x > 0, east with x + 1 < w) on a row y % R == 0 (the west crack names the visit when it owns both).  One walker per relay follows the
border to the next relay and records (next relay, steps, the raster-first candidate crack on the way and its step, the
raster-first pixel).  The relays of a border then form a cycle of segments; its length is the sum of the steps, its start — the
reference's discovery point — is the smallest candidate key over the segments, and a segment's points go to
(offset of the segment - offset of the start visit) mod n.  Borders that own no relay crack are left to the candidate walks.

    python tools/relay_proto.py [trials]
"""
import sys

import numpy as np

DX = [-1, -1, 0, 1, 1, 1, 0, -1]   # ring order w nw n ne e se s sw (screen clockwise)
DY = [0, -1, -1, -1, 0, 1, 1, 1]
W_, E_ = 0, 4
NONE = 0xffffffff


class Img:
    def __init__(self, mask):
        self.fg = np.asarray(mask) > 0
        self.h, self.w = self.fg.shape

    def on(self, x, y):
        return 0 <= x < self.w and 0 <= y < self.h and bool(self.fg[y, x])


def step(img, x, y, state):
    """One visit: (direction of the next pixel, examined directions).  k3_contours.cu:build_tables (fwd)."""
    examined = 0
    d = state
    for k in range(1, 8):
        c = (state - k) & 7
        if img.on(x + DX[c], y + DY[c]):
            d = c
            break
        examined |= 1 << c
    return d, examined


def owner_state(img, x, y, side):
    """State of the visit that owns crack (x, y, side), or None for an isolated pixel."""
    for k in range(8):
        d = (side + k) & 7
        if img.on(x + DX[d], y + DY[d]):
            return d
    return None


def trace_from(img, x, y, state):
    """All visits of the border through visit (x, y, state), starting there: [(x, y, state, examined)]."""
    out = []
    cx, cy, cs = x, y, state
    while True:
        d, ex = step(img, cx, cy, cs)
        out.append((cx, cy, cs, ex))
        cx, cy, cs = cx + DX[d], cy + DY[d], (d + 4) & 7
        if (cx, cy, cs) == (x, y, state):
            return out


def cand_keys(img, x, y, ex):
    """Candidate cracks of a visit as raster keys (pix << 1 | kind): west needs x > 0, east x + 1 < w."""
    pix = y * img.w + x
    keys = []
    if (ex >> W_) & 1 and x > 0:
        keys.append(pix << 1)
    if (ex >> E_) & 1 and x + 1 < img.w:
        keys.append(pix << 1 | 1)
    return keys


def relay_borders(mask, R, J=32):
    """The relay scheme: {start candidate key: (n, points)} for every border that owns a relay crack and a candidate."""
    img = Img(mask)
    w, h = img.w, img.h
    # ---- enumeration (the candidate scan): west / east cracks on relay rows, their owning visits; aliases dropped ----
    relays = []   # (x, y, state)
    index = {}    # (x, y, side) -> relay index (the kernels: word base + rank inside the word)
    for y in range(0, h, R):
        for x in range(w):
            if not img.on(x, y):
                continue
            for side in (W_, E_):
                if img.on(x + DX[side], y + DY[side]) or (side == W_ and x == 0) or (side == E_ and x + 1 == w):
                    continue  # relays are candidate cracks: west with x > 0, east with x + 1 < w
                st = owner_state(img, x, y, side)
                if st is None:
                    continue  # isolated pixel: the candidate walks own it
                _, ex = step(img, x, y, st)
                assert (ex >> side) & 1, "the owning visit examines its crack"
                index[(x, y, side)] = len(relays)
                alias = side == E_ and (ex >> W_) & 1 and x > 0
                relays.append(None if alias else (x, y, st))
    # ---- segments: one walker per relay ----
    seg = [None] * len(relays)
    for i, r in enumerate(relays):
        if r is None:
            continue
        x, y, st = r
        t, best, best_pos, min_pix = 0, NONE, 0, NONE
        while True:
            d, ex = step(img, x, y, st)
            if t > 0 and y % R == 0 and (((ex >> W_) & 1 and x > 0) or ((ex >> E_) & 1 and x + 1 < w)):
                side = W_ if ((ex >> W_) & 1 and x > 0) else E_
                nxt = index[(x, y, side)]
                assert relays[nxt] is not None, "a walker arrives at the canonical crack of a relay visit"
                break
            for k in cand_keys(img, x, y, ex):
                if k < best:
                    best, best_pos = k, t
            min_pix = min(min_pix, y * w + x)
            x, y, st = x + DX[d], y + DY[d], (d + 4) & 7
            t += 1
        seg[i] = (nxt, t, best, best_pos, min_pix)
    # ---- jumps: every relay sums up the next J segments (J dependent steps, all relays at once), so that a cycle of k segments is
    # followed in k / J steps: (relay after J hops, visits, best candidate and its offset, first pixel, smallest index landed on) ----
    jump = [None] * len(relays)
    for i, s in enumerate(seg):
        if s is None:
            continue
        j, total, best, mp, mi, hops = i, 0, (NONE, 0), NONE, NONE, 0
        while True:
            nxt, ln, cand, pos, pix = seg[j]
            if cand < best[0]:
                best = (cand, total + pos)
            mp = min(mp, pix)
            total += ln
            j = nxt
            hops += 1
            mi = min(mi, j)
            if hops == J or j == i:
                break
        jump[i] = (j, total, best[0], best[1], mp, mi)
    # ---- cycles: the smallest relay index of a cycle is its leader.  A walker jumps while every relay it would land on or pass
    # has a larger index, leaving (its index, visits so far) at the relays it jumps from (anchors); when its own index is
    # among the next J it closes the cycle step by step; when a smaller index is, it is not the leader.  Minimum wins everywhere:
    # the leader has the smallest index and passes everything. ----
    owner = [(NONE, 0)] * len(relays)
    anchor = [(NONE, 0)] * len(relays)
    info = {}
    for i, s in enumerate(seg):
        if s is None:
            continue
        j, total, best, leader = i, 0, (NONE, 0), True
        while True:
            dest, jt, jc, jo, jmp, jmi = jump[j]
            if jmi < i:
                leader = False
                break
            if jmi == i:  # the span from j comes back to me: close step by step
                while True:
                    nxt, ln, cand, pos, mp = seg[j]
                    owner[j] = min(owner[j], (i, total))
                    if cand < best[0]:
                        best = (cand, total + pos)
                    total += ln
                    j = nxt
                    if j == i:
                        break
                break
            anchor[j] = min(anchor[j], (i, total))
            if jc < best[0]:
                best = (jc, total + jo)
            total += jt
            j = dest
        if leader and best[0] != NONE:
            info[i] = (best[0], best[1], total)
    # ---- spread: every anchor hands (owner, visits so far) to the J segments after it ----
    for a_, (own, cum) in enumerate(anchor):
        if own == NONE:
            continue
        j = a_
        for _ in range(J):
            owner[j] = min(owner[j], (own, cum))
            cum += seg[j][1]
            j = seg[j][0]
    # ---- emission: one walker per relay again ----
    out = {k: (n, [None] * n) for (k, _, n) in info.values()}
    for j, s in enumerate(seg):
        if s is None or owner[j][0] not in info:
            continue
        key, start, n = info[owner[j][0]]
        x, y, st = relays[j]
        off = (owner[j][1] - start) % n
        for t in range(s[1]):
            out[key][1][(off + t) % n] = (x, y)
            d, _ = step(img, x, y, st)
            x, y, st = x + DX[d], y + DY[d], (d + 4) & 7
    for k, (n, pts) in out.items():
        assert all(p is not None for p in pts)
    return out


def plain_borders(mask, R):
    """Every border traced plainly from its raster-first candidate crack: {key: (n, points, owns a relay crack)}."""
    img = Img(mask)
    w, h = img.w, img.h
    seen = set()
    out = {}
    for y in range(h):
        for x in range(w):
            if not img.on(x, y):
                continue
            for side in (W_, E_):
                if img.on(x + DX[side], y + DY[side]):
                    continue
                st = owner_state(img, x, y, side)
                if st is None or (x, y, st) in seen:
                    continue
                tr = trace_from(img, x, y, st)
                for v in tr:
                    seen.add(v[:3])
                keys = [(k, i) for i, v in enumerate(tr) for k in cand_keys(img, v[0], v[1], v[3])]
                has_relay = any(v[1] % R == 0 and (((v[3] >> W_) & 1 and v[0] > 0) or ((v[3] >> E_) & 1 and v[0] + 1 < w)) for v in tr)
                if not keys:
                    continue
                k0, i0 = min(keys)
                pts = [(v[0], v[1]) for v in tr[i0:] + tr[:i0]]
                out[k0] = (len(tr), pts, has_relay)
    return out


def check(mask, R, J=32):
    want = plain_borders(mask, R)
    got = relay_borders(mask, R, J)
    want_relay = {k: v[:2] for k, v in want.items() if v[2]}
    assert set(got) == set(want_relay), (sorted(set(got) ^ set(want_relay))[:5])
    for k, (n, pts) in got.items():
        assert (n, pts) == want_relay[k], f"border {k}: {n} vs {want_relay[k][0]} points"
    return len(got), len(want) - len(want_relay)


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(7)
    nb = nn = 0
    for t in range(trials):
        h, w = int(rng.integers(3, 40)), int(rng.integers(3, 70))
        kind = t % 4
        if kind == 0:
            m = rng.random((h, w)) < rng.uniform(0.2, 0.8)
        elif kind == 1:  # blobs
            m = np.zeros((h, w), bool)
            for _ in range(int(rng.integers(1, 6))):
                x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
                m[y0:y0 + int(rng.integers(1, h)), x0:x0 + int(rng.integers(1, w))] ^= True
        elif kind == 2:  # thin lines and rings
            m = np.zeros((h, w), bool)
            for _ in range(int(rng.integers(1, 5))):
                x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
                x1, y1 = min(w, x0 + int(rng.integers(1, w))), min(h, y0 + int(rng.integers(1, h)))
                m[y0:y1, x0] = True; m[y0:y1, x1 - 1] = True; m[y0, x0:x1] = True; m[y1 - 1, x0:x1] = True
        else:
            m = rng.random((h, w)) < 0.5
            m[:, 0] = rng.random(h) < 0.7  # pressure on the x == 0 / x == w - 1 rules
            m[:, -1] = rng.random(h) < 0.7
        for R in (int(rng.choice([1, 2, 3, 4, 8])), 32 if t % 16 else 5):
            a, b = check(m, R, int(rng.choice([1, 2, 3, 5, 32])))
            nb += a; nn += b
    print(f"relay_proto ok: {trials} masks, {nb} relay borders identical to the plain trace, {nn} borders without a relay left to the candidate walks")


if __name__ == "__main__":
    main()
