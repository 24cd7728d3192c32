#!/usr/bin/env python
"""Author and validate the Pinball map fixtures under maps/.

The reference repository ships no map files, so the
`easy` and `hard` maps are authored here in the public Pinball text format:

    ball <r>
    target <x> <y> <r>
    start <x> <y> [<x> <y> ...]
    polygon <x1> <y1> <x2> <y2> ...      (closed implicitly)

`easy` is written by hand (36 edges).  `hard` is produced by the seeded
generator below (about 100 edges).  Both are validated: start and target in
free space, and the target reachable from every start through cells that are
at least one ball radius away from every obstacle.

Usage:  python tools/make_maps.py            (rewrites maps/*.cfg)
        python tools/make_maps.py --check    (validates the committed files)
"""
import argparse
import math
import os
import sys
from collections import deque

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAPS = os.path.join(ROOT, "maps")

WALLS = [
    [(0.0, 0.0), (1.0, 0.0), (1.0, 0.01), (0.0, 0.01)],      # bottom
    [(0.0, 0.99), (1.0, 0.99), (1.0, 1.0), (0.0, 1.0)],      # top
    [(0.0, 0.0), (0.01, 0.0), (0.01, 1.0), (0.0, 1.0)],      # left
    [(0.99, 0.0), (1.0, 0.0), (1.0, 1.0), (0.99, 1.0)],      # right
]

EASY = dict(
    ball=0.02,
    target=(0.9, 0.2, 0.04),
    starts=[(0.2, 0.9)],
    polygons=WALLS + [
        [(0.30, 0.55), (0.42, 0.45), (0.50, 0.60), (0.45, 0.75)],
        [(0.60, 0.40), (0.99, 0.42), (0.99, 0.48), (0.62, 0.47)],
        [(0.15, 0.20), (0.30, 0.15), (0.38, 0.28), (0.28, 0.38), (0.14, 0.32)],
        [(0.65, 0.75), (0.78, 0.65), (0.80, 0.85)],
        [(0.55, 0.01), (0.62, 0.01), (0.60, 0.25), (0.54, 0.22)],
    ],
)


def seg_point_dist(px, py, x1, y1, x2, y2):
    dx, dy = x2 - x1, y2 - y1
    t = ((px - x1) * dx + (py - y1) * dy) / (dx * dx + dy * dy)
    t = np.clip(t, 0.0, 1.0)
    return np.hypot(x1 + t * dx - px, y1 + t * dy - py)


def point_in_poly(px, py, poly):
    inside = np.zeros_like(px, dtype=bool)
    n = len(poly)
    for i in range(n):
        x1, y1 = poly[i]
        x2, y2 = poly[(i + 1) % n]
        cond = (y1 > py) != (y2 > py)
        with np.errstate(divide="ignore", invalid="ignore"):
            xi = (x2 - x1) * (py - y1) / (y2 - y1) + x1
        inside ^= cond & (px < xi)
    return inside


def clearance_grid(m, n=200):
    """Boolean n x n grid: cell centre is >= 1.25 ball radii from every obstacle."""
    c = (np.arange(n) + 0.5) / n
    px, py = np.meshgrid(c, c, indexing="ij")
    free = np.ones((n, n), dtype=bool)
    for poly in m["polygons"]:
        free &= ~point_in_poly(px, py, poly)
        k = len(poly)
        for i in range(k):
            x1, y1 = poly[i]
            x2, y2 = poly[(i + 1) % k]
            free &= seg_point_dist(px, py, x1, y1, x2, y2) > 1.25 * m["ball"]
    return free


def validate(m, name):
    n = 200
    free = clearance_grid(m, n)

    def cell(x, y):
        return min(int(x * n), n - 1), min(int(y * n), n - 1)

    tx, ty, tr = m["target"]
    assert free[cell(tx, ty)], f"{name}: target not in free space"
    seen = np.zeros_like(free)
    q = deque([cell(tx, ty)])
    seen[cell(tx, ty)] = True
    while q:
        i, j = q.popleft()
        for di, dj in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            a, b = i + di, j + dj
            if 0 <= a < n and 0 <= b < n and free[a, b] and not seen[a, b]:
                seen[a, b] = True
                q.append((a, b))
    for sx, sy in m["starts"]:
        assert free[cell(sx, sy)], f"{name}: start {sx, sy} not in free space"
        assert seen[cell(sx, sy)], f"{name}: target unreachable from start {sx, sy}"
    for poly in m["polygons"]:
        k = len(poly)
        assert k >= 3
        for i in range(k):
            assert poly[i] != poly[(i + 1) % k], f"{name}: degenerate edge"
    n_edges = sum(len(p) for p in m["polygons"])
    reach = seen.sum() / free.sum()
    print(f"{name}: {len(m['polygons'])} polygons, {n_edges} edges, "
          f"free {free.mean():.3f}, reachable-from-target {reach:.3f}")
    return n_edges


def make_hard(seed=7):
    """Seeded generator: irregular star-shaped polygons on a jittered 4x4 lattice."""
    rng = np.random.RandomState(seed)
    m = dict(ball=0.02, target=(0.9, 0.2, 0.04), starts=[(0.2, 0.9)], polygons=list(WALLS))
    tx, ty, _ = m["target"]
    sx, sy = m["starts"][0]
    centres = []
    for i in range(4):
        for j in range(4):
            cx = 0.14 + 0.24 * i + rng.uniform(-0.035, 0.035)
            cy = 0.14 + 0.24 * j + rng.uniform(-0.035, 0.035)
            if math.hypot(cx - tx, cy - ty) < 0.16 or math.hypot(cx - sx, cy - sy) < 0.16:
                continue
            centres.append((cx, cy))
    for cx, cy in centres:
        k = int(rng.randint(5, 8))
        ang = np.sort(rng.uniform(0, 2 * math.pi, k))
        # keep the polygon star-shaped and not too thin: spread angles a little
        ang = 0.5 * ang + 0.5 * (np.arange(k) * 2 * math.pi / k + rng.uniform(0, 2 * math.pi))
        ang = np.sort(np.mod(ang, 2 * math.pi))
        rad = rng.uniform(0.045, 0.085, k)
        poly = [(round(float(cx + r * math.cos(a)), 4), round(float(cy + r * math.sin(a)), 4))
                for a, r in zip(ang, rad)]
        poly = [(min(max(x, 0.03), 0.97), min(max(y, 0.03), 0.97)) for x, y in poly]
        m["polygons"].append(poly)
    return m


def write_cfg(m, path, header):
    with open(path, "w") as f:
        for line in header:
            f.write(f"# {line}\n")
        f.write(f"ball {m['ball']}\n")
        f.write("target {} {} {}\n".format(*m["target"]))
        f.write("start " + " ".join(f"{x} {y}" for x, y in m["starts"]) + "\n")
        for poly in m["polygons"]:
            f.write("polygon " + " ".join(f"{x} {y}" for x, y in poly) + "\n")


def read_cfg(path):
    m = dict(polygons=[], starts=[])
    with open(path) as f:
        for line in f:
            t = line.split("#")[0].split()
            if not t:
                continue
            v = [float(u) for u in t[1:]]
            if t[0] == "ball":
                m["ball"] = v[0]
            elif t[0] == "target":
                m["target"] = tuple(v)
            elif t[0] == "start":
                m["starts"] = list(zip(v[0::2], v[1::2]))
            elif t[0] == "polygon":
                m["polygons"].append(list(zip(v[0::2], v[1::2])))
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    os.makedirs(MAPS, exist_ok=True)
    if args.check:
        for name in ("easy", "hard"):
            validate(read_cfg(os.path.join(MAPS, name + ".cfg")), name)
        return 0
    validate(EASY, "easy")
    write_cfg(EASY, os.path.join(MAPS, "easy.cfg"),
              ["Pinball 'easy' map - authored for this repository (the reference ships no maps).",
               "Public Pinball text format; see tools/make_maps.py."])
    hard = make_hard()
    validate(hard, "hard")
    write_cfg(hard, os.path.join(MAPS, "hard.cfg"),
              ["Pinball 'hard' map - generated by tools/make_maps.py (seed 7); authored here.",
               "Public Pinball text format."])
    return 0


if __name__ == "__main__":
    sys.exit(main())
