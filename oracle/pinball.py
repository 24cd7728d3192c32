"""Pinball domain, CPU oracle (NumPy, fp32, one rounding per operation).  TEST INFRASTRUCTURE.

What it restates: the Pinball domain used by the paper the reference's README names
(/root/reference/README.md:2).  The reference has no code (README.md:1-2 is all of it),
so the constants follow BASELINE.json north_star ("4-D ball state, 5 discrete thrust
actions, polygonal obstacles with elastic reflection and drag") and SURVEY.md appendix A.1.
Whatever this file computes is the definition the CUDA step kernel is checked against.

Arithmetic contract: every quantity is fp32 and every +,-,*,/,sqrt
is individually rounded (no fused multiply-add), in exactly the order written in
`step_scalar`.  The CUDA kernel reproduces the same operation order with __fmul_rn /
__fadd_rn so collision decisions, terminal flags and next states are bit-identical.

One env step (action a in {0: +x, 1: +y, 2: -x, 3: -y, 4: none}):
  impulse +-1/5 on one velocity component, clipped to [-1, 1]; then 20 substeps of
    pos += vel * (ball_r / 20)
    test every edge: closest point on the segment; hit iff dist^2 <= r^2 and the ball
      moves toward the closest point ((c - p) . v > 0)
    exactly one hit  -> elastic reflection v' = v - 2 (v.n) n about that edge
                        (on the last substep the ball moves once more with v')
    two or more hits -> v' = -v
    goal test: |pos - target|^2 < target_r^2 -> done, reward +10000, return at once
  then drag v *= 0.995, bounds clamp (x > 1 -> 0.95, x < 0 -> 0.05, same for y),
  reward -1 (no thrust) or -5 (thrust).
"""
import os

import numpy as np

f32 = np.float32

N_ACTIONS = 5
ACC_X, ACC_Y, DEC_X, DEC_Y, ACC_NONE = range(5)
N_SUBSTEPS = 20
DRAG = f32(0.995)
IMPULSE = f32(1.0) / f32(5.0)
STEP_PENALTY = f32(-1.0)
THRUST_PENALTY = f32(-5.0)
GOAL_REWARD = f32(10000.0)

# flags word layout (shared with include/scg_b200.h)
FLAG_DONE = 1
HIT_NONE, HIT_REFLECT, HIT_REVERSE = 0, 1, 2
FLAG_KIND_SHIFT = 1
FLAG_EDGE_SHIFT = 8
FLAG_OBST_SHIFT = 16

MAPS_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maps")


def pack_flags(done, kind, obstacle, edge):
    w = np.asarray(done, dtype=np.int32) & 1
    w = w | (np.asarray(kind, dtype=np.int32) << FLAG_KIND_SHIFT)
    has = np.asarray(kind) != HIT_NONE
    w = w | np.where(has, np.asarray(edge, dtype=np.int32) << FLAG_EDGE_SHIFT, 0)
    w = w | np.where(has, np.asarray(obstacle, dtype=np.int32) << FLAG_OBST_SHIFT, 0)
    return w.astype(np.int32)


def unpack_flags(flags):
    """flags int32 -> (done bool, kind, obstacle, edge); obstacle/edge are -1 when no collision."""
    flags = np.asarray(flags, dtype=np.int32)
    done = (flags & FLAG_DONE) != 0
    kind = (flags >> FLAG_KIND_SHIFT) & 3
    edge = np.where(kind != 0, (flags >> FLAG_EDGE_SHIFT) & 0xFF, -1)
    obst = np.where(kind != 0, (flags >> FLAG_OBST_SHIFT) & 0xFFF, -1)
    return done, kind, obst.astype(np.int32), edge.astype(np.int32)


class PinballMap:
    """Parsed map + the fp32 edge table both backends must agree on bit for bit.

    edge table columns: x1, y1, dx, dy, inv_len2, nx, ny, 0
      dx = x2 - x1, dy = y2 - y1, len2 = dx*dx + dy*dy, inv_len2 = 1/len2,
      len = sqrt(len2), nx = dy/len, ny = (0 - dx)/len          (all fp32, in this order)
    """

    def __init__(self, ball_r, target, starts, polygons):
        self.ball_r = f32(ball_r)
        self.target = tuple(f32(v) for v in target)
        self.starts = np.asarray(starts, dtype=np.float32).reshape(-1, 2)
        self.polygons = [np.asarray(p, dtype=np.float32).reshape(-1, 2) for p in polygons]
        if len(self.starts) == 0:
            raise ValueError("map has no start position")
        rows, obst, local = [], [], []
        for o, poly in enumerate(self.polygons):
            n = len(poly)
            if n < 3:
                raise ValueError("polygon with fewer than 3 vertices")
            for j in range(n):
                x1, y1 = poly[j]
                x2, y2 = poly[(j + 1) % n]
                dx = f32(x2 - x1)
                dy = f32(y2 - y1)
                len2 = f32(f32(dx * dx) + f32(dy * dy))
                if not len2 > 0:
                    raise ValueError("degenerate edge")
                inv = f32(f32(1.0) / len2)
                ln = f32(np.sqrt(len2))
                nx = f32(dy / ln)
                ny = f32(f32(f32(0.0) - dx) / ln)
                rows.append([x1, y1, dx, dy, inv, nx, ny, 0.0])
                obst.append(o)
                local.append(j)
        self.edges = np.asarray(rows, dtype=np.float32)
        self.edge_obstacle = np.asarray(obst, dtype=np.int32)
        self.edge_local = np.asarray(local, dtype=np.int32)
        self.n_edges = len(rows)
        self.r2 = f32(self.ball_r * self.ball_r)
        self.h = f32(self.ball_r / f32(N_SUBSTEPS))
        self.tr2 = f32(self.target[2] * self.target[2])

    @classmethod
    def from_file(cls, path):
        ball, target, starts, polys = None, None, [], []
        with open(path) as f:
            for line in f:
                tok = line.split("#")[0].split()
                if not tok:
                    continue
                v = [float(u) for u in tok[1:]]
                if tok[0] == "ball":
                    ball = v[0]
                elif tok[0] == "target":
                    target = v[:3]
                elif tok[0] == "start":
                    starts = list(zip(v[0::2], v[1::2]))
                elif tok[0] == "polygon":
                    polys.append(list(zip(v[0::2], v[1::2])))
                else:
                    raise ValueError(f"unknown map directive {tok[0]!r}")
        if ball is None or target is None:
            raise ValueError("map needs 'ball' and 'target' lines")
        return cls(ball, target, starts, polys)

    @classmethod
    def from_name(cls, name):
        return cls.from_file(os.path.join(MAPS_DIR, name + ".cfg"))

    # -- synthetic benchmark inputs ------------------------------------------------------------
    def in_free_space(self, x, y, clearance=None):
        """True where the ball centre (x, y) is outside every polygon and farther than
        `clearance` (default 1.05 ball radii) from every edge.  fp64 geometry; used only to
        generate inputs, never inside the step."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        c = float(self.ball_r) * 1.05 if clearance is None else clearance
        ok = (x > 0) & (x < 1) & (y > 0) & (y < 1)
        for poly in self.polygons:
            p = poly.astype(np.float64)
            n = len(p)
            inside = np.zeros(x.shape, dtype=bool)
            for i in range(n):
                x1, y1 = p[i]
                x2, y2 = p[(i + 1) % n]
                if y1 != y2:
                    cond = (y1 > y) != (y2 > y)
                    xi = (x2 - x1) * (y - y1) / (y2 - y1) + x1
                    inside ^= cond & (x < xi)
                dx, dy = x2 - x1, y2 - y1
                t = np.clip(((x - x1) * dx + (y - y1) * dy) / (dx * dx + dy * dy), 0, 1)
                ok &= np.hypot(x1 + t * dx - x, y1 + t * dy - y) > c
            ok &= ~inside
        return ok

    def sample_free_states(self, rng, n, vmax=1.0):
        """n synthetic states: position uniform in free space, velocity uniform in [-vmax, vmax]^2."""
        out = np.empty((n, 4), dtype=np.float32)
        k = 0
        while k < n:
            m = max(2 * (n - k), 64)
            x = rng.uniform(0.0, 1.0, m)
            y = rng.uniform(0.0, 1.0, m)
            keep = self.in_free_space(x, y)
            x, y = x[keep][: n - k], y[keep][: n - k]
            out[k:k + len(x), 0] = x
            out[k:k + len(x), 1] = y
            k += len(x)
        out[:, 2:] = rng.uniform(-vmax, vmax, (n, 2)).astype(np.float32)
        return out


def _clip1(v):
    return min(max(v, f32(-1.0)), f32(1.0))


def step_scalar(m, state, action):
    """One env step for ONE env with explicit fp32 scalar arithmetic (the normative form).

    state: 4 floats; returns (next_state float32[4], reward float32, flags int32).
    """
    x, y, vx, vy = (f32(v) for v in state)
    a = int(action)
    if a == ACC_X:
        vx = _clip1(f32(vx + IMPULSE))
    elif a == ACC_Y:
        vy = _clip1(f32(vy + IMPULSE))
    elif a == DEC_X:
        vx = _clip1(f32(vx - IMPULSE))
    elif a == DEC_Y:
        vy = _clip1(f32(vy - IMPULSE))
    elif a != ACC_NONE:
        raise ValueError("action out of range")
    E = m.edges
    tx, ty, _ = m.target
    kind, h_obst, h_edge = HIT_NONE, 0, 0
    done = False
    for i in range(N_SUBSTEPS):
        x = f32(x + f32(vx * m.h))
        y = f32(y + f32(vy * m.h))
        nhit, first = 0, -1
        for e in range(m.n_edges):
            x1, y1, dx, dy, inv, _, _, _ = E[e]
            rx = f32(x - x1)
            ry = f32(y - y1)
            t = f32(f32(f32(rx * dx) + f32(ry * dy)) * inv)
            t = min(max(t, f32(0.0)), f32(1.0))
            cx = f32(x1 + f32(t * dx))
            cy = f32(y1 + f32(t * dy))
            ex = f32(cx - x)
            ey = f32(cy - y)
            d2 = f32(f32(ex * ex) + f32(ey * ey))
            if d2 <= m.r2 and f32(f32(ex * vx) + f32(ey * vy)) > 0:
                nhit += 1
                if first < 0:
                    first = e
        if nhit == 1:
            nx, ny = E[first, 5], E[first, 6]
            k = f32(f32(2.0) * f32(f32(vx * nx) + f32(vy * ny)))
            vx = f32(vx - f32(k * nx))
            vy = f32(vy - f32(k * ny))
            kind = HIT_REFLECT
            if i == N_SUBSTEPS - 1:
                x = f32(x + f32(vx * m.h))
                y = f32(y + f32(vy * m.h))
        elif nhit >= 2:
            vx = f32(-vx)
            vy = f32(-vy)
            kind = HIT_REVERSE
        if nhit >= 1:
            h_obst, h_edge = int(m.edge_obstacle[first]), int(m.edge_local[first])
        gx = f32(x - tx)
        gy = f32(y - ty)
        if f32(f32(gx * gx) + f32(gy * gy)) < m.tr2:
            done = True
            break
    if done:
        reward = GOAL_REWARD
    else:
        vx = f32(vx * DRAG)
        vy = f32(vy * DRAG)
        if x > 1:
            x = f32(0.95)
        if x < 0:
            x = f32(0.05)
        if y > 1:
            y = f32(0.95)
        if y < 0:
            y = f32(0.05)
        reward = STEP_PENALTY if a == ACC_NONE else THRUST_PENALTY
    flags = pack_flags(done, kind, h_obst, h_edge)
    return np.array([x, y, vx, vy], dtype=np.float32), f32(reward), np.int32(flags)


def step_batched(m, state, actions, chunk=8192):
    """Vectorised form of `step_scalar` over a batch (same operations, same order, per element).

    state float32 (B, 4), actions int (B,) -> (next_state (B, 4), reward (B,), flags int32 (B,)).
    """
    state = np.ascontiguousarray(state, dtype=np.float32)
    actions = np.asarray(actions)
    B = state.shape[0]
    if actions.shape != (B,):
        raise ValueError("actions must have shape (B,)")
    if B and (actions.min() < 0 or actions.max() >= N_ACTIONS):
        raise ValueError("action out of range")
    out = np.empty_like(state)
    reward = np.empty(B, dtype=np.float32)
    flags = np.empty(B, dtype=np.int32)
    for lo in range(0, B, chunk):
        hi = min(lo + chunk, B)
        out[lo:hi], reward[lo:hi], flags[lo:hi] = _step_chunk(m, state[lo:hi], actions[lo:hi])
    return out, reward, flags


def _step_chunk(m, state, a):
    x, y, vx, vy = (state[:, i].copy() for i in range(4))
    one = f32(1.0)
    vx = np.where(a == ACC_X, np.clip(vx + IMPULSE, -one, one), vx)
    vy = np.where(a == ACC_Y, np.clip(vy + IMPULSE, -one, one), vy)
    vx = np.where(a == DEC_X, np.clip(vx - IMPULSE, -one, one), vx)
    vy = np.where(a == DEC_Y, np.clip(vy - IMPULSE, -one, one), vy)
    E = m.edges
    X1, Y1, DX, DY, INV, NX, NY = (E[None, :, i] for i in range(7))
    tx, ty, _ = m.target
    n = len(x)
    kind = np.zeros(n, dtype=np.int32)
    h_first = np.zeros(n, dtype=np.int32)
    done = np.zeros(n, dtype=bool)
    idx = np.arange(n)
    for i in range(N_SUBSTEPS):
        if len(idx) == 0:
            break
        px = x[idx] + vx[idx] * m.h
        py = y[idx] + vy[idx] * m.h
        qx, qy = vx[idx], vy[idx]
        rx = px[:, None] - X1
        ry = py[:, None] - Y1
        t = (rx * DX + ry * DY) * INV
        t = np.minimum(np.maximum(t, f32(0.0)), f32(1.0))
        ex = (X1 + t * DX) - px[:, None]
        ey = (Y1 + t * DY) - py[:, None]
        d2 = ex * ex + ey * ey
        hit = (d2 <= m.r2) & ((ex * qx[:, None] + ey * qy[:, None]) > 0)
        nhit = hit.sum(axis=1)
        first = hit.argmax(axis=1)
        one_hit = nhit == 1
        multi = nhit >= 2
        nx, ny = E[first, 5], E[first, 6]
        k = f32(2.0) * (qx * nx + qy * ny)
        rvx = qx - k * nx
        rvy = qy - k * ny
        qx = np.where(one_hit, rvx, np.where(multi, -qx, qx))
        qy = np.where(one_hit, rvy, np.where(multi, -qy, qy))
        if i == N_SUBSTEPS - 1:
            px = np.where(one_hit, px + qx * m.h, px)
            py = np.where(one_hit, py + qy * m.h, py)
        any_hit = nhit >= 1
        kind[idx] = np.where(one_hit, HIT_REFLECT, np.where(multi, HIT_REVERSE, kind[idx]))
        h_first[idx] = np.where(any_hit, first, h_first[idx])
        x[idx], y[idx], vx[idx], vy[idx] = px, py, qx, qy
        gx = px - tx
        gy = py - ty
        goal = (gx * gx + gy * gy) < m.tr2
        done[idx[goal]] = True
        idx = idx[~goal]
    live = ~done
    vx = np.where(live, vx * DRAG, vx)
    vy = np.where(live, vy * DRAG, vy)
    x = np.where(live & (x > 1), f32(0.95), x)
    x = np.where(live & (x < 0), f32(0.05), x)
    y = np.where(live & (y > 1), f32(0.95), y)
    y = np.where(live & (y < 0), f32(0.05), y)
    reward = np.where(done, GOAL_REWARD, np.where(a == ACC_NONE, STEP_PENALTY, THRUST_PENALTY)).astype(np.float32)
    flags = pack_flags(done, kind, m.edge_obstacle[h_first], m.edge_local[h_first])
    return np.stack([x, y, vx, vy], axis=1).astype(np.float32), reward, flags


class PinballEnv:
    """Batched Pinball environment: the API the GPU backend is a drop-in for.

        env.reset(mask=None, states=None) -> state  float32 (B, 4)
        env.step(actions) -> (state, reward, done, hit_info)
            hit_info: int32 (B, 3) = (kind, obstacle, edge) of the last collision in the step,
            kind 0 none / 1 reflection / 2 reversal; obstacle and edge are -1 when kind == 0.
    Finished envs are not reset automatically; call reset(mask=done).
    """

    def __init__(self, pmap, batch=1, seed=0, scalar=False, env_offset=0):
        self.map = pmap if isinstance(pmap, PinballMap) else PinballMap.from_name(pmap)
        self.batch = int(batch)
        self.seed = int(seed)
        self.env_offset = int(env_offset)
        self.scalar = bool(scalar)
        self.state = np.zeros((self.batch, 4), dtype=np.float32)
        self.n_resets = 0
        self.last_flags = np.zeros(self.batch, dtype=np.int32)
        self.reset()

    def reset(self, mask=None, states=None, step=None):
        """Put the masked envs (default: all) at a start position with zero velocity, or at the
        given `states`.  The start is picked by Philox(seed; env id, step, STREAM_RESET); `step`
        defaults to the number of reset() calls so far."""
        from .philox import draws, uniform01, STREAM_RESET
        if mask is None:
            mask = np.ones(self.batch, dtype=bool)
        mask = np.asarray(mask, dtype=bool)
        if mask.shape != (self.batch,):
            raise ValueError("mask must have shape (B,)")
        if states is not None:
            states = np.asarray(states, dtype=np.float32).reshape(self.batch, 4)
            self.state[mask] = states[mask]
        else:
            ids = np.nonzero(mask)[0]
            key = self.n_resets if step is None else int(step)
            u = uniform01(draws(self.seed, ids + self.env_offset, key, STREAM_RESET)[:, 0])
            ns = len(self.map.starts)
            pick = np.minimum((u * f32(ns)).astype(np.int32), ns - 1)
            self.state[ids, 0:2] = self.map.starts[pick]
            self.state[ids, 2:4] = 0
        self.n_resets += 1
        return self.state.copy()

    def step(self, actions):
        actions = np.asarray(actions)
        if self.scalar:
            ns = np.empty_like(self.state)
            r = np.empty(self.batch, dtype=np.float32)
            fl = np.empty(self.batch, dtype=np.int32)
            for b in range(self.batch):
                ns[b], r[b], fl[b] = step_scalar(self.map, self.state[b], actions[b])
        else:
            ns, r, fl = step_batched(self.map, self.state, actions)
        self.state = ns
        self.last_flags = fl
        done, kind, obst, edge = unpack_flags(fl)
        return ns.copy(), r, done, np.stack([kind, obst, edge], axis=1).astype(np.int32)
