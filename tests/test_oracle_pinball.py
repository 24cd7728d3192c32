"""Analytic known-answer tests of the Pinball oracle.  The reference has no
tests or golden vectors (/root/reference/README.md:1-2 is the whole repository), so these hand-derivable
cases are what pins the oracle; tests/golden/ then pins it against regressions."""
import numpy as np
import pytest

import oracle
from oracle.pinball import (PinballMap, PinballEnv, step_scalar, step_batched, unpack_flags, pack_flags, N_SUBSTEPS,
                            HIT_NONE, HIT_REFLECT, HIT_REVERSE)

f32 = np.float32
WALLS = [[(0.0, 0.0), (1.0, 0.0), (1.0, 0.01), (0.0, 0.01)], [(0.0, 0.99), (1.0, 0.99), (1.0, 1.0), (0.0, 1.0)],
         [(0.0, 0.0), (0.01, 0.0), (0.01, 1.0), (0.0, 1.0)], [(0.99, 0.0), (1.0, 0.0), (1.0, 1.0), (0.99, 1.0)]]


def box_map(extra=(), target=(0.9, 0.2, 0.04)):
    return PinballMap(0.02, target, [(0.2, 0.9)], WALLS + list(extra))


def test_free_flight_is_twenty_sequential_fp32_adds_and_drag():
    m = box_map()
    s = np.array([0.5, 0.5, 0.3, -0.2], dtype=np.float32)
    ns, r, fl = step_scalar(m, s, 4)
    x, y = f32(0.5), f32(0.5)
    h = f32(f32(0.02) / f32(20))
    for _ in range(N_SUBSTEPS):
        x = f32(x + f32(f32(0.3) * h))
        y = f32(y + f32(f32(-0.2) * h))
    assert ns[0] == x and ns[1] == y
    assert ns[2] == f32(f32(0.3) * f32(0.995)) and ns[3] == f32(f32(-0.2) * f32(0.995))
    assert r == f32(-1.0) and fl == 0
    assert abs(float(ns[0]) - (0.5 + 0.3 * 0.02)) < 1e-6


def test_thrust_impulse_clip_and_reward():
    m = box_map()
    rest = np.array([0.5, 0.5, 0.0, 0.0], dtype=np.float32)
    for a, (ix, sign) in enumerate([(2, 1), (3, 1), (2, -1), (3, -1)]):
        ns, r, fl = step_scalar(m, rest, a)
        assert ns[ix] == f32(f32(sign) * f32(f32(1.0) / f32(5.0)) * f32(0.995)) and r == f32(-5.0)
    fast = np.array([0.5, 0.5, 0.95, -0.95], dtype=np.float32)
    assert step_scalar(m, fast, 0)[0][2] == f32(f32(1.0) * f32(0.995))        # clipped at +1
    assert step_scalar(m, fast, 3)[0][3] == f32(f32(-1.0) * f32(0.995))       # clipped at -1
    with pytest.raises(ValueError):
        step_scalar(m, rest, 5)


def test_wall_reflection_flips_normal_component_and_reports_indices():
    # a square obstacle (polygon 4) whose left face is x = 0.6; ball moving +x into it
    m = box_map([[(0.6, 0.3), (0.8, 0.3), (0.8, 0.7), (0.6, 0.7)]])
    s = np.array([0.575, 0.5, 0.5, 0.1], dtype=np.float32)
    ns, r, fl = step_scalar(m, s, 4)
    done, kind, obst, edge = unpack_flags(fl)
    assert not done and kind == HIT_REFLECT and obst == 4 and edge == 3       # edge 3: (0.6,0.7)->(0.6,0.3)
    assert ns[2] == f32(f32(-0.5) * f32(0.995)) and ns[3] == f32(f32(0.1) * f32(0.995))
    speed0, speed1 = np.hypot(0.5, 0.1) * 0.995, np.hypot(float(ns[2]), float(ns[3]))
    assert abs(speed0 - speed1) < 1e-6
    assert ns[0] < f32(0.6) - m.ball_r + f32(0.01)


def test_moving_away_from_an_overlapping_edge_is_not_a_collision():
    m = box_map([[(0.6, 0.3), (0.8, 0.3), (0.8, 0.7), (0.6, 0.7)]])
    s = np.array([0.59, 0.5, -0.5, 0.0], dtype=np.float32)                    # overlaps the face, leaving
    ns, r, fl = step_scalar(m, s, 4)
    assert unpack_flags(fl)[1] == HIT_NONE and ns[2] == f32(f32(-0.5) * f32(0.995))


def test_corner_two_edges_reverses_velocity():
    m = box_map([[(0.6, 0.3), (0.8, 0.3), (0.8, 0.7), (0.6, 0.7)]])
    # aim at the corner (0.6, 0.3) along the diagonal: both adjacent edges are within the radius
    s = np.array([0.6 - 0.02, 0.3 - 0.02, 0.7, 0.7], dtype=np.float32)
    ns, r, fl = step_scalar(m, s, 4)
    done, kind, obst, edge = unpack_flags(fl)
    assert kind == HIT_REVERSE and obst == 4
    assert ns[2] == f32(f32(-0.7) * f32(0.995)) and ns[3] == f32(f32(-0.7) * f32(0.995))


def test_two_obstacles_at_once_reverse():
    m = box_map([[(0.5, 0.3), (0.52, 0.3), (0.52, 0.7), (0.5, 0.7)], [(0.4, 0.52), (0.6, 0.52), (0.6, 0.54), (0.4, 0.54)]])
    s = np.array([0.485, 0.505, 0.6, 0.6], dtype=np.float32)                  # into the inner corner of a cross
    ns, r, fl = step_scalar(m, s, 4)
    assert unpack_flags(fl)[1] == HIT_REVERSE


def test_goal_ends_step_at_once_with_reward():
    m = box_map()
    s = np.array([0.9 - 0.045, 0.2, 1.0, 0.0], dtype=np.float32)
    ns, r, fl = step_scalar(m, s, 4)
    done = unpack_flags(fl)[0]
    assert done and r == f32(10000.0)
    assert ns[2] == f32(1.0)                                                  # no drag after the goal
    assert (ns[0] - f32(0.9)) ** 2 + (ns[1] - f32(0.2)) ** 2 < f32(0.04) ** 2
    assert ns[0] < f32(0.9 - 0.045) + f32(0.02)                               # stopped before 20 substeps
    inside = np.array([0.9, 0.2, 0.0, 0.0], dtype=np.float32)
    assert unpack_flags(step_scalar(m, inside, 4)[2])[0]


def test_bounds_clamp():
    m = PinballMap(0.02, (0.9, 0.2, 0.04), [(0.2, 0.9)], [[(0.4, 0.4), (0.5, 0.4), (0.45, 0.5)]])   # no border walls
    for s, want in [([0.999, 0.5, 1.0, 0.0], (0, 0.95)), ([0.001, 0.5, -1.0, 0.0], (0, 0.05)),
                    ([0.5, 0.999, 0.0, 1.0], (1, 0.95)), ([0.5, 0.001, 0.0, -1.0], (1, 0.05))]:
        ns, _, _ = step_scalar(m, np.array(s, dtype=np.float32), 4)
        assert ns[want[0]] == f32(want[1])


def test_flags_roundtrip():
    fl = pack_flags([True, False, False], [HIT_REVERSE, HIT_NONE, HIT_REFLECT], [17, 5, 4095], [3, 9, 255])
    done, kind, obst, edge = unpack_flags(fl)
    assert done.tolist() == [True, False, False] and kind.tolist() == [2, 0, 1]
    assert obst.tolist() == [17, -1, 4095] and edge.tolist() == [3, -1, 255]


@pytest.mark.parametrize("name", ["easy", "hard"])
def test_batched_equals_scalar_bit_for_bit(name):
    m = PinballMap.from_name(name)
    rng = np.random.default_rng(3)
    S = m.sample_free_states(rng, 300)
    e = m.edges[rng.integers(0, m.n_edges, 150)]                              # half of them next to an edge
    t = rng.uniform(0, 1, 150).astype(np.float32)
    off = rng.uniform(-1.5, 1.5, 150).astype(np.float32) * m.ball_r
    S[:150, 0] = e[:, 0] + t * e[:, 2] + off * e[:, 5]
    S[:150, 1] = e[:, 1] + t * e[:, 3] + off * e[:, 6]
    A = rng.integers(0, 5, 300)
    ns, r, fl = step_batched(m, S, A, chunk=128)
    hits = 0
    for b in range(300):
        s1, r1, f1 = step_scalar(m, S[b], A[b])
        assert np.array_equal(s1.view(np.uint32), ns[b].view(np.uint32)) and r1 == r[b] and f1 == fl[b]
        hits += unpack_flags(f1)[1] != 0
    assert hits > 30


def test_maps_parse_and_are_closed_world():
    for name, n_poly in (("easy", 9), ("hard", 18)):
        m = PinballMap.from_name(name)
        assert len(m.polygons) == n_poly and m.n_edges == sum(len(p) for p in m.polygons)
        assert m.in_free_space(*m.starts[0]) and m.in_free_space(float(m.target[0]), float(m.target[1]))
        assert np.all(np.abs(np.hypot(m.edges[:, 5], m.edges[:, 6]) - 1) < 1e-6)          # unit normals
        assert np.all(np.abs(m.edges[:, 2] * m.edges[:, 5] + m.edges[:, 3] * m.edges[:, 6]) < 1e-7)


def test_env_api_reset_step_and_empty_batch():
    env = PinballEnv("easy", batch=5, seed=1)
    s0 = env.reset()
    assert s0.shape == (5, 4) and np.all(s0[:, :2] == env.map.starts[0]) and np.all(s0[:, 2:] == 0)
    s1, r, done, hit = env.step(np.array([0, 1, 2, 3, 4]))
    assert s1.shape == (5, 4) and r.tolist() == [-5, -5, -5, -5, -1] and hit.shape == (5, 3) and not done.any()
    env.reset(mask=np.array([True, False, False, False, False]))
    assert np.all(env.state[0, 2:] == 0) and np.all(env.state[1:] == s1[1:])
    with pytest.raises(ValueError):
        env.step(np.array([0, 1]))
    with pytest.raises(ValueError):
        env.step(np.array([0, 1, 2, 3, 7]))
    e0 = PinballEnv("easy", batch=0)
    s, r, d, h = e0.step(np.zeros(0, dtype=np.int32))
    assert s.shape == (0, 4) and r.shape == (0,)


def test_scalar_env_mode_matches_batched():
    a = PinballEnv("hard", batch=16, seed=2, scalar=True)
    b = PinballEnv("hard", batch=16, seed=2, scalar=False)
    rng = np.random.default_rng(0)
    S = a.map.sample_free_states(rng, 16)
    a.reset(states=S); b.reset(states=S)
    for _ in range(5):
        A = rng.integers(0, 5, 16)
        sa = a.step(A); sb = b.step(A)
        assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1]) and np.array_equal(sa[3], sb[3])
