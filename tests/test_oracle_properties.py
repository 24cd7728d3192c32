"""Property tests of the Pinball oracle (hypothesis): invariants that hold for any state and action."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle
from oracle.pinball import step_scalar, step_batched, unpack_flags

EASY = oracle.PinballMap.from_name("easy")
HARD = oracle.PinballMap.from_name("hard")
coord = st.floats(min_value=0.03125, max_value=0.96875, allow_nan=False, width=32)
vel = st.floats(min_value=-1.0, max_value=1.0, allow_nan=False, width=32)


@settings(max_examples=150, deadline=None)
@given(x=coord, y=coord, vx=vel, vy=vel, a=st.integers(0, 4), hard=st.booleans())
def test_step_invariants(x, y, vx, vy, a, hard):
    m = HARD if hard else EASY
    s = np.array([x, y, vx, vy], dtype=np.float32)
    ns, r, fl = step_scalar(m, s, a)
    done, kind, obst, edge = unpack_flags(fl)
    assert np.all(np.isfinite(ns)) and 0.0 <= ns[0] <= 1.0 and 0.0 <= ns[1] <= 1.0
    assert float(r) in (-1.0, -5.0, 10000.0) and (float(r) == 10000.0) == bool(done)
    # elastic collisions never add energy: the speed is bounded by the thrust-adjusted start speed (then drag)
    v0 = np.array([vx, vy], dtype=np.float64)
    if a < 4:
        v0[a % 2] = np.clip(v0[a % 2] + (0.2 if a < 2 else -0.2), -1, 1)
    assert np.hypot(ns[2], ns[3]) <= np.hypot(*v0) * (1.0 if done else 0.995) + 1e-5
    if kind != 0:
        assert 0 <= obst < len(m.polygons) and 0 <= edge < len(m.polygons[obst])
    # the vectorised form agrees bit for bit
    bs, br, bf = step_batched(m, s[None], np.array([a]))
    assert np.array_equal(bs[0].view(np.uint32), ns.view(np.uint32)) and bf[0] == fl and br[0] == r


@settings(max_examples=50, deadline=None)
@given(x=coord, y=coord, vx=vel, vy=vel)
def test_fourier_features_are_bounded_and_even_in_reflection(x, y, vx, vy):
    fb = oracle.FourierBasis(3)
    s = np.array([[x, y, vx, vy]], dtype=np.float32)
    phi = fb.features(s)[0]
    assert phi[0] == 1.0 and np.abs(phi).max() <= 1.0
    # cos(pi c . s_hat) with s_hat -> 1 - s_hat picks up the sign (-1)^(sum c)
    mirror = np.array([[1 - x, 1 - y, -vx, -vy]], dtype=np.float32)
    assert np.allclose(fb.features(mirror)[0], phi * (-1.0) ** fb.C.sum(axis=1), atol=2e-5)
