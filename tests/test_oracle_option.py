"""Known-answer tests of the Fourier basis, per-option Sarsa(lambda) and logistic classifiers of the
oracle: hand-computed small cases and form-equivalence properties."""
import numpy as np
import pytest

import oracle
from oracle.option import OptionSet, Option, epsilon_greedy, logistic_features, sigmoid

f32 = np.float32


# ---- Fourier basis -------------------------------------------------------------------------------
@pytest.mark.parametrize("order", [1, 2, 3, 5])
def test_fourier_known_answers(order):
    fb = oracle.FourierBasis(order)
    n1 = order + 1
    assert fb.n_features == n1 ** 4 and fb.C.shape == (n1 ** 4, 4)
    # lexicographic, c0 slowest
    assert fb.C[0].tolist() == [0, 0, 0, 0] and fb.C[1].tolist() == [0, 0, 0, 1]
    assert fb.C[n1].tolist() == [0, 0, 1, 0] and fb.C[n1 ** 3].tolist() == [1, 0, 0, 0]
    assert fb.C[-1].tolist() == [order] * 4
    # state whose normalised form is 0 (x = y = 0, v = -2): every feature is 1
    zero = np.array([[0.0, 0.0, -2.0, -2.0]], dtype=np.float32)
    assert np.all(fb.features(zero) == 1.0)
    # normalised state 1 (x = y = 1, v = +2): phi_i = (-1)^(sum c_i)
    one = np.array([[1.0, 1.0, 2.0, 2.0]], dtype=np.float32)
    assert np.allclose(fb.features(one)[0], (-1.0) ** fb.C.sum(axis=1), atol=1e-6)
    rng = np.random.default_rng(order)
    S = rng.uniform(0, 1, (50, 4)).astype(np.float32)
    phi = fb.features(S)
    assert np.all(phi[:, 0] == 1.0) and np.abs(phi).max() <= 1.0
    # feature with c = (0, 1, 0, 0) is cos(pi y)
    assert np.allclose(phi[:, n1 ** 2], np.cos(np.pi * S[:, 1].astype(np.float64)), atol=1e-6)
    # step-size scale 1/||c||, and 1 for c = 0
    assert fb.alpha_scale[0] == 1.0 and fb.alpha_scale[1] == 1.0
    assert np.isclose(fb.alpha_scale[-1], 1.0 / (2.0 * order))
    with pytest.raises(ValueError):
        oracle.FourierBasis(3, n_dims=2)


def test_velocity_normalisation_range():
    s = oracle.FourierBasis.normalise(np.array([[0.3, 0.7, -2.0, 2.0], [0.3, 0.7, 0.0, 1.0]], dtype=np.float32))
    assert s[0].tolist() == [f32(0.3), f32(0.7), 0.0, 1.0] and s[1, 2] == 0.5 and s[1, 3] == 0.75


# ---- epsilon-greedy ------------------------------------------------------------------------------
def test_epsilon_greedy_ties_and_exploration():
    Q = np.array([[1.0, 3.0, 3.0, 0.0, 2.0], [0.0, 0.0, 0.0, 0.0, 0.0]], dtype=np.float32)
    ids = np.array([0, 1], dtype=np.uint32)
    assert epsilon_greedy(Q, 0.0, 1, ids, 0).tolist() == [1, 0]               # first maximum
    Qb = np.tile(Q[:1], (4000, 1))
    a = epsilon_greedy(Qb, 1.0, 7, np.arange(4000, dtype=np.uint32), 3)       # always explore
    assert set(a.tolist()) == {0, 1, 2, 3, 4} and abs((a == 4).mean() - 0.2) < 0.03
    a = epsilon_greedy(Qb, 0.25, 7, np.arange(4000, dtype=np.uint32), 3)
    assert abs((a != 1).mean() - 0.25 * 0.8) < 0.03


# ---- Sarsa(lambda) -------------------------------------------------------------------------------
def _transition(rng, B):
    S = rng.uniform(0.05, 0.95, (B, 4)).astype(np.float32)
    S[:, 2:] = rng.uniform(-1, 1, (B, 2))
    return S


def test_three_step_trace_by_hand():
    """B = 1, order 1: traces, TD errors and the weight update computed directly from the definitions."""
    g, lam, alpha = 0.9, 0.5, 0.1
    o = OptionSet(1, 1, 1, gamma=g, lam=lam, alpha=alpha)
    rng = np.random.default_rng(0)
    o.W[:] = rng.standard_normal(o.W.shape).astype(np.float32) * 0.1
    W = o.W[0].astype(np.float64).copy()
    fb = o.basis
    e = np.zeros((5, 16))
    states = [_transition(rng, 1) for _ in range(4)]
    acts, rews = [2, 0, 2, 4], [-1.0, -5.0, 3.0]
    for t in range(3):
        s, s2 = states[t], states[t + 1]
        a, a2 = np.array([acts[t]]), np.array([acts[t + 1]])
        phi, phi2 = fb.features(s)[0].astype(np.float64), fb.features(s2)[0].astype(np.float64)
        done = t == 2
        d_hand = rews[t] + (0.0 if done else g * (W[acts[t + 1]] @ phi2)) - W[acts[t]] @ phi
        e *= float(f32(f32(g) * f32(lam)))
        e[acts[t]] += phi
        d = o.update(s, a, np.array([rews[t]], dtype=np.float32), s2, a2, np.array([done]), np.array([0]))
        o.tick()
        assert abs(d[0] - d_hand) < 1e-5
        assert np.allclose(o.trace[0], 0 if done else e, atol=1e-6)
        W_new = W + alpha * fb.alpha_scale[None, :] * d_hand * e
        o.apply()                                                              # sync every step: classical rule
        assert np.allclose(o.W[0], W_new, atol=1e-6)
        W = o.W[0].astype(np.float64).copy()
    assert np.all(o.trace == 0)


def test_lambda_zero_is_one_step_sarsa():
    o = OptionSet(1, 2, 8, gamma=0.9, lam=0.0, alpha=0.05)
    rng = np.random.default_rng(1)
    o.W[:] = rng.standard_normal(o.W.shape).astype(np.float32) * 0.1
    for _ in range(3):
        S, S2 = _transition(rng, 8), _transition(rng, 8)
        A, A2 = rng.integers(0, 5, 8), rng.integers(0, 5, 8)
        r = rng.standard_normal(8).astype(np.float32)
        d = o.update(S, A, r, S2, A2, np.zeros(8, bool), np.zeros(8, int))
        phi = o.basis.features(S)
        want = np.zeros_like(o.trace)
        want[np.arange(8), A] = phi                                            # no memory of earlier steps
        assert np.array_equal(o.trace, want)
        dW = np.zeros((5, o.F))
        for b in range(8):
            dW[A[b]] += float(d[b]) * phi[b].astype(np.float64)
        assert np.allclose(o.dW[0], dW, atol=1e-9)
        o.tick(); o.apply()


def test_done_uses_reward_only_target_and_zeroes_trace():
    o = OptionSet(2, 1, 4, gamma=0.9, lam=0.9)
    rng = np.random.default_rng(2)
    o.W[:] = rng.standard_normal(o.W.shape).astype(np.float32)
    S, S2 = _transition(rng, 4), _transition(rng, 4)
    A, A2 = np.array([0, 1, 2, 3]), np.array([4, 4, 4, 4])
    opt = np.array([0, 1, 0, 1])
    done = np.array([True, True, False, False])
    r = np.ones(4, dtype=np.float32)
    d = o.td_error(S, A, r, S2, A2, done, opt)
    q = o.q(S, opt)[np.arange(4), A]
    assert np.allclose(d[:2], 1.0 - q[:2], atol=1e-6)
    o.update(S, A, r, S2, A2, done, opt)
    assert np.all(o.trace[:2] == 0) and np.all(np.abs(o.trace[2:]).sum(axis=(1, 2)) > 0)
    assert o.cnt.tolist() == [2, 2]


def test_apply_normalises_by_count_and_respects_alpha_scale():
    o = OptionSet(2, 1, 6, gamma=0.9, lam=0.5, alpha=0.2)
    o.dW[0, 3, :] = 6.0
    o.cnt[:] = [3, 0]
    o.window_steps = 2
    o.apply()
    assert np.allclose(o.W[0, 3], 0.2 * o.basis.alpha_scale * 6.0 * (2.0 / 3.0), rtol=1e-6)
    assert np.all(o.W[1] == 0) and np.all(o.dW == 0) and np.all(o.cnt == 0) and o.window_steps == 0


def test_windowed_form_equals_dense_form():
    """dense per-step sweep == forward-view window form (frozen weights),
    including masks, terminations, option changes after a termination, and several windows."""
    omap = oracle.PinballMap.from_name("easy")
    rng = np.random.default_rng(0)
    B, K = 48, 3
    hp = dict(gamma=0.95, lam=0.8, alpha=0.05, seed=1)
    d, w = OptionSet(K, 2, B, **hp), OptionSet(K, 2, B, windowed=True, **hp)
    W = (rng.standard_normal(d.W.shape) * 0.05).astype(np.float32)
    d.W[:] = W; w.W[:] = W
    opt = rng.integers(0, K, B).astype(np.int32)
    for it in range(24):
        S, S2 = omap.sample_free_states(rng, B), omap.sample_free_states(rng, B)
        A, A2 = rng.integers(0, 5, B), rng.integers(0, 5, B)
        r = rng.standard_normal(B).astype(np.float32)
        done = rng.random(B) < 0.15
        mask = None if it % 3 else rng.random(B) < 0.7
        dd, dw = d.update(S, A, r, S2, A2, done, opt, mask), w.update(S, A, r, S2, A2, done, opt, mask)
        assert np.allclose(dd, dw, atol=1e-6)
        d.tick(); w.tick()
        ch = done & (np.ones(B, bool) if mask is None else mask)
        opt = np.where(ch, rng.integers(0, K, B), opt).astype(np.int32)
        if it % 5 == 4:                                                        # window of 5 steps
            w.flush()
            assert np.abs(w.trace - d.trace).max() < 1e-6
            assert np.abs(w.dW - d.dW).max() < 1e-6 * max(1.0, np.abs(d.dW).max())
            assert np.array_equal(w.cnt, d.cnt)
        if it % 10 == 9:
            d.apply(); w.apply()
            assert np.abs(d.W - w.W).max() < 1e-6
    w.flush(); w.flush()                                                       # idempotent when empty


def test_batched_equals_per_env_and_sharding_invariance():
    """B envs in one OptionSet == the sum of B single-env OptionSets == 2 shards with summed deltas."""
    rng = np.random.default_rng(5)
    B, K = 12, 2
    hp = dict(gamma=0.9, lam=0.7, alpha=0.1)
    W = (rng.standard_normal((K, 5, 81)) * 0.1).astype(np.float32)
    full = OptionSet(K, 2, B, **hp); full.W[:] = W
    halves = [OptionSet(K, 2, B // 2, env_offset=i * B // 2, **hp) for i in range(2)]
    singles = [OptionSet(K, 2, 1, env_offset=b, **hp) for b in range(B)]
    for o in halves + singles:
        o.W[:] = W
    for it in range(4):
        S, S2 = _transition(rng, B), _transition(rng, B)
        A, A2 = rng.integers(0, 5, B), rng.integers(0, 5, B)
        r = rng.standard_normal(B).astype(np.float32)
        done = rng.random(B) < 0.2
        opt = rng.integers(0, K, B)
        df = full.update(S, A, r, S2, A2, done, opt); full.tick()
        for i, h in enumerate(halves):
            sl = slice(i * B // 2, (i + 1) * B // 2)
            assert np.array_equal(h.update(S[sl], A[sl], r[sl], S2[sl], A2[sl], done[sl], opt[sl]), df[sl]); h.tick()
        for b, o in enumerate(singles):
            sl = slice(b, b + 1)
            assert np.array_equal(o.update(S[sl], A[sl], r[sl], S2[sl], A2[sl], done[sl], opt[sl]), df[sl]); o.tick()
        assert np.allclose(sum(h.dW for h in halves), full.dW, atol=1e-12)
        assert np.allclose(sum(o.dW for o in singles), full.dW, atol=1e-12)
        assert np.array_equal(sum(h.cnt for h in halves), full.cnt)
    dW, cnt = sum(h.dW for h in halves), sum(h.cnt for h in halves)           # the "allreduce"
    full.apply()
    for h in halves:
        h.apply(dW.copy(), cnt.copy())
        assert np.array_equal(h.W, full.W)                                     # replicas stay bit-identical


# ---- initiation classifier -----------------------------------------------------------------------
def test_logistic_known_answers():
    o = OptionSet(2, 1, 1)
    X = np.array([[0.2, 0.4, 0, 0], [0.9, 0.1, 0, 0]], dtype=np.float32)
    assert np.all(o.initiation_prob(X) == 0.5) and o.initiation(X).all()        # theta = 0 -> p = 0.5 -> inside
    psi = logistic_features(X)
    assert np.allclose(psi[0], [1, 0.2, 0.4, 0.04, 0.08, 0.16])
    y = np.array([1, 0])
    g = o.clf_grad(0, X, y)                                                    # mean (0.5 - y) psi
    assert np.allclose(g, ((0.5 - y)[:, None] * psi).mean(axis=0), atol=1e-7)
    o.fit_initiation(0, X, y, steps=1, lr=2.0)
    assert np.allclose(o.theta[0], -2.0 * g, atol=1e-7)
    assert np.isclose(sigmoid(0.0), 0.5) and sigmoid(40.0) > 0.999999 and sigmoid(-40.0) < 1e-6


def test_logistic_separable_set_converges():
    rng = np.random.default_rng(0)
    X = rng.random((600, 2)).astype(np.float32)
    y = (X[:, 0] > 0.5).astype(np.uint8)
    o = OptionSet(1, 1, 1)
    o.fit_initiation(0, X, y, steps=3000, lr=5.0)
    pred = o.initiation(np.concatenate([X, np.zeros_like(X)], axis=1))[:, 0]
    assert (pred == y.astype(bool)).mean() > 0.97


def test_option_view():
    s = OptionSet(3, 1, 4, epsilon=0.0)
    s.W[1, 2] = 1.0
    opt = Option(s, 1)
    S = np.random.default_rng(0).uniform(0.1, 0.9, (4, 4)).astype(np.float32)
    assert opt.q(S).shape == (4, 5) and np.all(opt.q(S)[:, 0] == 0)
    assert opt.act(S).shape == (4,) and opt.initiation(S).shape == (4,)
