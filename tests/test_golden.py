"""The oracle against the committed golden fixtures (tests/golden/, made by tools/make_golden.py).
CPU only: pins the oracle's definition against regressions; the GPU side of the same fixtures is in
tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

import oracle
from oracle.pinball import step_batched, step_scalar

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("name", ["easy", "hard"])
def test_step_fixture(name):
    g = gold(f"step_{name}.npz")
    m = oracle.PinballMap.from_name(name)
    assert np.array_equal(m.edges.view(np.uint32), g["edges"].view(np.uint32))
    ns, r, fl = step_batched(m, g["state"], g["action"])
    assert np.array_equal(ns.view(np.uint32), g["next_state"].view(np.uint32))
    assert np.array_equal(r, g["reward"]) and np.array_equal(fl, g["flags"])
    for b in range(0, len(fl), 37):                          # spot-check the scalar (normative) form too
        s1, r1, f1 = step_scalar(m, g["state"][b], g["action"][b])
        assert np.array_equal(s1, g["next_state"][b]) and f1 == g["flags"][b]
    done, kind, obst, edge = oracle.unpack_flags(g["flags"])
    assert done.sum() >= 2 and (kind == 1).sum() > 50 and (kind == 2).sum() >= 1


@pytest.mark.parametrize("order", [3, 5])
def test_features_q_fixture(order):
    g = gold(f"features_q_o{order}.npz")
    assert np.array_equal(oracle.FourierBasis(order).features(g["state"]), g["phi"])
    o = oracle.OptionSet(g["W"].shape[0], order, len(g["state"]))
    o.W[:] = g["W"]
    assert np.array_equal(o.q(g["state"], g["option"]), g["Q"])


@pytest.mark.parametrize("windowed", [False, True])
def test_sarsa_fixture(windowed):
    g = gold("sarsa_o3.npz")
    gamma, lam, alpha = g["hp"]
    B = g["S"].shape[1]
    o = oracle.OptionSet(g["W0"].shape[0], 3, B, gamma=gamma, lam=lam, alpha=alpha, seed=1, windowed=windowed)
    o.W[:] = g["W0"]
    tol = 0 if not windowed else 2e-6
    for it in range(6):
        d = o.update(g["S"][it], g["A"][it], g["r"][it], g["S2"][it], g["A2"][it], g["done"][it], g["option"][it])
        o.tick()
        assert np.abs(d - g["delta"][it]).max() <= tol * max(1.0, np.abs(g["delta"][it]).max())
        if it == 2:
            o.flush()
            assert np.allclose(o.dW, g["dW3"], rtol=1e-5, atol=1e-6) and np.array_equal(o.cnt, g["cnt3"])
            assert np.allclose(o.trace.astype(np.float64).sum(axis=2), g["trace3_sum"], atol=1e-4)
            o.apply()
            assert np.allclose(o.W, g["W3"], atol=1e-7)
    o.flush()
    assert np.allclose(o.dW, g["dW_end"], rtol=1e-5, atol=1e-5)
    assert np.allclose(o.trace.astype(np.float64).sum(axis=2), g["trace_end_sum"], atol=1e-4)


def test_classifier_fixture():
    g = gold("classifier.npz")
    c = oracle.OptionSet(2, 1, 1)
    c.theta[0] = g["theta0"]
    assert np.array_equal(c.clf_grad(0, g["X"], g["y"]), g["grad0"])
    assert np.array_equal(c.fit_initiation(1, g["X"], g["y"], steps=100, lr=2.0), g["theta1_fit"])
    S4 = np.concatenate([g["X"], np.zeros_like(g["X"])], axis=1)
    assert np.array_equal(c.initiation_prob(S4), g["prob"])


def test_agent_step_fixture():
    g = gold("agent_step.npz")
    B, K = len(g["state"]), g["theta"].shape[0]
    m = oracle.PinballMap.from_name("easy")
    cfg = dict(map="easy", batch=B, order=3, max_options=K, seed=2, sync_interval=3, option_timeout=3, epsilon=0.2)
    ag = oracle.SkillChainAgent(oracle.AgentConfig(**cfg), m)
    ag.env.reset(states=g["state"])
    ag.start_xy = g["state"][:, :2].copy()
    ag.options.W[:] = g["W"]
    ag.options.theta[:] = g["theta"]
    ag.active[:2] = True
    ag.n_active = 2
    ag.parents[1], ag.parents[2] = 1, 2
    ag.option, ag.t_opt, ag.action = g["option"].copy(), g["t_opt"].copy(), g["action"].copy()
    out = ag.step()
    assert np.array_equal(out["state"], g["next_state"]) and np.array_equal(out["reward"], g["reward"])
    assert np.array_equal(out["term"], g["term"]) and np.array_equal(out["hit"], g["hit"])
    assert np.array_equal(out["delta"], g["delta"]) and np.array_equal(out["option"], g["next_option"])
    assert np.array_equal(out["action"], g["next_action"]) and np.array_equal(ag.options.cnt, g["cnt"])
    assert g["term"].sum() > 20 and g["hit"].sum() > 3
