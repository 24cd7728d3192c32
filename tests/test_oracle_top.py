"""Known-answer tests of the top-level SMDP learner of the CPU oracle (oracle/agent.py, top_level=True): the paper's
agent learns which option to run.  Written from the definition in the module docstring."""
import numpy as np

import oracle
from oracle.philox import STREAM_TOP, draws, uniform01
from oracle_replay import activate, default_theta

f32 = np.float32


def _agent(B=64, K=4, order=2, n_active=2, **kw):
    cfg = dict(map="easy", batch=B, order=order, max_options=K, top_level=True, option_timeout=3, sync_interval=1000,
               epsilon=0.2, seed=5)
    cfg.update(kw)
    ag = oracle.SkillChainAgent(oracle.AgentConfig(**cfg))
    rng = np.random.default_rng(1)
    ag.env.reset(states=ag.map.sample_free_states(rng, B))
    ag.start_xy = ag.env.state[:, :2].copy()
    ag.opt_s0 = ag.env.state.copy()
    activate(ag, default_theta(K), n_active)
    ag.options.W[:K] = (rng.standard_normal(ag.options.W[:K].shape) * 0.1).astype(np.float32)
    return ag, rng


def test_layout_of_the_top_level_slots():
    ag, rng = _agent(K=7)
    o = ag.options
    assert o.top_slots == 2 and o.K_all == 9 and o.W.shape == (9, 5, 81) and o.cnt.shape == (9,)
    o.W[7:] = rng.standard_normal(o.W[7:].shape).astype(np.float32)
    S = ag.env.state
    Qt = o.q_top(S)
    assert Qt.shape == (64, 7)
    for j in range(7):                                       # Q_top(s, j) = row j % 5 of slot K + j // 5
        assert np.array_equal(Qt[:, j], o.q(S, np.full(64, 7 + j // 5))[:, j % 5])


def test_zero_top_weights_reproduce_the_first_active_rule():
    a, _ = _agent(alpha_top=0.0, epsilon_top=0.0)
    b, _ = _agent(alpha_top=0.0, epsilon_top=0.0)
    b.cfg.top_level = False                                   # same weight layout, choice by the fixed rule
    for _ in range(12):
        oa, ob = a.step(), b.step()
        assert np.array_equal(oa["option"], ob["option"]) and np.array_equal(oa["state"], ob["state"])
    assert len(np.unique(a.option)) >= 2


def test_smdp_target_by_hand_single_env():
    """B = 1, the option runs until its time-out (3 steps): delta_top = r0 + g r1 + g^2 r2 + g^3 max_adm Q_top(s3) - Q_top(s0, o)."""
    ag, rng = _agent(B=1, n_active=0, epsilon=1.0, option_timeout=3, gamma=0.9)      # only the gestating slot 0 is admissible
    o = ag.options
    o.W[o.K:] = (rng.standard_normal(o.W[o.K:].shape) * 0.3).astype(np.float32)
    ag.env.reset(states=np.array([[0.5, 0.5, 0.0, 0.0]], dtype=np.float32))
    ag.opt_s0 = ag.env.state.copy()
    s0 = ag.env.state.copy()
    rs = []
    for t in range(3):
        out = ag.step()
        rs.append(float(out["reward"][0]))
        assert bool(out["term"][0]) == (t == 2)
    s3 = out["state"]
    g = 0.9
    R = f32(f32(f32(rs[0]) + f32(f32(g) * f32(rs[1]))) + f32(f32(f32(g) * f32(g)) * f32(rs[2])))
    disc = f32(f32(f32(g) * f32(g)) * f32(g))
    Wt = o.W.copy()                                           # frozen: no apply happened (sync_interval 1000)
    target = f32(R + f32(disc * o.q_top(s3)[0, 0]))
    want = f32(target - o.q_top(s0)[0, 0])
    assert ag.last_delta_top[0] == want
    # and it went to row 0 of the top-level slot, with one event counted
    phi0 = o.basis.features(s0)[0].astype(np.float64)
    assert np.allclose(o.dW[o.K, 0], float(want) * phi0, rtol=1e-12) and o.cnt[o.K] == 1
    assert np.abs(o.dW[o.K, 1:]).max() == 0
    # apply: mean over events (1), alpha_top, per-feature scale
    before = o.W[o.K, 0].copy()
    o.apply()
    step = (o.alpha_top * o.basis.alpha_scale) * (float(want) * phi0).astype(np.float32)
    assert np.allclose(o.W[o.K, 0], before + step.astype(np.float32), rtol=1e-6, atol=1e-9)
    assert np.array_equal(o.W[: o.K], Wt[: o.K]) or o.cnt[: o.K].sum() == 0


def test_episode_end_does_not_bootstrap():
    ag, rng = _agent(B=1, n_active=0, epsilon=1.0, option_timeout=50)
    o = ag.options
    o.W[o.K:] = (rng.standard_normal(o.W[o.K:].shape) * 0.3).astype(np.float32)
    tx, ty, tr = ag.map.target
    ag.env.reset(states=np.array([[tx - 0.03, ty, 1.0, 0.0]], dtype=np.float32))      # flies into the goal
    ag.opt_s0 = ag.env.state.copy()
    s0 = ag.env.state.copy()
    out = ag.step()
    assert bool(out["env_done"][0]) and bool(out["term"][0])
    assert ag.last_delta_top[0] == f32(f32(10000.0) - o.q_top(s0)[0, 0])


def test_choice_is_eps_greedy_over_the_admissible_slots():
    ag, rng = _agent(B=512, K=4, n_active=3, epsilon_top=0.3)
    o = ag.options
    o.W[o.K:] = rng.standard_normal(o.W[o.K:].shape).astype(np.float32)
    S = ag.env.state
    bits = ag.initiation_bits(S)
    choice, Qt, adm = ag.choose_option_top(bits, S, t=9)
    assert adm[:, 3].all()                                    # the gestating slot is always admissible
    assert adm[np.arange(512), choice].all()                  # never an inadmissible option
    r = draws(ag.cfg.seed, o.env_ids, 9, STREAM_TOP)
    explore = uniform01(r[:, 0]) < f32(0.3)
    assert 0.15 < explore.mean() < 0.45
    greedy = np.argmax(np.where(adm, Qt, -np.inf), axis=1)
    assert np.array_equal(choice[~explore], greedy[~explore])
    n_adm = adm.sum(axis=1)
    pick = np.minimum((uniform01(r[:, 1]) * n_adm.astype(np.float32)).astype(np.int32), n_adm - 1)
    for b in np.nonzero(explore)[0][:50]:
        assert choice[b] == np.nonzero(adm[b])[0][pick[b]]


def test_learns_to_prefer_the_option_with_the_higher_return():
    """A learning sanity check: after some updates Q_top differs between options and the weights moved."""
    ag, _ = _agent(B=256, n_active=2, alpha_top=0.05, epsilon_top=0.2, sync_interval=4, option_timeout=4)
    for _ in range(40):
        ag.step()
    assert np.abs(ag.options.W[ag.options.K:]).max() > 0
    assert np.isfinite(ag.options.W).all()
