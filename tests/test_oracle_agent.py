"""Oracle skill-chaining agent: the lock-step step's bookkeeping, the controller, chain vs graph mode."""
import numpy as np

import oracle
from oracle.agent import GOAL_BIT


def _agent(B=64, K=3, **kw):
    cfg = oracle.AgentConfig(map="easy", batch=B, order=2, max_options=K, seed=3, **kw)
    return oracle.SkillChainAgent(cfg)


def test_initial_state_and_first_step_shapes():
    ag = _agent(B=16)
    assert np.all(ag.option == 0) and ag.parents[0] == GOAL_BIT and ag.n_active == 0
    out = ag.step()
    for k in ("state", "reward", "env_done", "term", "hit", "delta", "option", "action"):
        assert len(out[k]) == 16
    assert ag.t == 1 and np.all(ag.t_opt == 1) and np.all(ag.ep_steps == 1)
    assert np.all((out["action"] >= 0) & (out["action"] < 5))


def test_goal_hit_terminates_records_example_and_resets():
    ag = _agent(B=8, option_timeout=1000)
    tx, ty, _ = (float(v) for v in ag.map.target)
    S = np.tile(np.array([tx - 0.045, ty, 1.0, 0.0], dtype=np.float32), (8, 1))
    S[4:] = [0.5, 0.9, 0.0, 0.0]                                  # far from the goal
    ag.env.reset(states=S)
    ag.start_xy = S[:, :2].copy()
    ag.action = np.full(8, 4, dtype=np.int32)
    out = ag.step()
    assert out["env_done"][:4].all() and not out["env_done"][4:].any()
    assert out["hit"][:4].all() and out["term"][:4].all() and not out["term"][4:].any()
    assert np.all(out["reward"][:4] == 10000.0)                    # no option bonus on top of the task goal
    assert ag.n_success[0] == 4 and ag.n_fail[0] == 0 and ag.ex_count[0] == 4
    assert np.all(ag.ex_label[0, :4] == 1) and np.allclose(ag.ex_xy[0, :4], S[:4, :2])
    assert np.all(ag.env.state[:4, :2] == ag.map.starts[0]) and np.all(ag.env.state[:4, 2:] == 0)
    assert ag.episodes[:4].tolist() == [1] * 4 and ag.goals[:4].tolist() == [1] * 4
    assert np.all(ag.options.trace[:4] == 0) and np.all(ag.t_opt[:4] == 0) and np.all(ag.t_opt[4:] == 1)


def test_option_timeout_gives_negative_example():
    ag = _agent(B=4, option_timeout=2)
    ag.step()
    out = ag.step()
    assert out["term"].all() and not out["hit"].any()
    assert ag.n_fail[0] == 4 and np.all(ag.ex_label[0, :4] == 0)


def test_manage_promotes_and_wires_chain_and_graph():
    for graph in (False, True):
        ag = _agent(B=32, K=4, gestation_successes=5, graph=graph)
        rng = np.random.default_rng(0)
        X = rng.random((200, 2)).astype(np.float32)
        ag.ex_xy[0, :200] = X
        ag.ex_label[0, :200] = (X[:, 0] > 0.6).astype(np.uint8)
        ag.ex_count[0] = 200
        assert ag.manage() is False                                 # not enough successes yet
        ag.n_success[0] = 5
        assert ag.manage() is True
        assert ag.n_active == 1 and ag.active[0] and not ag.active[1]
        assert ag.parents[1] == (np.uint32(1) | GOAL_BIT if graph else np.uint32(1))
        bits = ag.initiation_bits(np.array([[0.9, 0.5, 0, 0], [0.1, 0.5, 0, 0]], dtype=np.float32))
        assert bits.tolist() == [1, 0]
        assert ag.choose_option(bits).tolist() == [0, 1]            # inside I_0 -> option 0, else the gestating slot
        ag.n_success[1] = 5; ag.ex_count[1] = 10
        ag.ex_xy[1, :10] = X[:10]; ag.ex_label[1, :10] = 1
        assert ag.manage() is True
        assert ag.parents[2] == (np.uint32(3) | GOAL_BIT if graph else np.uint32(2))
        ag.n_success[2] = 99
        assert ag.manage() is True                                  # no examples: theta stays 0 (accepts everywhere)
        assert ag.n_active == 3 and np.all(ag.options.theta[2] == 0)
        ag.n_success[3] = 99
        assert ag.manage() is False and not ag.active[3]            # slot K-1 is never promoted


def test_subgoal_hit_pays_option_bonus():
    ag = _agent(B=6, K=3, option_bonus=123.0)
    ag.options.theta[0] = [-5.0, 10.0, 0, 0, 0, 0]                  # I_0 = {x >= 0.5}
    ag.active[0] = True; ag.n_active = 1; ag.parents[1] = 1
    S = np.tile(np.array([0.49, 0.9, 1.0, 0.0], dtype=np.float32), (6, 1))   # crosses x = 0.5 this step
    S[3:, 0] = 0.2
    ag.env.reset(states=S)
    ag.option[:] = 1
    ag.action = np.full(6, 4, dtype=np.int32)
    out = ag.step()
    assert out["hit"][:3].all() and not out["hit"][3:].any()
    assert np.all(out["reward"][:3] == np.float32(-1.0 + 123.0)) and np.all(out["reward"][3:] == -1.0)
    assert np.all(out["option"][:3] == 0) and np.all(out["option"][3:] == 1)
    assert ag.n_success[1] == 3


def test_run_episode_learns_to_reach_the_goal_from_nearby():
    cfg = oracle.AgentConfig(map="easy", batch=64, order=3, max_options=2, seed=0, alpha=5e-3, epsilon=0.1,
                             sync_interval=1, max_episode_steps=80, option_timeout=80)
    ag = oracle.SkillChainAgent(cfg)
    st = ag.run_episode(max_steps=120, manage_every=40)
    assert st["steps"] <= 120 and st["env_steps"] == st["steps"] * 64 and st["finished"] >= 60


def test_sync_interval_only_changes_when_weights_move():
    a, b = _agent(B=16, sync_interval=1), _agent(B=16, sync_interval=4)
    for i in range(4):
        a.step(); b.step()
        if i < 3:
            assert np.all(b.options.W == 0) and b.options.cnt.sum() == 16 * (i + 1)
    assert np.abs(a.options.W).max() > 0 and np.abs(b.options.W).max() > 0 and b.options.cnt.sum() == 0


def test_windowed_agent_equals_dense_agent():
    """The whole agent with Sarsa(lambda) in the forward-view window form == with the dense per-step sweep
    (weights only move at syncs, so the TD errors, hence the trajectories, are the same; traces and weights agree
    to rounding)."""
    kw = dict(map="hard", batch=96, order=2, max_options=3, seed=5, sync_interval=6, option_timeout=4, epsilon=0.2,
              alpha=0.01)
    d = oracle.SkillChainAgent(oracle.AgentConfig(**kw))
    w = oracle.SkillChainAgent(oracle.AgentConfig(windowed=True, **kw))
    rng = np.random.default_rng(0)
    W = (rng.standard_normal(d.options.W.shape) * 0.2).astype(np.float32)
    d.options.W[:] = W; w.options.W[:] = W
    for t in range(20):
        od, ow = d.step(), w.step()
        assert np.array_equal(od["state"], ow["state"]) and np.array_equal(od["action"], ow["action"])
        assert np.allclose(od["delta"], ow["delta"], atol=1e-5 * max(1.0, np.abs(od["delta"]).max()))
        if (t + 1) % 6 == 0:
            assert np.abs(d.options.W - w.options.W).max() < 1e-6
    w.options.flush()
    assert np.abs(d.options.trace - w.options.trace).max() < 1e-5
    assert (d.n_success + d.n_fail).sum() > 50
