"""CPU tests of the test infrastructure itself: the oracle's follow mode (oracle/agent.py step(follow=)) and the sharded
replay helper (tests/oracle_replay.py) that the full-size GPU parity tests rest on."""
import numpy as np

import oracle
from oracle.compare import assert_close, mismatch, robust_scale
from oracle_replay import activate, default_theta, replay_sharded

HP = dict(map="easy", order=2, max_options=4, seed=21, sync_interval=1000, epsilon=0.05, alpha=1e-3, option_timeout=3,
          max_episode_steps=2000, graph=False)


def _free_run(B, T):
    omap = oracle.PinballMap.from_name("easy")
    S = omap.sample_free_states(np.random.default_rng(77), B)
    W = (np.random.default_rng(5).standard_normal((4, 5, 81)) * 0.1).astype(np.float32)
    theta = default_theta(4)
    ag = oracle.SkillChainAgent(oracle.AgentConfig(batch=B, **HP), omap)
    ag.env.reset(states=S)
    ag.start_xy = ag.env.state[:, :2].copy()
    ag.options.W[:] = W
    activate(ag, theta, 2)
    ag.option = (np.arange(B) % 3).astype(np.int32)
    ag.action = ag.options.act(ag.env.state, ag.option, step=0xFFFFFFFF, stream=2)
    acts, opts, dl = [ag.action.copy()], [ag.option.copy()], []
    for _ in range(T):
        out = ag.step()
        acts.append(out["action"]); opts.append(out["option"]); dl.append(out["delta"])
    return ag, S, W, theta, np.stack(acts), np.stack(opts), np.stack(dl)


def test_following_its_own_choices_reproduces_the_free_run_and_shards_sum():
    B, T = 64, 5
    ag, S, W, theta, acts, opts, dl = _free_run(B, T)
    jobs = [dict(cfg=dict(batch=32, env_offset=lo, **HP), S0=S[lo:lo + 32], W=W, theta=theta, n_active=2, graph=False,
                 actions=acts[:, lo:lo + 32], options=opts[:, lo:lo + 32], probe=[1, 5]) for lo in (0, 32)]
    for procs in (1, 2):                                   # inline and in spawned processes
        res = replay_sharded(jobs, procs)
        assert np.array_equal(np.concatenate([r["delta"] for r in res], axis=1), dl)
        assert sum(r["n_dis_action"] for r in res) == 0 and sum(r["n_dis_option"] for r in res) == 0
        assert np.array_equal(np.concatenate([r["state"] for r in res]), ag.env.state)
        assert np.array_equal(sum(r["cnt"] for r in res), ag.options.cnt)
        assert np.abs(sum(r["dW"] for r in res) - ag.options.dW).max() <= 1e-9 * np.abs(ag.options.dW).max()
        assert np.array_equal(np.concatenate([r["trace_probe"] for r in res]), ag.options.trace[[1, 5, 33, 37]])
        assert sum(r["n_success"] for r in res).sum() == ag.n_success.sum()


def test_follow_reports_disagreements():
    B, T = 32, 2
    ag, S, W, theta, acts, opts, dl = _free_run(B, T)
    wrong = acts.copy()
    wrong[1, 3] = (wrong[1, 3] + 1) % 5                    # a followed action that is not the oracle's choice
    res = replay_sharded([dict(cfg=dict(batch=B, env_offset=0, **HP), S0=S, W=W, theta=theta, n_active=2,
                               actions=wrong, options=opts)], 1)
    assert res[0]["n_dis_action"] >= 1


def test_assert_close_is_elementwise():
    b = np.array([5.0, -3.0, 10000.0, 2.0, 0.0, 4.0])
    assert robust_scale(b) < 10                            # the goal reward does not inflate the scale
    a = b.copy()
    a[0] *= 1.01                                           # a 1 % error on an ordinary element must fail ...
    assert mismatch(a, b)[0] == 1
    a = b.copy()
    a[2] += 0.5                                            # ... while 5e-5 relative on the large one passes
    assert_close(a, b)
    a = b.copy()
    a[4] = 1e-4                                            # an exact zero may be off by rtol * typical magnitude
    assert_close(a, b)
    a[4] = 1e-2
    assert mismatch(a, b)[0] == 1
    assert mismatch(np.array([np.nan]), np.array([1.0]))[0] == 1
