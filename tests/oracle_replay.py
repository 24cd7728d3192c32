"""Replay of a GPU agent's recorded steps by the CPU oracle in follow mode (oracle/agent.py SkillChainAgent.step
`follow=`), optionally sharded over host processes so that full BASELINE.json batch sizes finish in seconds.

The GPU agent runs free (its own carried Q values, its own arg-max and eps-greedy choices, any number of steps per
launch).  The oracle then steps from the same initial state and, after each of its steps, adopts the action and option
the GPU took, while still computing its own.  What is compared is therefore the arithmetic of every step (states bit for
bit, TD errors, traces, weight deltas, weights element-wise), and the choices themselves (they must agree except where
the oracle's own Q row has a near-tie at the top).

Sharding uses the property the oracle defines and tests on the CPU (tests/test_oracle_agent.py sharding invariance):
envs are independent while the weights are frozen, dW and cnt are sums over envs.
"""
import numpy as np


def activate(oag, theta, n_active, graph=False):
    """Put an oracle agent in the state 'options 0..n_active-1 active with classifiers theta' (chain or graph parents)."""
    oag.options.theta[:] = theta
    oag.active[:] = False
    oag.active[:n_active] = True
    oag.n_active = n_active
    oag.parents[:] = 0
    oag.parents[0] = np.uint32(1 << 31)
    for n in range(1, min(n_active + 1, oag.options.K)):
        oag.parents[n] = (np.uint32((1 << n) - 1) | np.uint32(1 << 31)) if graph else np.uint32(1 << (n - 1))


def default_theta(K):
    """Up to 8 distinct, overlapping initiation regions (half planes and one disc) so that terminations, hits and
    re-selections of several different options happen within a few steps."""
    base = np.array([[-6.0, 10.0, 0.0, 0.0, 0.0, 0.0],        # x >= 0.6
                     [4.5, 0.0, -10.0, 0.0, 0.0, 0.0],        # y <= 0.45
                     [-3.0, 0.0, 10.0, 0.0, 0.0, 0.0],        # y >= 0.3
                     [2.0, -10.0, 0.0, 0.0, 0.0, 0.0],        # x <= 0.2
                     [-4.0, 5.0, 5.0, 0.0, 0.0, 0.0],         # x + y >= 0.8
                     [-6.5, 20.0, 20.0, -20.0, 0.0, -20.0],   # disc of radius ~0.42 around (0.5, 0.5)
                     [1.0, 4.0, -6.0, 0.0, 0.0, 0.0],
                     [0.5, -3.0, 2.0, 0.0, 0.0, 0.0]], dtype=np.float32)
    theta = np.zeros((K, 6), dtype=np.float32)
    theta[:min(K, len(base))] = base[:K]
    return theta


def replay_shard(job):
    """One shard of the oracle replay (top-level so that it can run in a spawned process).
    job: dict(cfg=oracle AgentConfig kwargs for the shard, S0 (b, 4), A0, O0, t_opt0, ep0, start_xy0 or None, W, theta,
              n_active, graph, trace0 or None, actions (T+1, b) / options (T+1, b): row t = what the env holds BEFORE
              step t (row T = after the last step), apply_inside (bool))"""
    import oracle
    from oracle.agent import AgentConfig, SkillChainAgent
    cfg = AgentConfig(**job["cfg"])
    ag = SkillChainAgent(cfg, oracle.PinballMap.from_name(cfg.map))
    ag.env.reset(states=job["S0"])
    ag.start_xy = (job["start_xy0"] if job.get("start_xy0") is not None else ag.env.state[:, :2]).copy()
    ag.options.W[:] = job["W"]
    activate(ag, job["theta"], job["n_active"], job.get("graph", False))
    ag.action = job["actions"][0].astype(np.int32).copy()
    ag.option = job["options"][0].astype(np.int32).copy()
    if job.get("t_opt0") is not None:
        ag.t_opt = job["t_opt0"].astype(np.int32).copy()
    if job.get("ep0") is not None:
        ag.ep_steps = job["ep0"].astype(np.int32).copy()
    if job.get("trace0") is not None:
        ag.options.trace[:] = job["trace0"]
    ag.t = int(job.get("t0", 0))
    T = len(job["actions"]) - 1
    deltas, states, n_dis_a, n_dis_o, worst_gap = [], [], 0, 0, 0.0
    for t in range(T):
        states.append(ag.env.state.copy())
        out = ag.step(follow=dict(action=job["actions"][t + 1], option=job["options"][t + 1]))
        deltas.append(out["delta"])
        dis = out["own_action"] != job["actions"][t + 1]
        if dis.any():
            # a legitimate disagreement is an arg-max near-tie: the followed action's Q is within rounding of the best
            Q = out["Qsel"][dis].astype(np.float64)
            gap = Q.max(axis=1) - Q[np.arange(len(Q)), job["actions"][t + 1][dis]]
            worst_gap = max(worst_gap, float((gap / max(1e-30, float(np.abs(out["Qsel"]).mean()))).max()))
        n_dis_a += int(dis.sum())
        n_dis_o += int((out["own_option"] != job["options"][t + 1]).sum())
    o = ag.options
    if not job.get("apply_inside"):
        o.flush() if o.windowed else None
    return dict(delta=np.stack(deltas) if deltas else np.zeros((0, cfg.batch), np.float32),
                pre_state=np.stack(states) if states else np.zeros((0, cfg.batch, 4), np.float32),
                state=ag.env.state.copy(), dW=o.dW.copy(), cnt=o.cnt.copy(), W=o.W.copy(),
                trace=o.trace.copy() if job.get("want_trace") else None,
                trace_rowsum=o.trace.astype(np.float64).sum(axis=2),
                trace_probe=o.trace[np.asarray(job.get("probe", []), dtype=np.int64)].copy(),
                t_opt=ag.t_opt.copy(), ep_steps=ag.ep_steps.copy(), n_success=ag.n_success.copy(),
                n_fail=ag.n_fail.copy(), ex_count=ag.ex_count.copy(), ex_xy=ag.ex_xy.copy(), ex_label=ag.ex_label.copy(),
                start_xy=ag.start_xy.copy(), n_dis_action=n_dis_a, n_dis_option=n_dis_o, worst_gap=worst_gap)


def replay_sharded(jobs, procs):
    """Run the shard jobs (in order) on up to `procs` spawned processes; a single job runs inline."""
    if len(jobs) == 1 or procs <= 1:
        return [replay_shard(j) for j in jobs]
    import multiprocessing as mp
    import os
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(v, "1")
    with mp.get_context("spawn").Pool(min(procs, len(jobs))) as pool:
        return pool.map(replay_shard, jobs, chunksize=1)
