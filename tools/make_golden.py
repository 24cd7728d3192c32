#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the CPU oracle.

The reference repository ships no code, tests or vectors (/root/reference/README.md:1-2), so these
fixtures are outputs of the committed oracle on seeded inputs; they pin the oracle against
regressions (tests/test_golden.py, CPU) and are a second, file-based parity target for the CUDA path
(tests/test_gpu_parity.py, GPU).  Re-run only when the oracle's definition changes on purpose:

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle.pinball import step_scalar  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def near_edge_states(m, rng, n):
    S = m.sample_free_states(rng, n)
    e = m.edges[rng.integers(0, m.n_edges, n // 2)]
    t = rng.uniform(0, 1, n // 2).astype(np.float32)
    off = rng.uniform(-1.5, 1.5, n // 2).astype(np.float32) * m.ball_r
    S[: n // 2, 0] = e[:, 0] + t * e[:, 2] + off * e[:, 5]
    S[: n // 2, 1] = e[:, 1] + t * e[:, 3] + off * e[:, 6]
    return S


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- K1: single-step transitions (scalar, normative form) ----
    for name in ("easy", "hard"):
        m = oracle.PinballMap.from_name(name)
        rng = np.random.default_rng(2024)
        S = near_edge_states(m, rng, 768)
        tx, ty, _ = (float(v) for v in m.target)
        S[-8:] = [[tx - 0.05, ty, 1, 0], [tx, ty, 0, 0], [1.5, 0.5, 0, 0], [-0.2, -0.3, 0, 0], [0.5, 0.5, 0, 0],
                  [0.2, 0.9, 1, 1], [tx, ty + 0.045, 0, -1], [0.015, 0.015, -1, -1]]
        A = rng.integers(0, 5, len(S)).astype(np.int32)
        out = [step_scalar(m, S[b], A[b]) for b in range(len(S))]
        np.savez_compressed(os.path.join(OUT, f"step_{name}.npz"), state=S, action=A,
                            next_state=np.stack([o[0] for o in out]), reward=np.array([o[1] for o in out], dtype=np.float32),
                            flags=np.array([o[2] for o in out], dtype=np.int32), edges=m.edges)
    # ---- K2: features and Q ----
    m = oracle.PinballMap.from_name("easy")
    for order in (3, 5):
        rng = np.random.default_rng(order)
        S = m.sample_free_states(rng, 48 if order == 3 else 12)
        S[:, 2:] *= 1.4
        K = 3
        W = (rng.standard_normal((K, 5, (order + 1) ** 4)) * 0.05).astype(np.float32)
        opt = rng.integers(0, K, len(S)).astype(np.int32)
        os_ = oracle.OptionSet(K, order, len(S))
        os_.W[:] = W
        np.savez_compressed(os.path.join(OUT, f"features_q_o{order}.npz"), state=S, option=opt, W=W,
                            phi=oracle.FourierBasis(order).features(S), Q=os_.q(S, opt))
    # ---- K3: N updates on a fixed transition batch ----
    rng = np.random.default_rng(77)
    B, K, order = 96, 3, 3
    hp = dict(gamma=0.95, lam=0.8, alpha=0.05, seed=1)
    o = oracle.OptionSet(K, order, B, **hp)
    W0 = (rng.standard_normal(o.W.shape) * 0.05).astype(np.float32)
    o.W[:] = W0
    steps = []
    deltas = []
    opt = rng.integers(0, K, B).astype(np.int32)
    for it in range(6):
        S, S2 = m.sample_free_states(rng, B), m.sample_free_states(rng, B)
        A, A2 = rng.integers(0, 5, B).astype(np.int32), rng.integers(0, 5, B).astype(np.int32)
        r = rng.standard_normal(B).astype(np.float32)
        done = rng.random(B) < 0.15
        deltas.append(o.update(S, A, r, S2, A2, done, opt))
        o.tick()
        steps.append((S, A, r, S2, A2, done, opt.copy()))
        opt = np.where(done, rng.integers(0, K, B), opt).astype(np.int32)
        if it == 2:
            dW3, cnt3, trace3 = o.dW.astype(np.float32), o.cnt.copy(), o.trace.copy()
            o.apply()
            W3 = o.W.copy()
    np.savez_compressed(os.path.join(OUT, "sarsa_o3.npz"), W0=W0, hp=np.array([hp["gamma"], hp["lam"], hp["alpha"]]),
                        S=np.stack([s[0] for s in steps]), A=np.stack([s[1] for s in steps]),
                        r=np.stack([s[2] for s in steps]), S2=np.stack([s[3] for s in steps]),
                        A2=np.stack([s[4] for s in steps]), done=np.stack([s[5] for s in steps]),
                        option=np.stack([s[6] for s in steps]), delta=np.stack(deltas), dW3=dW3, cnt3=cnt3,
                        trace3_sum=trace3.astype(np.float64).sum(axis=2), W3=W3, dW_end=o.dW.astype(np.float32),
                        trace_end_sum=o.trace.astype(np.float64).sum(axis=2))
    # ---- K4: classifier ----
    rng = np.random.default_rng(8)
    X = rng.random((400, 2)).astype(np.float32)
    y = ((X[:, 0] - 0.6) ** 2 + (X[:, 1] - 0.4) ** 2 < 0.08).astype(np.uint8)
    c = oracle.OptionSet(2, 1, 1)
    c.theta[0] = rng.standard_normal(6).astype(np.float32)
    grad = c.clf_grad(0, X, y)
    fit = c.fit_initiation(1, X, y, steps=100, lr=2.0)
    S4 = np.concatenate([X, np.zeros_like(X)], axis=1)
    np.savez_compressed(os.path.join(OUT, "classifier.npz"), X=X, y=y, theta0=c.theta[0], grad0=grad, theta1_fit=fit,
                        prob=c.initiation_prob(S4))
    # ---- fused agent step ----
    B, K = 256, 4
    cfg = dict(map="easy", batch=B, order=3, max_options=K, seed=2, sync_interval=3, option_timeout=3, epsilon=0.2)
    rng = np.random.default_rng(5)
    S = near_edge_states(m, rng, B)
    A = rng.integers(0, 5, B).astype(np.int32)
    ag = oracle.SkillChainAgent(oracle.AgentConfig(**cfg), m)
    ag.env.reset(states=S)
    ag.start_xy = S[:, :2].copy()
    W = (rng.standard_normal(ag.options.W.shape) * 0.5).astype(np.float32)
    ag.options.W[:] = W
    theta = np.zeros((K, 6), dtype=np.float32)
    theta[0] = [-1.0, 2.0, 0, 0, 0, 0]
    theta[1] = [-0.6, 0.0, 2.0, 0, 0, 0]
    ag.options.theta[:] = theta
    ag.active[:2] = True
    ag.n_active = 2
    ag.parents[1], ag.parents[2] = 1, 2
    opt = rng.integers(0, 3, B).astype(np.int32)
    tq = rng.integers(0, 3, B).astype(np.int32)
    ag.option, ag.t_opt, ag.action = opt.copy(), tq.copy(), A.copy()
    out = ag.step()
    np.savez_compressed(os.path.join(OUT, "agent_step.npz"), state=S, action=A, W=W, theta=theta, option=opt, t_opt=tq,
                        next_state=out["state"], reward=out["reward"], term=out["term"], hit=out["hit"],
                        delta=out["delta"], next_option=out["option"], next_action=out["action"],
                        t_opt_after=ag.t_opt, n_success=ag.n_success, n_fail=ag.n_fail, cnt=ag.options.cnt,
                        dW=ag.options.dW.astype(np.float32), trace_sum=ag.options.trace.astype(np.float64).sum(axis=2))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
