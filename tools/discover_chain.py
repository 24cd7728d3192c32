#!/usr/bin/env python
"""Run full skill chaining on a map until the option chain reaches the start state (or a wall-clock budget ends), then
measure what the discovered chain is worth: BASELINE.json's "a full skill chain discovered on the hard map in under one
wall-clock minute".

    python tools/discover_chain.py --map hard --batch 65536 --options 8 [--graph] [--no-top-level]

What makes the discovery meaningful (round 1's was not: with one start position and examples labelled only by "hit", the
first classifier already covered the start):
  * many start states: episodes begin at the map's start and at --starts random free-space positions, so option executions
    begin all over the map and the classifiers see positives and negatives everywhere;
  * an example is positive only if the option reached its target within --horizon steps of where it started
    (AgentConfig.init_horizon), so an initiation set is a neighbourhood of its target and the chain has to grow link by
    link from the goal back to the start;
  * options are chosen by the learned top-level SMDP value function (AgentConfig.top_level), not by slot order.
The controller runs on the device (one kernel per manage() call); the host only polls the mirror to time-stamp promotions.
After discovery a fresh batch of envs, all at the map's own start, runs the chain greedily (epsilon = 0) for one episode
each: the success rate and the steps to the goal are reported, next to the same numbers for the untrained agent.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def evaluate(scg, torch, base_map, src, args, n_env=4096, trained=True):
    """Greedy episodes from the map's own start with (a copy of) the agent's weights, classifiers and option graph."""
    cfg = scg.AgentConfig(map=args.map, batch=n_env, order=args.order, max_options=args.options, gamma=0.99, lam=0.9,
                          alpha=0.0, epsilon=0.0, sync_interval=8, option_timeout=args.option_timeout,
                          max_episode_steps=args.max_episode_steps, gestation_successes=1 << 30, graph=args.graph,
                          top_level=not args.no_top_level, alpha_top=0.0, epsilon_top=0.0, seed=123)
    ev = scg.SkillChainAgent(cfg, base_map)
    if trained:
        ev.options.set_weights(src.options.W)
        ev.options.theta.copy_(src.options.theta)
        c = src.controller_state(sync=True)
        ev.parents_host[:] = np.array(c["parents"], dtype=np.uint32)
        ev._ctl.n_active, ev._ctl.active_mask = c["n_active"], c["active_mask"]
        ev._push_ctl()
    st = ev.run_episode(max_steps=args.max_episode_steps, manage_every=64)
    goals = int(ev.stats[1])
    return dict(success_rate=goals / n_env, finished=st["finished"], mean_return=st["mean_return"], steps=st["steps"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default="hard")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--options", type=int, default=8)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--no-top-level", action="store_true")
    ap.add_argument("--budget", type=float, default=55.0)
    ap.add_argument("--alpha", type=float, default=2e-3)
    ap.add_argument("--alpha-top", type=float, default=1e-3)
    ap.add_argument("--epsilon", type=float, default=0.1)
    ap.add_argument("--gestation", type=int, default=20000)
    ap.add_argument("--train-seconds", type=float, default=25.0, help="keep learning this long after the chain is complete")
    ap.add_argument("--horizon", type=int, default=120)
    ap.add_argument("--option-timeout", type=int, default=120)
    ap.add_argument("--max-episode-steps", type=int, default=1500)
    ap.add_argument("--starts", type=int, default=255)
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--min-chain", type=int, default=3)
    args = ap.parse_args()
    import torch
    import skill_chaining_with_graphs_b200 as scg
    base = scg.PinballMap.from_name(args.map)
    rng = np.random.default_rng(0)
    extra = base.sample_free_states(rng, args.starts)[:, :2] if args.starts > 0 else np.zeros((0, 2), np.float32)
    starts = [tuple(float(v) for v in p) for p in base.starts] + [(float(x), float(y)) for x, y in extra]
    train_map = scg.PinballMap(float(base.ball_r), tuple(float(v) for v in base.target), starts,
                               [p.tolist() for p in base.polygons])
    cfg = scg.AgentConfig(map=args.map, batch=args.batch, order=args.order, max_options=args.options, gamma=0.99, lam=0.9,
                          alpha=args.alpha, epsilon=args.epsilon, sync_interval=8, option_timeout=args.option_timeout,
                          max_episode_steps=args.max_episode_steps, gestation_successes=args.gestation,
                          example_capacity=8192, clf_steps=300, clf_lr=2.0, graph=args.graph, option_bonus=1000.0,
                          top_level=not args.no_top_level, alpha_top=args.alpha_top, epsilon_top=0.1,
                          init_horizon=args.horizon)
    ag = scg.SkillChainAgent(cfg, train_map)
    before = evaluate(scg, torch, base, ag, args, trained=False)
    start = torch.tensor([[float(base.starts[0][0]), float(base.starts[0][1]), 0.0, 0.0]], device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    log, steps, seen, done = [], 0, 0, False

    def start_covered(n_active):
        return bool(ag.options.initiation(start)[0, :n_active].any()) if n_active else False

    # skill chaining stops creating options once an initiation set covers the start state (or the slots run out)
    while time.perf_counter() - t0 < args.budget and not done:
        ag.run(args.chunk)
        ag.manage()                                   # one kernel on the stream; nothing waits for it
        steps += args.chunk
        c = ag.controller_state(sync=False)           # the host mirror: promotions show up here a moment later
        if c["n_promotions"] > seen:
            seen = c["n_promotions"]
            covered = start_covered(c["n_active"])
            log.append(dict(t=round(time.perf_counter() - t0, 3), steps=steps, n_active=c["n_active"],
                            promoted_at_step=c["last_promotion_step"], start_covered=covered))
            print(json.dumps(log[-1]), flush=True)
            done = (covered and c["n_active"] >= args.min_chain) or c["n_active"] >= args.options - 1
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    # the chain is in place: keep learning the option policies and the top-level values for a while
    t1 = time.perf_counter()
    train_steps = 0
    while time.perf_counter() - t1 < args.train_seconds:
        ag.run(16 * args.chunk)
        ag.manage()
        train_steps += 16 * args.chunk
        torch.cuda.synchronize()
    train_wall = time.perf_counter() - t1
    c = ag.controller_state(sync=True)
    cnt = ag.counters()
    # how much of the free space each initiation set covers, and whether it contains the map's start
    probe = torch.as_tensor(base.sample_free_states(np.random.default_rng(1), 20000)).cuda()
    inside = ag.options.initiation(probe)[:, :c["n_active"]].float().mean(dim=0).cpu().numpy().tolist() if c["n_active"] else []
    after = evaluate(scg, torch, base, ag, args, trained=True)
    print(json.dumps(dict(map=args.map, batch=args.batch, graph=args.graph, top_level=not args.no_top_level,
                          wall_s=round(wall, 3), train_after_s=round(train_wall, 2), train_after_steps=train_steps, steps=steps, env_steps=steps * args.batch, n_active=c["n_active"],
                          chain_complete=done, start_covered=start_covered(c["n_active"]), parents=c["parents"],
                          initiation_set_fraction_of_free_space=[round(v, 3) for v in inside],
                          goals=cnt["goals"], episodes=cnt["episodes"], mean_return=cnt["mean_return"],
                          n_success=cnt["n_success"].tolist(), promotions=log,
                          greedy_from_start_untrained=before, greedy_from_start_after_discovery=after)))


if __name__ == "__main__":
    main()
