#!/usr/bin/env python
"""Run full skill chaining on a map until the option chain is complete (or a wall-clock budget ends) and report the
wall time: BASELINE.json's "a full skill chain discovered on the hard map in under one wall-clock minute".

    python tools/discover_chain.py --map hard --batch 65536 --options 8 [--graph]
"""
import argparse
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map", default="hard")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--options", type=int, default=8)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--budget", type=float, default=60.0)
    ap.add_argument("--alpha", type=float, default=2e-3)
    ap.add_argument("--epsilon", type=float, default=0.1)
    ap.add_argument("--gestation", type=int, default=2000)
    ap.add_argument("--chunk", type=int, default=64)
    args = ap.parse_args()
    import torch
    import skill_chaining_with_graphs_b200 as scg
    cfg = scg.AgentConfig(map=args.map, batch=args.batch, order=args.order, max_options=args.options, gamma=0.99, lam=0.9,
                          alpha=args.alpha, epsilon=args.epsilon, sync_interval=8, option_timeout=200,
                          max_episode_steps=1000, gestation_successes=args.gestation, example_capacity=8192,
                          clf_steps=300, clf_lr=2.0, graph=args.graph, option_bonus=1000.0)
    ag = scg.SkillChainAgent(cfg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    log = []
    steps = 0
    start = torch.tensor([[float(ag.map.starts[0][0]), float(ag.map.starts[0][1]), 0.0, 0.0]], device="cuda")

    def start_covered():
        return bool((ag.options.initiation(start)[0, :ag.n_active]).any()) if ag.n_active else False

    # skill chaining stops creating options once an initiation set covers the start state (or the slots run out)
    done = False
    while time.perf_counter() - t0 < args.budget and not done:
        ag.run(args.chunk)
        steps += args.chunk
        if ag.manage():
            done = start_covered() or ag.n_active >= args.options - 1
            torch.cuda.synchronize()
            c = ag.counters()
            log.append(dict(t=round(time.perf_counter() - t0, 3), steps=steps, n_active=ag.n_active, goals=c["goals"],
                            episodes=c["episodes"]))
            print(json.dumps(log[-1]), flush=True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    c = ag.counters()
    covered = start_covered()
    print(json.dumps(dict(map=args.map, batch=args.batch, graph=args.graph, wall_s=round(wall, 3), steps=steps,
                          env_steps=steps * args.batch, n_active=ag.n_active, chain_complete=done,
                          start_covered=covered, goals=c["goals"], episodes=c["episodes"],
                          mean_return=c["mean_return"], n_success=c["n_success"].tolist(), promotions=log)))


if __name__ == "__main__":
    main()
