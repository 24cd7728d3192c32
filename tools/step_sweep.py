#!/usr/bin/env python
"""BASELINE.json configs[3]: step-only throughput sweep of the batched Pinball step (K1, scg_step) over batch sizes,
on the easy and the hard map, against the HBM roofline (44 algorithmic bytes per env-step).  One JSON line per point.

    python tools/step_sweep.py [--min 4096] [--max 16777216] [--iters 20]
Envs are independent, so N GPUs run N copies of this with no exchange (the multi-GPU number is N x the single one).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min", type=int, default=4096)
    ap.add_argument("--max", type=int, default=1 << 24)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--grid", type=int, default=0, help="broad-phase grid resolution (0 = library default)")
    args = ap.parse_args()
    import torch
    import skill_chaining_with_graphs_b200 as scg
    from skill_chaining_with_graphs_b200._lib import check, ptr, current_stream
    lib = scg.load_library()
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    for name in ("easy", "hard"):
        gmap = scg.PinballMap.from_name(name, args.grid)
        rng = np.random.default_rng(0)
        base = gmap.sample_free_states(rng, 1 << 16)
        B = args.min
        while B <= args.max:
            reps = (B + len(base) - 1) // len(base)
            S = torch.as_tensor(np.tile(base, (reps, 1))[:B]).cuda().t().contiguous()
            S2 = torch.empty_like(S)
            A = torch.randint(0, 5, (B,), dtype=torch.int32, device="cuda")
            r = torch.empty(B, device="cuda")
            f = torch.empty(B, dtype=torch.int32, device="cuda")

            def step(src, dst):
                check(lib.scg_step(gmap.handle, B, ptr(src[0]), ptr(src[1]), ptr(src[2]), ptr(src[3]), ptr(A), ptr(dst[0]),
                                   ptr(dst[1]), ptr(dst[2]), ptr(dst[3]), ptr(r), ptr(f), 1, current_stream()))
            for _ in range(3):
                step(S, S2)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.iters):          # ping-pong so every launch steps fresh states from HBM
                step(S, S2) if i % 2 == 0 else step(S2, S)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            rate = B / (ms * 1e-3)
            print(json.dumps({"config": "configs[3] step-only sweep", "map": name, "envs": B, "ms_per_launch": ms,
                              "env_steps_per_s": rate, "algorithmic_GBps": rate * 44 / 1e9, "hbm_peak_GBps": peak,
                              "frac_of_hbm_peak": rate * 44 / 1e9 / peak, "edges": gmap.n_edges, "grid": args.grid or 64,
                              "grid_candidates": gmap.n_candidates}), flush=True)
            del S, S2, A, r, f
            B *= 4


if __name__ == "__main__":
    main()
