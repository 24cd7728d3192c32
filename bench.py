#!/usr/bin/env python
"""bench.py - Pinball env-steps/s including option Q-eval + Sarsa(lambda) update (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one lock-step agent step over the whole env batch: a fused kernel does env step (K1) ->
initiation classifiers + Q evaluation + eps-greedy + TD error (K2+K4) -> 32-byte step record, for the
`sync_interval` consecutive steps of a window in one launch (nothing couples the envs while the weights are
frozen); then the window's records are folded into the traces and weight deltas by one Sarsa(lambda) sweep
(K3, forward-view form), followed by the cross-GPU exchange and the weight apply.
Workload at every N: BASELINE.json configs[1] per GPU - Pinball 'easy', 65,536 envs, order-3 Fourier
basis, 4 option slots with 2 active logistic initiation classifiers (weak scaling: each rank owns
its own 65,536-env slice).  Synthetic data: random free-space start states, random-init weights.

Prints ONE JSON line (rank 0).  `value` = env-steps of all ranks / max-over-ranks device time with
state resident in HBM; `e2e` = same through SkillChainAgent.step_host (host buffers, H2D and D2H
inside the timed region); `roofline` = the window trace-sweep kernel's algorithmic bytes / its CUDA-event
time against the measured HBM copy peak; `cpu_baseline` = the NumPy oracle timed on this box.
--impl reference times the stand-in reference (the NumPy oracle: the reference repository has no
code) on all host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Pinball env-steps/s incl. option Q-eval+Sarsa(lambda) update"
UNIT = "env-steps/s"
MANAGE_EVERY = 64          # steps between calls of the option-creation controller (SkillChainAgent.manage)


def workload(args):
    return dict(map=args.map, batch=args.batch, order=args.order, max_options=args.options,
                sync_interval=args.sync_interval, gamma=0.99, lam=0.9, alpha=1e-3, epsilon=0.05, seed=0,
                option_timeout=250, max_episode_steps=2000, graph=bool(getattr(args, "graph", False)))


def gpu_only(args):
    """AgentConfig fields that only the GPU backend has."""
    return dict(sync_backend=args.sync_backend, window=args.window)


def config_json(args, n_gpus):
    F = (args.order + 1) ** 4
    return {
        "workload": f"configs[{4 if getattr(args, 'graph', False) else 1}]{' option-graph variant' if getattr(args, 'graph', False) else ''}: "
                    f"Pinball '{args.map}', {args.batch} envs per GPU, order-{args.order} Fourier "
                    f"(F={F}), {args.options} option slots (2 active initiation classifiers), "
                    f"sync every {args.sync_interval} steps",
        "controller": f"SkillChainAgent.manage() every {MANAGE_EVERY} steps inside the timed region",
        "envs_per_gpu": args.batch, "global_envs": args.batch * n_gpus, "order": args.order,
        "options": args.options, "sync_interval": args.sync_interval, "map": args.map,
        "parallelism": f"env-sharded x{n_gpus}, sum of (dW, cnt) over ranks every sync interval "
                       f"({'one NVLink peer-memory exchange+apply kernel' if getattr(args, 'sync_backend', 'p2p') == 'p2p' else 'NCCL all-reduce'})",
        "l2": f"traces {args.batch * 5 * F * 4 / 2**20:.0f} MiB per GPU > 126 MiB L2, swept once per window between 8 "
              "step kernels (inputs larger than L2; no explicit flush)",
    }


def setup_classifiers(theta):
    """Two synthetic active options so the classifier / termination / re-selection path runs:
    option 0 accepts x >= 0.6, option 1 accepts y <= 0.45 (quadratic features unused)."""
    theta[:] = 0
    theta[0, :3] = [-6.0, 10.0, 0.0]
    theta[1, :3] = [4.5, 0.0, -10.0]


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU legs (the only places that execute oracle/) ----------------------------------------------
def _cpu_worker(wl, batch, steps, seed, q, warmup=1):
    import oracle
    wl = dict(wl, batch=batch, seed=seed)
    ag = oracle.SkillChainAgent(oracle.AgentConfig(**wl))
    rng = np.random.default_rng(seed)
    ag.env.reset(states=ag.map.sample_free_states(rng, batch))
    ag.options.W[:] = (rng.standard_normal(ag.options.W.shape) * 0.1).astype(np.float32)
    setup_classifiers(ag.options.theta)
    ag.active[:2] = True
    ag.n_active = 2
    ag.parents[1], ag.parents[2] = (np.uint32(1 | (1 << 31)), np.uint32(3 | (1 << 31))) if wl.get("graph") else (1, 2)
    for _ in range(max(warmup, 1)):             # untimed warm-up
        ag.step()
    ag.manage()
    t0 = time.perf_counter()
    for i in range(steps):
        ag.step()
        if (i + 1) % MANAGE_EVERY == 0:         # the option-creation controller, as in run_episode
            ag.manage()
    q.put((batch * steps, time.perf_counter() - t0))


def cpu_oracle_rate(wl, batch, steps, procs, warmup=1):
    """env-steps/s of the NumPy oracle agent: `procs` processes, each its own `batch`-env agent."""
    import multiprocessing as mp
    if procs > 1:      # one single-threaded oracle per core instead of oversubscribed BLAS pools
        for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[v] = "1"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_cpu_worker, args=(wl, batch, steps, 1000 + i, q, warmup)) for i in range(procs)]
    t0 = time.perf_counter()
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return total / slowest, wall


def run_reference(args):
    """The reference arm: the stand-in reference (NumPy oracle) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    procs = max(1, cores)
    # bounded sample: size each step (one pass over `batch` envs per core) so K steps take about a minute
    batch = int(min(args.cpu_batch, max(256, (60 * 12000 // max(args.steps, 1)) // 256 * 256)))
    wl = workload(args)
    t_all = time.perf_counter()
    rate, _ = cpu_oracle_rate(wl, batch, args.steps, procs, warmup=min(args.warmup, 10))
    wall = time.perf_counter() - t_all
    ms = 1e3 * (batch * procs) / rate
    sample = (f"{procs} processes x {batch} envs x {args.steps} steps of the same workload "
              f"(NumPy oracle SkillChainAgent.step; the reference repository has no code)")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_json(args, args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm -----------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import skill_chaining_with_graphs_b200 as scg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise scg.ScgError("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = scg.load_library()
    wl = workload(args)
    B = args.batch
    gmap = scg.PinballMap.from_name(args.map)
    rng = np.random.default_rng(1234 + rank)
    S = gmap.sample_free_states(rng, B)
    cfg = scg.AgentConfig(**wl, env_offset=rank * B, **gpu_only(args))
    ag = scg.SkillChainAgent(cfg, gmap, initial_states=S)
    if world > 1 and args.sync_backend == "p2p" and ag._xchg is None:
        args.sync_backend = "nccl"                      # the agent fell back (it said why on stderr): report what ran
    wrng = np.random.default_rng(7)                     # same weights on every rank
    ag.options.set_weights((wrng.standard_normal(tuple(ag.options.W.shape)) * 0.1).astype(np.float32))
    theta = np.zeros((args.options, 6), dtype=np.float32)
    setup_classifiers(theta)
    ag.options.theta.copy_(torch.as_tensor(theta))
    ag.active_mask, ag.n_active = 3, 2
    GOAL = 1 << 31
    ag.parents_host[1], ag.parents_host[2] = (1 | GOAL, 3 | GOAL) if args.graph else (1, 2)
    ag._push_parents()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_warm = max(args.warmup, 16)            # at least two windows, so the gestating option has qualified by the end
    ag.run(n_warm)
    ag.warm_up_controller()
    ag.manage()              # with these classifiers the gestating option qualifies within the warm-up: promote it here
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region: exactly K steps, device-resident ----
    launches0 = lib.scg_launch_count()
    # CUDA events around the dominant kernel (the window sweep), live in the timed region.  Events around every launch
    # would cost ~5 us per step here (each record is a stream operation between back-to-back kernels): the other
    # stages are timed in a short separate pass below.
    ag.profile_begin(args.steps + 16, kinds=(1,))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for lo in range(0, args.steps, MANAGE_EVERY):               # full skill chaining: steps + the controller
        ag.run(min(MANAGE_EVERY, args.steps - lo))
        ag.manage()
    ag.flush()                                                  # a partial last window is swept inside the timed region too
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kind_ms, kind_n = ag.profile_end()
    launches = lib.scg_launch_count() - launches0
    n_side = 8 * max(args.sync_interval, ag.win_cap)
    ag.profile_begin(4 * n_side + 16, kinds=(0, 2, 3))          # untimed side pass: step kernel, reduction, apply / exchange
    ag.run(n_side)
    torch.cuda.synchronize()
    side_ms, side_n = ag.profile_end()
    for k in (0, 2, 3):
        kind_ms[k], kind_n[k] = side_ms[k] * args.steps / n_side, side_n[k] * args.steps // n_side
    # ---- e2e: same K steps through the host-buffer API ----
    hs = ag.s.cpu().numpy().copy()
    ha = ag.action.cpu().numpy().copy()
    for _ in range(3):
        s2, r, f, a2, d = ag.step_host(hs, ha)
        hs, ha = s2, a2
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        s2, r, f, a2, d = ag.step_host(hs, ha)
        hs, ha = s2, a2                              # views of the pinned result buffers, fed straight back
        if (i + 1) % MANAGE_EVERY == 0:
            ag.manage()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    # informational: the same loop for a caller that needs the state only once per sync interval (run_host)
    T = args.sync_interval
    hs, ha = ag.s.cpu().numpy().copy(), ag.action.cpu().numpy().copy()
    for _ in range(2):
        hs, ha, r, f, d = ag.run_host(hs, ha, T)
    barrier()
    t0 = time.perf_counter()
    n_calls = max(args.steps // T, 1)
    for _ in range(n_calls):
        hs, ha, r, f, d = ag.run_host(hs, ha, T)
    torch.cuda.synchronize()
    e2e_win_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, e2e_ms, e2e_win_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, e2e_win_ms = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        F = (args.order + 1) ** 4
        total_env_steps = B * world * args.steps
        value = total_env_steps / (ms * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, which = 6650.0, "fallback (B200_PROFILING.md)"
        T = ag.win_cap
        bytes_per_env = 8 * 5 * F + 32 * T              # trace read + write once per window, T 32-byte records
        avg = lambda k: kind_ms[k] / kind_n[k] if kind_n[k] else 0.0
        k3_ms = avg(1)
        achieved = B * bytes_per_env / (k3_ms * 1e-3) / 1e9 if k3_ms > 0 else 0.0
        tot_ms = sum(kind_ms)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "k3_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                traffic = tj.get("dram_bytes_per_launch") if tj.get("kernel") == "k_window" else None
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_json(args, world),
            "roofline": {"kernel": "k_window (K3 Sarsa(lambda) trace sweep, forward-view window form)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": which, "algorithmic_bytes_per_launch": B * bytes_per_env,
                         "bytes_per_env_step": bytes_per_env / T, "window_steps": T,
                         "avg_launch_ms": k3_ms, "launches": kind_n[1],
                         "share_of_step": (kind_ms[1] / tot_ms) if tot_ms else None},
            "stages_ms_per_step": {"fused_step_k1_k2_k4": kind_ms[0] / args.steps, "k3_window_sweep": kind_ms[1] / args.steps,
                                   "dw_reduce": kind_ms[2] / args.steps, "apply": kind_ms[3] / args.steps,
                                   "fused_step_avg_launch_ms": avg(0), "steps_per_fused_launch": args.sync_interval,
                                   "note": "k3_window_sweep timed live in the timed region; the other stages in a separate "
                                           f"{n_side}-step pass (events around every launch cost ~5 us per step)"},
            # north_star also asks for the FP32 view of the feature + Q part (K2): 18 F flop per env-step, counted
            # against the nominal FP32 FMA rate at the maximum SM clock (148 SMs x 128 lanes x 2 flop)
            "k2_fp32": {"kernel": "k_agent_step (K1+K2+K4; only K2's 18*F flop per env-step are counted)",
                        "flop_per_env_step": 18 * F, "achieved_tflops": B * 18 * F / (kind_ms[0] / args.steps * 1e-3) / 1e12
                        if kind_ms[0] > 0 else None, "peak_tflops_nominal": 148 * 128 * 2 * 1.965e9 / 1e12},
            "e2e": {"value": total_env_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": B * ag.HOST_H2D_BYTES_PER_ENV,
                    "d2h_bytes_per_step": B * ag.HOST_D2H_BYTES_PER_ENV,
                    "api": "SkillChainAgent.step_host -> scg_agent_step_host (pinned host buffers)"},
            "e2e_per_sync_interval": {"value": B * world * n_calls * T / (e2e_win_ms * 1e-3), "unit": UNIT,
                                      "h2d_bytes_per_call": B * 20, "d2h_bytes_per_call": B * 32, "steps_per_call": T,
                                      "api": "SkillChainAgent.run_host: host state in / out once per sync interval "
                                             "(informational; `e2e` above copies every step)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            cores = max(1, len(os.sched_getaffinity(0)))
            rate, wall = cpu_oracle_rate(wl, args.cpu_batch, args.cpu_steps, cores)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"NumPy oracle SkillChainAgent.step, {args.cpu_batch} envs x {args.cpu_steps} steps of the "
                          f"same workload in each of {cores} single-threaded processes ({wall:.1f} s wall)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--map", default="easy")
    ap.add_argument("--batch", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--options", type=int, default=4)
    ap.add_argument("--sync-interval", type=int, default=8)
    ap.add_argument("--cpu-batch", type=int, default=4096)
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--sync-backend", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--window", type=int, default=0, help="steps per trace sweep (0 = min(sync interval, 8))")
    ap.add_argument("--graph", action="store_true", help="option-graph variant (configs[4]): an option's targets are the "
                    "initiation sets of ALL earlier options and the goal, so chains merge")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
