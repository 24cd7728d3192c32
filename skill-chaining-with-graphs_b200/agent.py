"""Batched skill-chaining agent on B200: the fused lock-step step (one kernel: K1 + K2 + K4), the
windowed Sarsa(lambda) sweep (K3, once per window), the low-rate option-creation controller, and the
cross-GPU weight-delta sync.  Same interface and semantics as the CPU oracle's SkillChainAgent
(oracle/agent.py, which stands in for the reference - the reference has no code:
/root/reference/README.md:1-2).

    agent = SkillChainAgent(AgentConfig(map="easy", batch=65536, max_options=4))
    stats = agent.run_episode(max_steps=2000)

Multi-GPU: one process per GPU, each owning a contiguous env slice (`env_offset`); every
`sync_interval` steps dW and cnt are all-reduced (sum) over NCCL and every rank applies the same
update (weights stay bit-identical across ranks).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check, ptr, AgentStruct, GOAL_BIT, N_ACTIONS
from .option import OptionSet
from .pinball import PinballMap
from .sync import allreduce_deltas, allreduce_scalar_sum, world_size


@dataclass
class AgentConfig:
    map: str = "easy"
    batch: int = 1
    order: int = 3
    max_options: int = 4
    gamma: float = 0.99
    lam: float = 0.9
    alpha: float = 1e-3
    epsilon: float = 0.05
    sync_interval: int = 1
    seed: int = 0
    env_offset: int = 0
    option_bonus: float = 1000.0
    option_timeout: int = 250
    max_episode_steps: int = 2000
    gestation_successes: int = 32
    example_capacity: int = 4096
    clf_steps: int = 200
    clf_lr: float = 1.0
    graph: bool = False
    cull: bool = True
    window: int = 0          # steps per trace sweep; 0 = min(sync_interval, 8)
    deterministic: bool = False   # fixed-order reduction of the weight deltas: runs reproduce bit for bit (slightly slower)
    sync_backend: str = "p2p"   # multi-rank weight-delta exchange: "p2p" (one kernel over NVLink peer memory) or "nccl"


class SkillChainAgent:
    def __init__(self, cfg, pmap=None, process_group=None, initial_states=None):
        import torch
        if not torch.cuda.is_available():
            raise _lib.ScgError("SkillChainAgent needs a CUDA device: there is no CPU fallback")
        self.torch = torch
        self.lib = _lib.load()
        self.cfg = cfg
        self.map = pmap if pmap is not None else PinballMap.from_name(cfg.map)
        self.pg = process_group
        B, K = cfg.batch, cfg.max_options
        dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.options = OptionSet(K, cfg.order, B, cfg.gamma, cfg.lam, cfg.alpha, cfg.epsilon, cfg.seed,
                                 cfg.env_offset, dev, deterministic=cfg.deterministic)
        self.options._pre_read = self.flush        # dW / trace reads see the open window folded in
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.win_cap = int(cfg.window) if cfg.window else max(1, min(int(cfg.sync_interval), 8))
        if not 1 <= self.win_cap <= _lib.WIN_MAX:
            raise ValueError(f"window must be in 1..{_lib.WIN_MAX}")
        self._sbuf = (torch.zeros((4, B), **f32), torch.zeros((4, B), **f32))
        # reward, flags, action, delta back to back: the host-buffer step returns them with one copy
        self._out = torch.zeros((4, B), **f32)
        self.reward, self.delta = self._out[0], self._out[3]
        self.flags, self.action = self._out[1].view(torch.int32), self._out[2].view(torch.int32)
        self.option = torch.zeros(B, **i32)
        self.t_opt = torch.zeros(B, **i32)
        self.ep_steps = torch.zeros(B, **i32)
        self.start_xy = torch.zeros((B, 2), **f32)
        self.ep_return = torch.zeros(B, **f32)
        self.q_carry = torch.zeros(B, **f32)
        self.win_rec = torch.zeros((self.win_cap, B, 8), **f32)
        self.parents = torch.zeros(K, dtype=torch.int32, device=dev)
        self.parents_host = np.zeros(K, dtype=np.uint32)
        self.parents_host[0] = GOAL_BIT
        self._push_parents()
        self.ex_xy = torch.zeros((K, cfg.example_capacity, 2), **f32)
        self.ex_label = torch.zeros((K, cfg.example_capacity), dtype=torch.uint8, device=dev)
        self.ex_count = torch.zeros(K, **i32)
        self.n_success = torch.zeros(K, **i32)
        self.n_fail = torch.zeros(K, **i32)
        self.n_success_global = torch.zeros(K, **i32)
        self.stats = torch.zeros(4, dtype=torch.int64, device=dev)
        self.active_mask = 0
        self.n_active = 0
        # initial state, option and action (oracle/agent.py __init__)
        s = self._sbuf[0]
        if initial_states is not None:
            st = torch.as_tensor(np.asarray(initial_states, dtype=np.float32)).to(dev).reshape(B, 4)
            s.copy_(st.t())
        else:
            check(self.lib.scg_reset(self.map.handle, B, None, ptr(s[0]), ptr(s[1]), ptr(s[2]), ptr(s[3]), cfg.seed, 0,
                                     cfg.env_offset, _lib.current_stream()))
        self.start_xy.copy_(s[:2].t())
        self.options.pack()
        self.action.copy_(self.options.act(None, self.option, step=0xFFFFFFFF, stream=_lib.STREAM_RESELECT, soa=s))
        self._struct = self._make_struct()
        self.options._on_weights_changed = self._weights_changed
        self._xchg = None
        if world_size(self.pg) > 1 and cfg.sync_backend == "p2p":
            # every rank must end up on the same backend: agree on whether the peer mapping worked everywhere
            try:
                x, err = self._connect_peers(), None
            except Exception as e:          # e.g. CUDA IPC unavailable between these devices
                x, err = None, e
            ok = torch.tensor([1 if x is not None else 0], dtype=torch.int32, device=dev)
            ok = int(allreduce_scalar_sum(ok, self.pg))
            if ok == world_size(self.pg):
                self._xchg = x
            else:
                if x is not None:
                    self.lib.scg_xchg_destroy(x)
                import sys
                print(f"[scg] peer-memory sync unavailable on {world_size(self.pg) - ok} rank(s) ({err}); "
                      "using the NCCL all-reduce path", file=sys.stderr)

    def _connect_peers(self):
        """Create this rank's exchange buffer and map every peer's through CUDA IPC (handles travel over
        torch.distributed once).  The handle exchange is collective, so a rank that failed locally still takes part
        and every rank then sees the failure."""
        dist = self.torch.distributed
        rank, world = dist.get_rank(self.pg), dist.get_world_size(self.pg)
        x, mine, err = C.c_void_p(), None, None
        try:
            check(self.lib.scg_xchg_create(self.options.ctx, rank, world, C.byref(x)))
            buf = (C.c_ubyte * self.lib.scg_xchg_handle_bytes())()
            check(self.lib.scg_xchg_handle(x, buf))
            mine = bytes(buf)
        except Exception as e:
            err = e
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=self.pg)
        try:
            if err is not None or any(h is None for h in handles):
                raise err or _lib.ScgError("a peer could not export its exchange buffer")
            check(self.lib.scg_xchg_connect(x, C.c_char_p(b"".join(handles))))
        except Exception:
            if x:
                self.lib.scg_xchg_destroy(x)
            raise
        return x

    def __del__(self):
        try:
            if getattr(self, "_xchg", None):
                self.lib.scg_xchg_destroy(self._xchg)
                self._xchg = None
        except Exception:
            pass

    # -- plumbing --------------------------------------------------------------------------------
    def _push_parents(self):
        self.parents.copy_(self.torch.from_numpy(self.parents_host.view(np.int32)))

    def _make_struct(self):
        cfg, o, g = self.cfg, self.options, AgentStruct()
        g.B, g.K, g.order = cfg.batch, o.K, o.order
        g.env_offset, g.example_capacity, g.seed = cfg.env_offset, cfg.example_capacity, cfg.seed
        g.gamma, g.lam, g.epsilon, g.option_bonus = cfg.gamma, cfg.lam, cfg.epsilon, cfg.option_bonus
        g.option_timeout, g.max_episode_steps, g.cull = cfg.option_timeout, cfg.max_episode_steps, int(cfg.cull)
        g.alpha, g.win_cap = cfg.alpha, self.win_cap
        g.step = g.window_steps = g.win_len = g.carry_valid = 0
        s, s2 = self._sbuf
        g.x, g.y, g.vx, g.vy = (s[i].data_ptr() for i in range(4))
        g.x2, g.y2, g.vx2, g.vy2 = (s2[i].data_ptr() for i in range(4))
        for name in ("action", "option", "t_opt", "ep_steps", "start_xy", "ep_return", "reward", "flags", "delta",
                     "q_carry", "win_rec", "parents", "ex_xy", "ex_label", "ex_count", "n_success", "n_fail",
                     "n_success_global", "stats"):
            setattr(g, name, getattr(self, name).data_ptr())
        g.trace, g.W, g.Wt, g.theta = o._trace.data_ptr(), o.W.data_ptr(), o.Wt.data_ptr(), o.theta.data_ptr()
        g.dW, g.cnt = o._dW.data_ptr(), o.cnt.data_ptr()
        return g

    def _sync_struct(self):
        """Push the host-side controller state the kernels read."""
        g = self._struct
        g.n_active, g.active_mask = self.n_active, self.active_mask
        return g

    @property
    def t(self):
        return int(self._struct.step)

    @property
    def s(self):
        """(4, B) SoA tensor holding the current state."""
        a, b = self._sbuf
        return a if self._struct.x == a[0].data_ptr() else b

    @property
    def state(self):
        """(B, 4) copy of the current state."""
        return self.s.t().contiguous()

    def _weights_changed(self, applied=False):
        """Hook called by the OptionSet when its weights change (set_weights, apply)."""
        self._struct.carry_valid = 0
        if applied:
            self._struct.window_steps = 0

    def invalidate(self):
        """Call after changing state, action, option or weights from outside: the carried Q_o(s, a)
        is stale and the next step re-evaluates it."""
        self._struct.carry_valid = 0

    # -- the hot path ----------------------------------------------------------------------------
    def step(self):
        """One lock-step agent step for the whole batch (oracle/agent.py SkillChainAgent.step)."""
        self.run(1)

    def run(self, n_steps):
        """`n_steps` lock-step agent steps, with the weight sync every `sync_interval` steps.  The whole loop
        runs inside the library (scg_agent_run): on one rank the sync is the apply kernel, across ranks the
        peer-memory exchange kernel.  With sync_backend="nccl" it returns to Python at every sync for the
        NCCL all-reduce."""
        g = self._sync_struct()
        st = _lib.current_stream()
        n, T = int(n_steps), int(self.cfg.sync_interval)
        if world_size(self.pg) == 1 or self._xchg is not None:
            check(self.lib.scg_agent_run(self.map.handle, self.options.ctx, C.byref(g), n, T, self._xchg, st))
            self.options.window_steps = int(g.window_steps)
            return
        while n > 0:
            k = min(n, T - int(g.window_steps))
            check(self.lib.scg_agent_run(self.map.handle, self.options.ctx, C.byref(g), k, 0, None, st))
            n -= k
            if int(g.window_steps) >= T:
                self.sync()

    def flush(self):
        """Fold the open window's step records into dW and the traces (scg_agent_flush)."""
        g = self._struct
        if int(g.win_len) > 0:
            check(self.lib.scg_agent_flush(self.options.ctx, C.byref(g), _lib.current_stream()))

    def step_host(self, state, action):
        """The same step for a caller that keeps state and actions in host memory (as a user of the
        oracle's NumPy agent does): `state` float32 (4, B) SoA and `action` int32 (B,) are copied
        host->device, the fused step runs, and (next_state (4, B), reward, flags, next_action,
        td_error) come back as views of pinned host buffers that the next call overwrites.  Traces
        and weights stay on the device."""
        torch = self.torch
        B = self.cfg.batch
        if getattr(self, "_host", None) is None:
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            def mk():
                out = pin((4, B), torch.float32)      # reward, flags, next action, TD error rows, back to back
                return dict(s=pin((4, B), torch.float32), a=pin((B,), torch.int32), s2=pin((4, B), torch.float32),
                            r=out[0], f=out[1].view(torch.int32), a2=out[2].view(torch.int32), d=out[3])
            self._host = [mk(), mk()]                       # ping-pong: results of call i are inputs of call i+1
            self._host_np = [{k: v.numpy() for k, v in h.items()} for h in self._host]
            self._host_i = 0
        prev, prev_np = self._host[self._host_i], self._host_np[self._host_i]
        cur, cur_np = self._host[1 - self._host_i], self._host_np[1 - self._host_i]
        if state is prev_np["s2"]:
            src_s = prev["s2"]                              # previous result fed straight back: already pinned
        else:
            cur_np["s"][...] = state
            src_s = cur["s"]
        if action is prev_np["a2"]:
            src_a = prev["a2"]
        else:
            cur_np["a"][...] = action
            src_a = cur["a"]
        g = self._sync_struct()
        check(self.lib.scg_agent_step_host(self.map.handle, self.options.ctx, C.byref(g), ptr(src_s), ptr(src_a),
                                           ptr(cur["s2"]), ptr(cur["r"]), ptr(cur["f"]), ptr(cur["a2"]), ptr(cur["d"]),
                                           _lib.current_stream()))
        self._host_i = 1 - self._host_i
        self.options.window_steps = int(g.window_steps)
        if int(g.window_steps) >= self.cfg.sync_interval:
            self.sync()
        return cur_np["s2"], cur_np["r"], cur_np["f"], cur_np["a2"], cur_np["d"]

    def run_host(self, state, action, n_steps):
        """`n_steps` steps for a caller that only needs the state every `n_steps` steps: `state` (4, B) and `action`
        (B,) go host->device once, the steps run device-resident (with the syncs that fall due), and (state, action,
        last reward, last flags, last td_error) come back once, as views of pinned host buffers that the next call
        overwrites (feeding them straight back skips the host-side copy)."""
        torch = self.torch
        B = self.cfg.batch
        if getattr(self, "_hostw", None) is None:
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            h = dict(s=pin((4, B), torch.float32), a=pin((B,), torch.int32), s2=pin((4, B), torch.float32),
                     out=pin((4, B), torch.float32))
            h["a2"] = h["out"][2].view(torch.int32)
            self._hostw = h
            self._hostw_np = dict(s=h["s"].numpy(), a=h["a"].numpy(), s2=h["s2"].numpy(), a2=h["a2"].numpy(),
                                  r=h["out"][0].numpy(), f=h["out"][1].view(torch.int32).numpy(), d=h["out"][3].numpy())
        h, hn = self._hostw, self._hostw_np
        if state is hn["s2"]:
            src_s = h["s2"]
        else:
            hn["s"][...] = state
            src_s = h["s"]
        if action is hn["a2"]:
            src_a = h["a2"]
        else:
            hn["a"][...] = action
            src_a = h["a"]
        self.s.copy_(src_s, non_blocking=True)
        self.action.copy_(src_a, non_blocking=True)
        self.invalidate()
        self.run(n_steps)
        h["s2"].copy_(self.s, non_blocking=True)
        h["out"].copy_(self._out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return hn["s2"], hn["a2"], hn["r"], hn["f"], hn["d"]

    HOST_H2D_BYTES_PER_ENV = 20      # state 16 + action 4
    HOST_D2H_BYTES_PER_ENV = 32      # next state 16 + reward 4 + flags 4 + next action 4 + TD error 4

    def profile_begin(self, max_events, kinds=(0, 1, 2, 3)):
        """Record CUDA events around the launches of the given kinds (0 step, 1 window sweep, 2 reduce, 3 apply)."""
        check(self.lib.scg_profile_begin(self.options.ctx, int(max_events), sum(1 << k for k in kinds)))

    def profile_end(self):
        """-> (ms per kind, launches per kind) for [fused step, window sweep, dW reduction, apply]."""
        ms = (C.c_float * 4)()
        n = (C.c_int * 4)()
        check(self.lib.scg_profile_end(self.options.ctx, ms, n))
        return [float(v) for v in ms], [int(v) for v in n]

    def sync(self):
        """Flush the window, sum its dW / cnt over ranks (if any) and apply."""
        o, g = self.options, self._struct
        self.flush()
        if self._xchg is not None:
            check(self.lib.scg_xchg_sync(self._xchg, o.order, o.K, ptr(o.W), ptr(o.Wt), ptr(o._dW), ptr(o.cnt),
                                         self.cfg.alpha, max(int(g.window_steps), 1), ptr(self.n_success),
                                         ptr(self.n_success_global), _lib.current_stream()))
            o.window_steps = 0
        else:
            allreduce_deltas(o._dW, o.cnt, self.pg)
            o.window_steps = int(g.window_steps)
            o.apply()
        g.window_steps = 0
        g.carry_valid = 0

    def peer_sync_timed_out(self):
        """True if a peer failed to show up at some sync (the exchange kernel gives up after ~2 s)."""
        if self._xchg is None:
            return False
        v = C.c_int()
        check(self.lib.scg_xchg_status(self._xchg, C.byref(v)))
        return bool(v.value)

    # -- low-rate controller ---------------------------------------------------------------------
    def examples(self, k):
        n = int(min(int(self.ex_count[k]) & 0xFFFFFFFF, self.cfg.example_capacity))
        return self.ex_xy[k, :n].clone(), self.ex_label[k, :n].clone()

    def _fit_slot(self, g, X, y):
        """Fit option g's initiation classifier on (X, y); across ranks every rank fits on its own examples and the
        parameters are averaged."""
        cfg = self.cfg
        self.options.theta[g].zero_()
        if X.shape[0] > 0:
            self.options.fit_initiation(g, X, y, cfg.clf_steps, cfg.clf_lr)
        ws = world_size(self.pg)
        if ws > 1:
            th = allreduce_scalar_sum(self.options.theta[g].clone(), self.pg)
            self.options.theta[g].copy_(th / ws)

    def manage(self):
        """Promote the gestating option once it has enough successes (oracle/agent.py manage).
        Multi-GPU: success counts are summed over ranks so every rank promotes at the same step;
        each rank fits on its own examples and theta is averaged."""
        cfg, K, torch = self.cfg, self.options.K, self.torch
        g = self.n_active
        if g >= K - 1:
            return False
        if self._xchg is not None:
            # the peer-memory sync already left the sum over ranks (as of the last sync) on every GPU: same
            # decision everywhere, no collective in the loop
            n_succ = int(self.n_success_global[g])
        else:
            # the device counter is 32 bits and wraps: read it as unsigned, sum over ranks in 64 bits
            n_succ = int(allreduce_scalar_sum(self.n_success[g:g + 1].to(torch.int64) & 0xFFFFFFFF, self.pg))
        if n_succ < cfg.gestation_successes:
            return False
        X, y = self.examples(g)
        self._fit_slot(g, X, y)
        self.active_mask |= (1 << g)
        self.n_active += 1
        n = self.n_active
        self.parents_host[n] = ((1 << n) - 1) | GOAL_BIT if cfg.graph else (1 << (n - 1))
        self._push_parents()
        return True

    def warm_up_controller(self):
        """Run the controller's device code path once on scratch data (torch loads its kernels lazily and NCCL sets
        up a collective on first use: the first promotion would otherwise pay milliseconds inside the caller's loop)."""
        torch, K = self.torch, self.options.K
        saved = self.options.theta[K - 1].clone()
        X = torch.tensor([[0.1, 0.2], [0.8, 0.7], [0.3, 0.9], [0.6, 0.1]], device=self.device)
        y = torch.tensor([0, 1, 0, 1], dtype=torch.uint8, device=self.device)
        _ = int(allreduce_scalar_sum(self.n_success[0:1].to(torch.int64) & 0xFFFFFFFF, self.pg))
        _ = int(self.n_success_global[0])
        _ = self.examples(0)
        clf_steps, self.cfg.clf_steps = self.cfg.clf_steps, 1
        self._fit_slot(K - 1, X, y)
        self.cfg.clf_steps = clf_steps
        self.options.theta[K - 1].copy_(saved)
        self._push_parents()
        torch.cuda.synchronize()

    # -- checkpoint / resume ----------------------------------------------------
    _CKPT_TENSORS = ("action", "option", "t_opt", "ep_steps", "start_xy", "ep_return", "ex_xy", "ex_label", "ex_count",
                     "n_success", "n_fail", "stats", "q_carry")

    def save(self, path):
        """Write everything needed to resume this rank (option weights, classifiers, option graph, per-env state and
        traces, counters) to `path` (.npz).  The open window is folded in first."""
        self.flush()
        o, g = self.options, self._struct
        arrs = {k: getattr(self, k).cpu().numpy() for k in self._CKPT_TENSORS}
        arrs.update(W=o.W.cpu().numpy(), theta=o.theta.cpu().numpy(), trace=o._trace.cpu().numpy(),
                    dW=o._dW.cpu().numpy(), cnt=o.cnt.cpu().numpy(), state=self.s.cpu().numpy(),
                    parents=self.parents_host.copy(),
                    meta=np.array([self.n_active, self.active_mask, int(g.step), int(g.window_steps)], dtype=np.int64))
        np.savez(path, **arrs)

    def load(self, path):
        """Restore a checkpoint written by save() into an agent built with the same AgentConfig."""
        torch = self.torch
        z = np.load(path)
        o, g = self.options, self._struct
        if tuple(z["W"].shape) != tuple(o.W.shape) or z["state"].shape != tuple(self.s.shape):
            raise ValueError("checkpoint does not match this agent's configuration")
        self.flush()
        for k in self._CKPT_TENSORS:
            getattr(self, k).copy_(torch.as_tensor(z[k]))
        o.W.copy_(torch.as_tensor(z["W"]))
        o.theta.copy_(torch.as_tensor(z["theta"]))
        o._trace.copy_(torch.as_tensor(z["trace"]))
        o._dW.copy_(torch.as_tensor(z["dW"]))
        o.cnt.copy_(torch.as_tensor(z["cnt"]))
        o.pack()
        self.s.copy_(torch.as_tensor(z["state"]))
        self.parents_host[:] = z["parents"]
        self._push_parents()
        self.n_active, self.active_mask = int(z["meta"][0]), int(z["meta"][1])
        g.step, g.window_steps, g.win_len = int(z["meta"][2]), int(z["meta"][3]), 0
        o.window_steps = int(z["meta"][3])
        g.carry_valid = 0

    def counters(self):
        """Host copy of the global statistics: episodes, goals, mean finished return, per-option counts."""
        st = self.stats.cpu().numpy()
        ep = int(st[0])
        ret = float(st[2:3].view(np.float64)[0])
        return dict(episodes=ep, goals=int(st[1]), mean_return=(ret / ep) if ep else float("nan"),
                    n_success=self.n_success.cpu().numpy().view(np.uint32).astype(np.int64),
                    n_fail=self.n_fail.cpu().numpy().view(np.uint32).astype(np.int64),
                    n_active=self.n_active)

    def run_episode(self, max_steps=2000, manage_every=64):
        """Step until every env has finished at least one more episode (checked every
        `manage_every` steps, where the controller also runs), or `max_steps` steps."""
        base = int(self.stats[0])
        steps = 0
        B = self.cfg.batch
        while steps < max_steps:
            k = min(manage_every, max_steps - steps)
            self.run(k)
            steps += k
            self.manage()
            done = self.stats[0:1] - base >= B
            if world_size(self.pg) > 1:          # every rank must leave the loop at the same step (the syncs are collective)
                done = allreduce_scalar_sum(done.to(self.torch.int32), self.pg) == world_size(self.pg)
            if bool(done):
                break
        c = self.counters()
        return dict(steps=steps, finished=c["episodes"] - base, goals=c["goals"], mean_return=c["mean_return"],
                    n_active=self.n_active, env_steps=steps * B)
