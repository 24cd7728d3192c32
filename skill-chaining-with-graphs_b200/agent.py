"""Batched skill-chaining agent on B200: the fused lock-step step (K1 -> K2+K4 -> K3), the low-rate
option-creation controller, and the cross-GPU weight-delta sync.  Same interface and semantics as
the CPU oracle's SkillChainAgent (oracle/agent.py, which stands in for the reference - the
reference has no code: /root/reference/README.md:1-2).

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
from .sync import allreduce_deltas


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
                                 cfg.env_offset, dev)
        F = self.options.F
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self.s = torch.zeros((4, B), **f32)
        self.s2 = torch.zeros((4, B), **f32)
        self.action = torch.zeros(B, **i32)
        self.option = torch.zeros(B, **i32)
        self.t_opt = torch.zeros(B, **i32)
        self.ep_steps = torch.zeros(B, **i32)
        self.start_xy = torch.zeros((B, 2), **f32)
        self.ep_return = torch.zeros(B, **f32)
        self.reward = torch.zeros(B, **f32)
        self.flags = torch.zeros(B, **i32)
        self.delta = torch.zeros(B, **f32)
        self.rec = torch.zeros((B, 12), **f32)
        self.parents = torch.zeros(K, dtype=torch.int32, device=dev)
        self.parents_host = np.zeros(K, dtype=np.uint32)
        self.parents_host[0] = GOAL_BIT
        self._push_parents()
        self.ex_xy = torch.zeros((K, cfg.example_capacity, 2), **f32)
        self.ex_label = torch.zeros((K, cfg.example_capacity), dtype=torch.uint8, device=dev)
        self.ex_count = torch.zeros(K, **i32)
        self.n_success = torch.zeros(K, **i32)
        self.n_fail = torch.zeros(K, **i32)
        self.stats = torch.zeros(4, **i32)
        self.active_mask = 0
        self.n_active = 0
        self.t = 0
        self._swapped = False
        # initial state, option and action (oracle/agent.py __init__)
        if initial_states is not None:
            st = torch.as_tensor(np.asarray(initial_states, dtype=np.float32)).to(dev).reshape(B, 4)
            self.s.copy_(st.t())
        else:
            check(self.lib.scg_reset(self.map.handle, B, None, ptr(self.s[0]), ptr(self.s[1]), ptr(self.s[2]),
                                     ptr(self.s[3]), cfg.seed, 0, cfg.env_offset, _lib.current_stream()))
        self.start_xy.copy_(self.s[:2].t())
        self.options.pack()
        self.action.copy_(self.options.act(None, self.option, step=0xFFFFFFFF, stream=_lib.STREAM_RESELECT,
                                           soa=self.s))
        self._struct = AgentStruct()

    # -- plumbing --------------------------------------------------------------------------------
    def _push_parents(self):
        self.parents.copy_(self.torch.from_numpy(self.parents_host.view(np.int32)))

    @property
    def state(self):
        """(B, 4) copy of the current state."""
        return self.s.t().contiguous()

    def _fill_struct(self):
        cfg, o, g = self.cfg, self.options, self._struct
        g.B, g.K, g.order, g.n_active = cfg.batch, o.K, o.order, self.n_active
        g.active_mask, g.env_offset = self.active_mask, cfg.env_offset
        g.step, g.example_capacity, g.seed = self.t & 0xFFFFFFFF, cfg.example_capacity, cfg.seed
        g.gamma, g.lam, g.epsilon, g.option_bonus = cfg.gamma, cfg.lam, cfg.epsilon, cfg.option_bonus
        g.option_timeout, g.max_episode_steps, g.cull = cfg.option_timeout, cfg.max_episode_steps, int(cfg.cull)
        s, s2 = self.s, self.s2
        g.x, g.y, g.vx, g.vy = (s[i].data_ptr() for i in range(4))
        g.x2, g.y2, g.vx2, g.vy2 = (s2[i].data_ptr() for i in range(4))
        for name in ("action", "option", "t_opt", "ep_steps", "start_xy", "ep_return", "reward", "flags", "delta",
                     "rec", "parents", "ex_xy", "ex_label", "ex_count", "n_success", "n_fail", "stats"):
            setattr(g, name, getattr(self, name).data_ptr())
        g.trace, g.W, g.Wt, g.theta = o.trace.data_ptr(), o.W.data_ptr(), o.Wt.data_ptr(), o.theta.data_ptr()
        g.dW, g.cnt = o.dW.data_ptr(), o.cnt.data_ptr()
        return g

    # -- the hot path ----------------------------------------------------------------------------
    def step(self):
        """One lock-step agent step for the whole batch (oracle/agent.py SkillChainAgent.step)."""
        g = self._fill_struct()
        check(self.lib.scg_agent_step(self.map.handle, self.options.ctx, C.byref(g), _lib.current_stream()))
        self.s, self.s2 = self.s2, self.s
        self.options.tick()
        self.t += 1
        if self.t % self.cfg.sync_interval == 0:
            self.sync()

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
            self._host = dict(s=pin((4, B), torch.float32), a=pin((B,), torch.int32), s2=pin((4, B), torch.float32),
                              r=pin((B,), torch.float32), f=pin((B,), torch.int32), a2=pin((B,), torch.int32),
                              d=pin((B,), torch.float32))
            self._host_np = {k: v.numpy() for k, v in self._host.items()}
        h, hn = self._host, self._host_np
        hn["s"][...] = state
        hn["a"][...] = action
        g = self._fill_struct()
        check(self.lib.scg_agent_step_host(self.map.handle, self.options.ctx, C.byref(g), ptr(h["s"]), ptr(h["a"]),
                                           ptr(h["s2"]), ptr(h["r"]), ptr(h["f"]), ptr(h["a2"]), ptr(h["d"]),
                                           _lib.current_stream()))
        self.s, self.s2 = self.s2, self.s
        self.options.tick()
        self.t += 1
        if self.t % self.cfg.sync_interval == 0:
            self.sync()
        return hn["s2"], hn["r"], hn["f"], hn["a2"], hn["d"]

    HOST_H2D_BYTES_PER_ENV = 20      # state 16 + action 4
    HOST_D2H_BYTES_PER_ENV = 32      # next state 16 + reward 4 + flags 4 + next action 4 + TD error 4

    def profile_begin(self, max_steps):
        check(self.lib.scg_profile_begin(self.options.ctx, int(max_steps)))

    def profile_end(self):
        """-> (ms per stage [K1 step, K2+K4 control, K3 trace sweep, dW reduction], steps recorded)."""
        ms = (C.c_float * 4)()
        n = C.c_int()
        check(self.lib.scg_profile_end(self.options.ctx, ms, C.byref(n)))
        return [float(v) for v in ms], n.value

    def sync(self):
        """All-reduce the window's dW / cnt over ranks (if any) and apply."""
        o = self.options
        allreduce_deltas(o.dW, o.cnt, self.pg)
        o.apply()

    # -- low-rate controller ---------------------------------------------------------------------
    def examples(self, k):
        n = int(min(int(self.ex_count[k]), self.cfg.example_capacity))
        return self.ex_xy[k, :n].clone(), self.ex_label[k, :n].clone()

    def manage(self):
        """Promote the gestating option once it has enough successes (oracle/agent.py manage).
        Multi-GPU: success counts are summed over ranks so every rank promotes at the same step;
        each rank fits on its own examples and theta is averaged."""
        cfg, K, torch = self.cfg, self.options.K, self.torch
        g = self.n_active
        if g >= K - 1:
            return False
        n_succ = self.n_success[g:g + 1].clone()
        distributed = torch.distributed.is_available() and torch.distributed.is_initialized() \
            and torch.distributed.get_world_size() > 1
        if distributed:
            torch.distributed.all_reduce(n_succ, group=self.pg)
        if int(n_succ) < cfg.gestation_successes:
            return False
        X, y = self.examples(g)
        self.options.theta[g].zero_()
        if X.shape[0] > 0:
            self.options.fit_initiation(g, X, y, cfg.clf_steps, cfg.clf_lr)
        if distributed:
            th = self.options.theta[g].clone()
            torch.distributed.all_reduce(th, group=self.pg)
            self.options.theta[g].copy_(th / torch.distributed.get_world_size())
        self.active_mask |= (1 << g)
        self.n_active += 1
        n = self.n_active
        self.parents_host[n] = ((1 << n) - 1) | GOAL_BIT if cfg.graph else (1 << (n - 1))
        self._push_parents()
        return True

    def counters(self):
        """Host copy of the global statistics: episodes, goals, mean finished return, per-option counts."""
        st = self.stats.cpu().numpy()
        ep = int(st[0])
        ret = float(st[2:3].view(np.float32)[0])
        return dict(episodes=ep, goals=int(st[1]), mean_return=(ret / ep) if ep else float("nan"),
                    n_success=self.n_success.cpu().numpy().copy(), n_fail=self.n_fail.cpu().numpy().copy(),
                    n_active=self.n_active)

    def run_episode(self, max_steps=2000, manage_every=64):
        """Step until every env has finished at least one more episode (checked every
        `manage_every` steps, where the controller also runs), or `max_steps` steps."""
        base = int(self.stats[0])
        steps = 0
        B = self.cfg.batch
        while steps < max_steps:
            self.step()
            steps += 1
            if steps % manage_every == 0:
                self.manage()
                if int(self.stats[0]) - base >= B:
                    break
        c = self.counters()
        return dict(steps=steps, finished=c["episodes"] - base, goals=c["goals"], mean_return=c["mean_return"],
                    n_active=self.n_active, env_steps=steps * B)
