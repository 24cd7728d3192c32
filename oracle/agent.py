"""Batched skill-chaining agent, CPU oracle.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates skill chaining (the paper named in /root/reference/README.md:2; the reference has no
code) for a batch of envs in lock-step, as BASELINE.json configs[1..4] need it:
"full skill chaining (up to 4 options + logistic initiation classifiers)" and the
"option-graph (multiple chains merging at shared subgoal initiation sets)" variant.
SURVEY.md appendix A.5 / A.6.

Option slots 0..K-1.  Slots 0..n_active-1 are ACTIVE (trained classifier, executable only inside
their initiation set); slot g = n_active is GESTATING (initiation set = everywhere, it is the
learner that reaches for the newest target event); later slots are unused.  Slot K-1 is never
promoted, so a gestating option always exists as the fallback.
  parents[k]: bit j set -> the initiation set of option j is a target of option k;
              bit 31 (GOAL_BIT) -> the task goal is a target of option k.
  chain mode: parents[g] = {g-1} (GOAL for g = 0).  graph mode: parents[g] = all active | GOAL.

One agent step t, for every env (this is the definition the fused GPU pipeline mirrors):
  1. s2, r_env, env_done = env.step(a)
  2. I_k(s2) = active[k] and sigmoid(theta_k . psi(s2)) >= 0.5, for all k
  3. hit   = (GOAL in parents[o] and env_done) or any_j(j in parents[o] and I_j(s2))
     t_opt += 1; ep_steps += 1
     term  = env_done or hit or t_opt >= option_timeout or (active[o] and not I_o(s2))
             or ep_steps >= max_episode_steps
     r     = r_env + (option_bonus if hit and not env_done)
  4. a2 = eps-greedy(Q_o(s2, .); Philox(seed; env, t, STREAM_ACTION))
  5. Sarsa(lambda) update of option o with (s, a, r, s2, a2, done=term)
  6. where term: append (option start position, label = hit and t_opt <= init_horizon) to option o's example ring;
     n_success[o] += hit, n_fail[o] += not hit
  7. where env_done or episode timed out: env reset to a start state, ep_steps = 0
  8. where term: o' = first active k with I_k(s_next), else g;  t_opt = 0; start position = s_next
                 a' = eps-greedy(Q_o'(s_next, .); Philox(seed; env, t, STREAM_RESELECT))
     else:       o' = o, a' = a2
  9. every sync_interval steps: OptionSet.apply()
Top-level SMDP learner (AgentConfig.top_level=True; SURVEY.md section 8 f-3 - the paper's agent learns which option to
run): Q_top(s, j) over the option slots j, linear in the same Fourier features, stored in extra slots of the OptionSet's
weight tables (oracle/option.py).  The admissible set at s is A(s) = {active k with I_k(s)} + {the gestating slot g}
(g's initiation set is everywhere: it is the flat learner through which the primitive actions stay reachable).
  * per env, since its option o started at s0:  R += disc * r_env;  disc *= gamma   (task reward, no option bonus)
  * at step 8 (option terminated), BEFORE re-selection, with s2 the state the step reached (before any reset):
        target    = R + (0 if the episode ended or timed out else disc * max_{j in A(s2)} Q_top(s2, j))
        delta_top = target - Q_top(s0, o)
        dW[top slot of o] += delta_top * phi(s0);   cnt[top slots] += 1        (applied with alpha_top, mean over events)
  * re-selection: o' = eps_top-greedy over A(s_next) of Q_top(s_next, .): with (u0, u1) = Philox(seed; env, t, STREAM_TOP),
        explore iff u0 < epsilon_top: the min(int(u1 * |A|), |A| - 1)-th admissible slot in ascending order; otherwise the
        first admissible slot with the maximal value.  With zero top-level weights this is the "first active" rule above.

manage() is the low-rate controller: when the gestating option has collected enough successes its
classifier is fit on its examples, it becomes active and the next slot starts gestating.
"""
from dataclasses import dataclass

import numpy as np

from .option import OptionSet, epsilon_greedy
from .philox import STREAM_ACTION, STREAM_TOP, draws, uniform01
from .pinball import PinballEnv, PinballMap

f32 = np.float32
GOAL_BIT = np.uint32(1 << 31)
STREAM_RESELECT = 2


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
    init_horizon: int = 1 << 30   # an example is positive iff the option hit a target within this many steps of its start
    merge_overlap: float = 0.0    # graph mode: an older option's initiation set becomes a target of the next option iff at
                                  # least this fraction of the promoted option's positive examples lies inside it (0: all)
    graph: bool = False
    windowed: bool = False      # Sarsa(lambda) in the forward-view window form (OptionSet.flush) instead of the dense sweep
    top_level: bool = False     # option choice by the learned SMDP value function Q_top instead of "first active"
    alpha_top: float = 1e-3
    epsilon_top: float = 0.05


class SkillChainAgent:
    def __init__(self, cfg, pmap=None):
        self.cfg = cfg
        self.map = pmap if pmap is not None else PinballMap.from_name(cfg.map)
        B, K = cfg.batch, cfg.max_options
        self.env = PinballEnv(self.map, B, seed=cfg.seed, env_offset=cfg.env_offset)
        self.options = OptionSet(K, cfg.order, B, cfg.gamma, cfg.lam, cfg.alpha, cfg.epsilon,
                                 cfg.seed, cfg.env_offset, windowed=cfg.windowed,
                                 top_slots=(-(-K // 5) if cfg.top_level else 0), alpha_top=cfg.alpha_top)
        self.active = np.zeros(K, dtype=bool)
        self.parents = np.zeros(K, dtype=np.uint32)
        self.parents[0] = GOAL_BIT
        self.n_active = 0
        self.t = 0
        self.option = np.zeros(B, dtype=np.int32)
        self.t_opt = np.zeros(B, dtype=np.int32)
        self.ep_steps = np.zeros(B, dtype=np.int32)
        self.start_xy = self.env.state[:, :2].copy()
        self.opt_s0 = self.env.state.copy()              # top-level learner: state where the current option started,
        self.opt_R = np.zeros(B, dtype=np.float32)       # discounted task reward since then,
        self.opt_disc = np.ones(B, dtype=np.float32)     # and gamma^(steps since then)
        self.last_delta_top = np.zeros(B, dtype=np.float32)
        self.ex_xy = np.zeros((K, cfg.example_capacity, 2), dtype=np.float32)
        self.ex_label = np.zeros((K, cfg.example_capacity), dtype=np.uint8)
        self.ex_count = np.zeros(K, dtype=np.int64)
        self.n_success = np.zeros(K, dtype=np.int64)
        self.n_fail = np.zeros(K, dtype=np.int64)
        self.episodes = np.zeros(B, dtype=np.int64)
        self.goals = np.zeros(B, dtype=np.int64)
        self.ep_return = np.zeros(B, dtype=np.float64)
        self.last_return = np.full(B, np.nan)
        self.last_delta = np.zeros(B, dtype=np.float32)
        self.action = self.options.act(self.env.state, self.option, step=0xFFFFFFFF, stream=STREAM_RESELECT)

    # -- helpers ------------------------------------------------------------------------------
    def initiation_bits(self, state):
        """uint32 (B,): bit k set iff option k is active and its classifier accepts the state."""
        I = self.options.initiation(state) & self.active[None, :]
        w = (np.uint32(1) << np.arange(self.options.K, dtype=np.uint32))[None, :]
        return (I * w).sum(axis=1).astype(np.uint32)

    def choose_option_top(self, bits, state, t):
        """eps_top-greedy over the admissible slots of Q_top(state, .) (module docstring).  -> (choice, Q_top rows)"""
        K = self.options.K
        g = min(self.n_active, K - 1)
        adm = ((bits[:, None] >> np.arange(K, dtype=np.uint32)[None, :]) & np.uint32(1)).astype(bool)
        adm[:, g] = True
        Qt = self.options.q_top(state)
        masked = np.where(adm, Qt, -np.inf)
        greedy = np.argmax(masked, axis=1).astype(np.int32)                # first admissible maximum
        r = draws(self.cfg.seed, self.options.env_ids, t, STREAM_TOP)
        u0, u1 = uniform01(r[:, 0]), uniform01(r[:, 1])
        n_adm = adm.sum(axis=1).astype(np.int32)
        pick = np.minimum((u1 * n_adm.astype(np.float32)).astype(np.int32), n_adm - 1)
        rank = np.cumsum(adm, axis=1) - 1                                  # rank of each admissible slot
        rand = np.argmax(adm & (rank == pick[:, None]), axis=1).astype(np.int32)
        return np.where(u0 < f32(self.cfg.epsilon_top), rand, greedy).astype(np.int32), Qt, adm

    def choose_option(self, bits):
        """First active option whose initiation set holds, else the gestating slot."""
        K = self.options.K
        g = min(self.n_active, K - 1)
        out = np.full(bits.shape, g, dtype=np.int32)
        for k in range(K - 1, -1, -1):
            out = np.where((bits >> np.uint32(k)) & np.uint32(1) != 0, k, out)
        return out.astype(np.int32)

    # -- one lock-step agent step -------------------------------------------------------------
    def step(self, follow=None):
        """One lock-step agent step.  `follow` (tests only) = dict(action=int32 (B,), option=int32 (B,)): the action and
        option every env holds AFTER this step in another implementation's run of the same step.  The oracle still
        computes its own choices (returned as own_a2 / own_action / own_option, with the Q rows they were taken
        from) but continues with the followed ones, so two implementations can be compared over many steps without
        argmax near-ties (which legitimately differ between fp32 summation orders) sending the trajectories apart."""
        cfg, env, opts = self.cfg, self.env, self.options
        t = self.t
        B = cfg.batch
        idx = np.arange(B)
        s = env.state.copy()
        a, o = self.action, self.option
        s2, r_env, env_done, _ = env.step(a)
        bits = self.initiation_bits(s2)
        pm = self.parents[o]
        hit = (((pm & GOAL_BIT) != 0) & env_done) | ((bits & pm & ~GOAL_BIT) != 0)
        self.t_opt += 1
        self.ep_steps += 1
        in_own = ((bits >> o.astype(np.uint32)) & np.uint32(1)) != 0
        left = self.active[o] & ~in_own
        ep_timeout = (self.ep_steps >= cfg.max_episode_steps) & ~env_done
        term = env_done | hit | (self.t_opt >= cfg.option_timeout) | left | ep_timeout
        r = (r_env + np.where(hit & ~env_done, f32(cfg.option_bonus), f32(0.0))).astype(np.float32)
        Q2 = opts.q(s2, o)
        a2 = epsilon_greedy(Q2, opts.epsilon, opts.seed, opts.env_ids, t, STREAM_ACTION)
        own_a2 = a2
        if follow is not None:      # where the option terminated a2 does not enter the update (done = term)
            a2 = np.where(term, a2, np.asarray(follow["action"], dtype=np.int32)).astype(np.int32)
        delta = opts.update(s, a, r, s2, a2, term, o)
        self.last_delta = delta
        self.ep_return += r_env
        if cfg.top_level:       # SMDP return of the running option (fp32, this operation order)
            self.opt_R = (self.opt_R + (self.opt_disc * r_env).astype(np.float32)).astype(np.float32)
            self.opt_disc = (self.opt_disc * f32(cfg.gamma)).astype(np.float32)
            if term.any():
                K = opts.K
                g = min(self.n_active, K - 1)
                adm2 = ((bits[:, None] >> np.arange(K, dtype=np.uint32)[None, :]) & np.uint32(1)).astype(bool)
                adm2[:, g] = True
                m2 = np.where(adm2, opts.q_top(s2), -np.inf).max(axis=1).astype(np.float32)
                ended = env_done | ep_timeout
                boot = np.where(ended, f32(0.0), (self.opt_disc * m2).astype(np.float32)).astype(np.float32)
                q0 = opts.q_top(self.opt_s0)[idx, o]
                d_top = ((self.opt_R + boot).astype(np.float32) - q0).astype(np.float32)
                self.last_delta_top = np.where(term, d_top, f32(0.0)).astype(np.float32)
                opts.top_update(self.opt_s0, o, d_top, term)
        for b in np.nonzero(term)[0]:
            k = o[b]
            slot = self.ex_count[k] % cfg.example_capacity
            self.ex_xy[k, slot] = self.start_xy[b]
            self.ex_label[k, slot] = 1 if (hit[b] and self.t_opt[b] <= cfg.init_horizon) else 0
            self.ex_count[k] += 1
            if hit[b]:
                self.n_success[k] += 1
            else:
                self.n_fail[k] += 1
        reset = env_done | ep_timeout
        if reset.any():
            env.reset(mask=reset, step=t)
            self.episodes[reset] += 1
            self.goals[env_done] += 1
            self.last_return[reset] = self.ep_return[reset]
            self.ep_return[reset] = 0
            self.ep_steps[reset] = 0
        s_next = env.state
        a_next, o_next = a2.copy(), o.copy()
        own_action, own_o, Qsel = own_a2.copy(), o.copy(), Q2
        if term.any():
            bits_next = self.initiation_bits(s_next)
            if cfg.top_level:
                o_sel, Qtop, adm = self.choose_option_top(bits_next, s_next, t)
                self._last_Qtop = np.where(adm, Qtop, -np.inf)
            else:
                o_sel = self.choose_option(bits_next)
            o_next = np.where(term, o_sel, o).astype(np.int32)
            if follow is not None:  # evaluate the first action under the followed option
                own_o = o_next
                o_next = np.where(term, np.asarray(follow["option"], dtype=np.int32), o).astype(np.int32)
            Qn = opts.q(s_next, o_next)
            a_sel = epsilon_greedy(Qn, opts.epsilon, opts.seed, opts.env_ids, t, STREAM_RESELECT)
            a_next = np.where(term, a_sel, a2).astype(np.int32)
            own_action = np.where(term, a_sel, own_a2).astype(np.int32)
            Qsel = np.where(term[:, None], Qn, Q2)
            self.t_opt[term] = 0
            self.start_xy[term] = s_next[term, :2]
            self.opt_s0[term] = s_next[term]
            self.opt_R[term] = 0
            self.opt_disc[term] = 1
        extra = {}
        if follow is not None:
            # own_action / own_option: what the oracle would have chosen itself; Qsel: the Q row each own_action was
            # the eps-greedy choice from (under the followed option where the option was re-selected)
            extra = dict(own_action=own_action, own_option=own_o, Qsel=Qsel)
            if cfg.top_level:       # the top-level values the option choice was made from (-inf: not admissible)
                extra["Qtop"] = self._last_Qtop if term.any() else None
                extra["delta_top"] = self.last_delta_top
            a_next = np.asarray(follow["action"], dtype=np.int32).copy()
        self.option, self.action = o_next, a_next
        opts.tick()
        self.t += 1
        if self.t % cfg.sync_interval == 0:
            opts.apply()
        return dict(state=s_next.copy(), reward=r, env_done=env_done, term=term, hit=hit, delta=delta,
                    option=o_next.copy(), action=a_next.copy(), **extra)

    # -- low-rate controller ------------------------------------------------------------------
    def examples(self, k):
        n = int(min(self.ex_count[k], self.cfg.example_capacity))
        return self.ex_xy[k, :n].copy(), self.ex_label[k, :n].copy()

    def manage(self):
        """Promote the gestating option once it has enough successes.  Returns True on promotion."""
        cfg, K = self.cfg, self.options.K
        g = self.n_active
        if g >= K - 1 or self.n_success[g] < cfg.gestation_successes:
            return False
        X, y = self.examples(g)
        self.options.theta[g] = 0
        if len(X):
            self.options.fit_initiation(g, X, y, cfg.clf_steps, cfg.clf_lr)
        self.active[g] = True
        self.n_active += 1
        n = self.n_active
        if cfg.graph:
            self.parents[n] = self.merged_parents(g, X, y)
        else:
            self.parents[n] = np.uint32(1 << (n - 1))
        return True

    def merged_parents(self, g, X, y):
        """Option graph, merge detection ("a new option may adopt an existing option's initiation
        set as an additional target when their chains meet").  The next option's targets are the initiation set of the
        option just promoted (g), every older active option j whose initiation set MEETS g's - at least
        cfg.merge_overlap of g's positive examples (start positions from which g reached its target) lie inside I_j -
        and the goal if it lies inside I_g or merge_overlap is 0.  With merge_overlap = 0 every older option and the
        goal qualify: the plain graph mode."""
        cfg = self.cfg
        pm = np.uint32(1 << g)
        pos = np.asarray(X, dtype=np.float32).reshape(-1, 2)[np.asarray(y).reshape(-1) != 0]
        if len(pos):
            st = np.concatenate([pos, np.zeros_like(pos)], axis=1)
            inside = self.options.initiation_logit(st) >= f32(0.0)                   # (n_pos, K)
        for j in range(g):
            frac = f32(inside[:, j].sum()) / f32(len(pos)) if len(pos) else f32(0.0)
            if frac >= f32(cfg.merge_overlap):
                pm |= np.uint32(1 << j)
        tx, ty, _ = self.map.target
        goal_in = self.options.initiation_logit(np.array([[tx, ty, 0, 0]], dtype=np.float32))[0, g] >= f32(0.0)
        if cfg.merge_overlap <= 0.0 or goal_in:
            pm |= GOAL_BIT
        return pm

    def run_episode(self, max_steps=2000, manage_every=64):
        """Step until every env has finished at least one more episode, or `max_steps` steps."""
        base = self.episodes.copy()
        steps = 0
        while steps < max_steps and (self.episodes == base).any():
            self.step()
            steps += 1
            if steps % manage_every == 0:
                self.manage()
        fin = self.episodes > base
        return dict(steps=steps, finished=int(fin.sum()), goals=int(self.goals.sum()),
                    mean_return=float(np.nanmean(self.last_return[fin])) if fin.any() else float("nan"),
                    n_active=int(self.n_active), env_steps=steps * self.cfg.batch)
