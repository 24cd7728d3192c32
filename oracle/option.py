"""Per-option linear Sarsa(lambda) and logistic initiation classifiers, CPU oracle.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, for a batch of envs, the option learner of the paper named in
/root/reference/README.md:2 (the reference itself has no code): BASELINE.json north_star
"per-option Sarsa(lambda) over an order-n Fourier basis, and per-option logistic-regression
initiation classifiers"; SURVEY.md appendix A.3 / A.4.

Batched semantics (definition):
  * every env b carries its own eligibility trace e_b (A x F) for the option it is executing
  * Q_o(s, a) = W[o, a] . phi(s)
  * delta_b = r_b + gamma (1 - done_b) Q_o(s2_b, a2_b) - Q_o(s_b, a_b)
  * e_b <- gamma*lambda e_b ;  e_b[a_b] += phi(s_b)      (accumulating traces)
  * dW[o] += sum over envs on o of delta_b e_b ;  cnt[o] += number of such envs
  * e_b <- 0 where done_b (option or episode ended)
  * weights are frozen between syncs; OptionSet.apply() does, for every option with cnt > 0,
        W[o] += alpha * alpha_scale (.) dW[o] * (window_steps / cnt[o])
    then zeroes dW, cnt.  With B = 1 and a sync every step this is classical Sarsa(lambda).
  * multi-GPU: dW and cnt are summed over ranks before apply() (sharding-invariant).

Windowed (forward-view) form, `OptionSet(..., windowed=True)`: because the
weights are frozen between syncs, the TD errors of a window do not depend on the traces, so the
per-step sweep "decay e, add phi, dW += delta e" can be replaced by recording (s, a, delta, done,
option) per step and, at the end of the window, for each env over its recorded steps i = 0..n-1
      G_i = delta_i + (0 if done_i else gamma*lambda * G_{i+1}),  G_n = 0
      dW[o_i][a_i] += G_i phi(s_i) ;   dW[o_0] += gamma*lambda * G_0 * e_start
      e_end = (0 if any done else (gamma*lambda)^n) e_start + sum_i c_i phi(s_i) (x) a_i,
      c_i = (gamma*lambda)^(n-1-i) if no done at any step >= i else 0
  which is algebraically the same sum (tests/test_oracle_option.py checks the two forms equal).
  It assumes what the agent guarantees: an env changes option only after a step with done = True.

Top-level value function (OptionSet(..., top_slots=n), used by oracle/agent.py with top_level=True; SURVEY.md
section 8 f-3): the SMDP learner over options shares the basis and the weight tables.  Q_top(s, j), j = 0..K-1, is row
j % 5 of the extra slot K + j // 5 of W (n = ceil(K / 5) slots), so the same q() evaluates it; its window delta lives
in dW[K:], and apply() steps those slots with alpha_top and the MEAN over the window's termination events
(cnt[K:] all hold the event count):  W[k] += alpha_top * alpha_scale (.) dW[k] / cnt[k]  for k >= K.

Initiation classifier of option k: p = sigmoid(theta_k . psi(x, y)), psi = (1, x, y, x^2, xy, y^2);
I_k(s) = (p >= 0.5), decided on the fp32 logit z = theta_k . psi >= 0 (initiation_logit: fixed operation order, one
rounding per operation, bit-reproducible).  fit: theta -= lr * mean_i (p_i - y_i) psi_i, a fixed number of steps.
"""
import numpy as np

from .fourier import FourierBasis
from .philox import draws, uniform01, STREAM_ACTION

f32 = np.float32
N_ACTIONS = 5
N_PSI = 6


def logistic_features(state):
    s = np.asarray(state, dtype=np.float32)
    s = s.reshape(-1, s.shape[-1])
    x, y = s[:, 0], s[:, 1]
    return np.stack([np.ones_like(x), x, y, x * x, x * y, y * y], axis=1).astype(np.float32)


def sigmoid(z):
    z = np.asarray(z, dtype=np.float64)
    return (1.0 / (1.0 + np.exp(-z))).astype(np.float32)


def epsilon_greedy(Q, epsilon, seed, env_ids, step, stream=STREAM_ACTION):
    """Q (B, A) -> actions int32 (B,).  Draw (u0, u1) = Philox(seed; env id, step, stream):
    explore iff u0 < epsilon, then a = min(int(u1 * A), A-1); else the first maximal action."""
    r = draws(seed, env_ids, step, stream)
    u0 = uniform01(r[:, 0])
    u1 = uniform01(r[:, 1])
    greedy = np.argmax(Q, axis=1).astype(np.int32)      # first maximum: ties -> lowest index
    rand = np.minimum((u1 * f32(N_ACTIONS)).astype(np.int32), N_ACTIONS - 1)
    return np.where(u0 < f32(epsilon), rand, greedy).astype(np.int32)


class OptionSet:
    """K options sharing one env batch: weights W (K, A, F), classifiers theta (K, 6),
    per-env traces (B, A, F), window accumulators dW (K, A, F) / cnt (K,)."""

    def __init__(self, n_options, order, batch, gamma=0.99, lam=0.9, alpha=1e-3, epsilon=0.05,
                 seed=0, env_offset=0, windowed=False, top_slots=0, alpha_top=1e-3):
        self.K = int(n_options)
        self.top_slots = int(top_slots)
        self.K_all = self.K + self.top_slots
        self.alpha_top = f32(alpha_top)
        self.windowed = bool(windowed)
        self._win = []
        self.basis = FourierBasis(order)
        self.F = self.basis.n_features
        self.B = int(batch)
        self.gamma = f32(gamma)
        self.lam = f32(lam)
        self.alpha = f32(alpha)
        self.epsilon = f32(epsilon)
        self.seed = int(seed)
        self.env_ids = np.arange(self.B, dtype=np.uint32) + np.uint32(env_offset)
        self.W = np.zeros((self.K_all, N_ACTIONS, self.F), dtype=np.float32)
        self.theta = np.zeros((self.K, N_PSI), dtype=np.float32)
        self.trace = np.zeros((self.B, N_ACTIONS, self.F), dtype=np.float32)
        self.dW = np.zeros((self.K_all, N_ACTIONS, self.F), dtype=np.float64)
        self.cnt = np.zeros(self.K_all, dtype=np.int64)
        self.window_steps = 0

    # -- value function -----------------------------------------------------------------------
    def q(self, state, option_ids, phi=None):
        """Q_o(s, .) for each env's option: (B, A) float32 (fp64 accumulation, rounded once)."""
        if phi is None:
            phi = self.basis.features(state)
        option_ids = np.asarray(option_ids)
        Q = np.zeros((phi.shape[0], N_ACTIONS), dtype=np.float32)
        for k in np.unique(option_ids):
            mk = option_ids == k
            Q[mk] = (phi[mk].astype(np.float64) @ self.W[k].T.astype(np.float64)).astype(np.float32)
        return Q

    def q_top(self, state, phi=None):
        """Q_top(s, j) for every option slot j: (B, K) float32 (rows of the top-level slots of W)."""
        if phi is None:
            phi = self.basis.features(state)
        Q = np.zeros((phi.shape[0], self.K), dtype=np.float32)
        for j in range(self.K):
            w = self.W[self.K + j // N_ACTIONS, j % N_ACTIONS].astype(np.float64)
            Q[:, j] = (phi.astype(np.float64) @ w).astype(np.float32)
        return Q

    def top_update(self, s0, option_ids, delta_top, mask):
        """SMDP update of the top-level learner for the envs in `mask` (their option just terminated): dW of row
        `option` of the top-level slots += delta_top * phi(s0); every top-level slot's cnt += number of events."""
        sel = np.nonzero(mask)[0]
        if len(sel) == 0 or self.top_slots == 0:
            return
        phi = self.basis.features(np.asarray(s0, dtype=np.float32)[sel]).astype(np.float64)
        d = np.asarray(delta_top, dtype=np.float64)[sel]
        o = np.asarray(option_ids)[sel]
        for j in np.unique(o):
            mj = o == j
            self.dW[self.K + j // N_ACTIONS, j % N_ACTIONS] += d[mj] @ phi[mj]
        self.cnt[self.K:] += len(sel)

    def act(self, state, option_ids, step, stream=STREAM_ACTION, env_ids=None):
        Q = self.q(state, option_ids)
        ids = self.env_ids if env_ids is None else env_ids
        return epsilon_greedy(Q, self.epsilon, self.seed, ids, step, stream)

    def td_error(self, s, a, r, s2, a2, done, option_ids):
        idx = np.arange(len(a))
        q_sa = self.q(s, option_ids)[idx, a]
        q_s2 = self.q(s2, option_ids)[idx, a2]
        notdone = (~np.asarray(done, dtype=bool)).astype(np.float32)
        return (np.asarray(r, dtype=np.float32) + self.gamma * notdone * q_s2 - q_sa).astype(np.float32)

    # -- Sarsa(lambda) --------------------------------------------------------------------------
    def update(self, s, a, r, s2, a2, done, option_ids, mask=None):
        """One Sarsa(lambda) update for every env (or those in `mask`).  Returns delta (B,)."""
        a = np.asarray(a)
        a2 = np.asarray(a2)
        option_ids = np.asarray(option_ids)
        done = np.asarray(done, dtype=bool)
        B = self.B
        if mask is None:
            mask = np.ones(B, dtype=bool)
        delta = self.td_error(s, a, r, s2, a2, done, option_ids)
        delta = np.where(mask, delta, f32(0.0)).astype(np.float32)
        if self.windowed:
            self._win.append((np.array(s, dtype=np.float32).reshape(B, 4), a.astype(np.int32).copy(), delta.copy(),
                              done.copy(), option_ids.astype(np.int32).copy(), np.asarray(mask, dtype=bool).copy()))
            self.cnt[: self.K] += np.bincount(option_ids[mask], minlength=self.K)[: self.K]
            return delta
        phi = self.basis.features(s)
        gl = f32(self.gamma * self.lam)
        sel = np.nonzero(mask)[0]
        self.trace[sel] *= gl
        self.trace[sel, a[sel]] += phi[sel]
        for k in range(self.K):
            mk = sel[option_ids[sel] == k]
            if len(mk):
                self.dW[k] += np.tensordot(delta[mk].astype(np.float64),
                                           self.trace[mk].astype(np.float64), axes=(0, 0))
                self.cnt[k] += len(mk)
        self.trace[sel[done[sel]]] = 0
        return delta

    def flush(self):
        """Windowed mode: fold the recorded steps into dW and the traces (see module docstring)."""
        if not self._win:
            return
        T, B = len(self._win), self.B
        gl = np.float64(f32(self.gamma * self.lam))
        act = np.stack([w[5] for w in self._win])                  # (T, B)
        dn = np.stack([w[3] for w in self._win]) & act
        dl = np.stack([w[2] for w in self._win]).astype(np.float64)
        G = np.zeros((T + 1, B))
        c = np.zeros((T, B))
        cc = np.ones(B)
        dead = np.zeros(B, dtype=bool)
        nxt = np.zeros(B)
        for t in range(T - 1, -1, -1):
            a_t = act[t]
            g = dl[t] + np.where(dn[t], 0.0, gl * nxt)
            G[t] = np.where(a_t, g, nxt)
            nxt = G[t]
            dead |= dn[t]
            c[t] = np.where(a_t & ~dead, cc, 0.0)
            cc = np.where(a_t, cc * gl, cc)
        e_scale = np.where(dead, 0.0, cc)
        # carry-in: the trace at the window start belongs to the option of the env's first recorded step
        first = np.argmax(act, axis=0)
        has = act.any(axis=0)
        o0 = np.stack([w[4] for w in self._win])[first, np.arange(B)]
        carry = np.where(has, gl * G[0], 0.0)
        e0 = self.trace.astype(np.float64)
        for k in range(self.K):
            mk = has & (o0 == k)
            if mk.any():
                self.dW[k] += np.tensordot(carry[mk], e0[mk], axes=(0, 0))
        e = e0 * e_scale[:, None, None]
        idx = np.arange(B)
        for t in range(T):
            s_t, a_t, _, _, o_t, m_t = self._win[t]
            phi = self.basis.features(s_t).astype(np.float64)
            for k in range(self.K):
                mk = m_t & (o_t == k)
                if not mk.any():
                    continue
                for a in range(N_ACTIONS):
                    mka = mk & (a_t == a)
                    if mka.any():
                        self.dW[k, a] += G[t][mka] @ phi[mka]
            e[idx, a_t] += np.where(m_t, c[t], 0.0)[:, None] * phi
        self.trace[:] = e.astype(np.float32)
        self._win = []

    def tick(self):
        """Call once per env step of the window (after update)."""
        self.window_steps += 1

    def apply(self, dW=None, cnt=None):
        """Fold the window's accumulated delta into W (after any cross-rank sum)."""
        self.flush()
        dW = self.dW if dW is None else dW
        cnt = self.cnt if cnt is None else cnt
        steps = max(self.window_steps, 1)
        for k in range(self.K_all):
            if cnt[k] > 0:
                top = k >= self.K        # top-level slots: mean over the window's termination events, own step size
                scale = f32(f32(1 if top else steps) / f32(cnt[k]))
                alpha = self.alpha_top if top else self.alpha
                step = (alpha * self.basis.alpha_scale)[None, :] * (dW[k].astype(np.float32) * scale)
                self.W[k] += step.astype(np.float32)
        self.dW[:] = 0
        self.cnt[:] = 0
        self.window_steps = 0

    # -- initiation classifiers -----------------------------------------------------------------
    def initiation_prob(self, state):
        psi = logistic_features(state).astype(np.float64)
        return sigmoid(psi @ self.theta.T.astype(np.float64))       # (B, K)

    def initiation_logit(self, state):
        """theta_k . psi(x, y) in fp32 with one rounding per operation, in this order (B, K):
               z = ((((t0 + t1*x) + t2*y) + t3*(x*x)) + t4*(x*y)) + t5*(y*y)
        The initiation DECISION is taken on this value, so that it is reproducible bit for bit (a decision taken on a
        rounded probability would flip for states within an ulp of the boundary, e.g. a start position that lies
        exactly on a classifier's edge)."""
        s = np.asarray(state, dtype=np.float32)
        s = s.reshape(-1, s.shape[-1])
        x, y = s[:, 0:1], s[:, 1:2]
        t = self.theta.astype(np.float32)[None, :, :]
        xx, xy, yy = x * x, x * y, y * y                       # fp32 products
        z = t[:, :, 0] + t[:, :, 1] * x
        z = z + t[:, :, 2] * y
        z = z + t[:, :, 3] * xx
        z = z + t[:, :, 4] * xy
        z = z + t[:, :, 5] * yy
        return z.astype(np.float32)

    def initiation(self, state):
        """I_k(s)  <=>  sigmoid(z) >= 0.5  <=>  z >= 0, with z = initiation_logit(s)."""
        return self.initiation_logit(state) >= f32(0.0)

    def clf_grad(self, k, X, y):
        """mean_i (p_i - y_i) psi_i for option k: float32 (6,)."""
        psi = logistic_features(X).astype(np.float64)
        p = 1.0 / (1.0 + np.exp(-(psi @ self.theta[k].astype(np.float64))))
        g = ((p - np.asarray(y, dtype=np.float64))[:, None] * psi).mean(axis=0)
        return g.astype(np.float32)

    def fit_initiation(self, k, X, y, steps=200, lr=1.0):
        for _ in range(int(steps)):
            self.theta[k] = (self.theta[k] - f32(lr) * self.clf_grad(k, X, y)).astype(np.float32)
        return self.theta[k].copy()


class Option:
    """One option's view of an OptionSet: the (initiation, act, update) triple of the paper."""

    def __init__(self, option_set, k):
        self.set = option_set
        self.k = int(k)

    def _ids(self):
        return np.full(self.set.B, self.k, dtype=np.int32)

    def q(self, state):
        return self.set.q(state, self._ids())

    def initiation(self, state):
        return self.set.initiation(state)[:, self.k]

    def act(self, state, step=0):
        return self.set.act(state, self._ids(), step)

    def update(self, s, a, r, s2, a2, done):
        return self.set.update(s, a, r, s2, a2, done, self._ids())

    def fit_initiation(self, X, y, steps=200, lr=1.0):
        return self.set.fit_initiation(self.k, X, y, steps, lr)
