"""Order-n Fourier basis over the 4-D Pinball state, CPU oracle.  TEST INFRASTRUCTURE.

phi_i(s) = cos(pi * c_i . s_hat), c_i in {0..n}^4, F = (n+1)^4 (BASELINE.json north_star:
"order-n Fourier basis", "phi = cos(pi C s)").  The reference has no
code to follow (/root/reference/README.md:1-2).

Pinned conventions (tests/test_oracle_fourier.py):
  * multi-index order is lexicographic with c0 slowest:
        f = ((c0*(n+1) + c1)*(n+1) + c2)*(n+1) + c3
  * s_hat = (x, y, (vx + 2)/4, (vy + 2)/4): positions are already in [0, 1]; a velocity
    component is clipped to [-1, 1] on thrust but a reflection can rotate speed sqrt(2) onto
    one axis, so velocities are normalised over [-2, 2]
  * per-feature step-size scale 1/||c_i||_2, and 1 for c = 0 (Konidaris et al., Fourier basis)
"""
import numpy as np

f32 = np.float32
PI32 = f32(np.pi)


class FourierBasis:
    def __init__(self, order, n_dims=4):
        if n_dims != 4:
            raise ValueError("the Pinball state is 4-D")
        self.order = int(order)
        n1 = self.order + 1
        self.n_features = n1 ** 4
        f = np.arange(self.n_features)
        self.C = np.stack([(f // n1 ** 3) % n1, (f // n1 ** 2) % n1, (f // n1) % n1, f % n1], axis=1).astype(np.int32)
        norm = np.sqrt((self.C.astype(np.float64) ** 2).sum(axis=1))
        norm[norm == 0] = 1.0
        self.alpha_scale = (1.0 / norm).astype(np.float32)

    @staticmethod
    def normalise(state):
        s = np.asarray(state, dtype=np.float32).reshape(-1, 4)
        out = np.empty_like(s)
        out[:, 0] = s[:, 0]
        out[:, 1] = s[:, 1]
        out[:, 2] = (s[:, 2] + f32(2.0)) * f32(0.25)
        out[:, 3] = (s[:, 3] + f32(2.0)) * f32(0.25)
        return out

    def features(self, state):
        """state (B, 4) -> phi float32 (B, F).  The angle is formed in fp64 and cos evaluated in
        fp64, then rounded to fp32: the oracle is the correctly rounded value of the definition,
        so any fp32 evaluation scheme on the GPU (direct cos or phasor products) is judged
        against the same target."""
        sh = self.normalise(state).astype(np.float64)
        ang = sh @ self.C.T.astype(np.float64)
        return np.cos(np.pi * ang).astype(np.float32)
