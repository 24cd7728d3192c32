"""Philox4x32-10 counter-based RNG (Salmon et al., "Parallel random numbers: as easy
as 1, 2, 3", SC'11) in NumPy uint32/uint64.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Used so that epsilon-greedy draws depend only on (seed, global env id, step, stream) and
not on how the env batch is sharded over GPUs.  The reference
ships no RNG (/root/reference/README.md:1-2 is the whole repository); the algorithm is
the published one and is pinned by its published known-answer vectors in
tests/test_oracle_philox.py.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32(ctr, key, rounds=10):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 (broadcastable) -> (..., 4) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    c0, c1, c2, c3 = (ctr[..., i].astype(np.uint64) for i in range(4))
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> _S32, p0 & _MASK
        hi1, lo1 = p1 >> _S32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def uniform01(u32):
    """uint32 -> float32 in [0, 1): top 24 bits times 2^-24 (exact in fp32)."""
    return (np.asarray(u32, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


# Stream ids: which decision a draw is for.  Counter = (env id, step, stream, 0).
STREAM_ACTION = 0
STREAM_RESET = 1
# (2 = STREAM_RESELECT, oracle/agent.py: first action under a newly selected option)
STREAM_TOP = 3         # the top-level learner's choice of the next option


def draws(seed, env_ids, step, stream):
    """Four uint32 per env for (seed, env id, step, stream)."""
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    ctr = np.zeros(env_ids.shape + (4,), dtype=np.uint32)
    ctr[..., 0] = env_ids
    ctr[..., 1] = np.uint32(step & 0xFFFFFFFF)
    ctr[..., 2] = np.uint32(stream)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32(ctr, key)
