"""world_size-2 test of the sharded path on CPU (gloo): each rank owns a contiguous env slice of the
oracle agent, the product's allreduce_deltas() sums dW / cnt over ranks, every rank applies the same
delta.  The result must equal the unsharded run."""
import os
import socket

import numpy as np
import pytest

import oracle

B, K, T = 32, 2, 3


def _cfg(batch, env_offset):
    return oracle.AgentConfig(map="easy", batch=batch, order=2, max_options=K, seed=11, env_offset=env_offset,
                              sync_interval=10 ** 9, epsilon=0.3, alpha=0.01)


def _run_window(ag, n):
    for _ in range(n):
        ag.step()


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from skill_chaining_with_graphs_b200.sync import allreduce_deltas, env_slice, world_size
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert world_size() == world
    lo, hi = env_slice(B, rank, world)
    ag = oracle.SkillChainAgent(_cfg(hi - lo, lo))
    rng = np.random.default_rng(0)
    W = (rng.standard_normal(ag.options.W.shape) * 0.1).astype(np.float32)
    ag.options.W[:] = W
    for _ in range(2):                                        # two windows of T steps
        _run_window(ag, T)
        dW = torch.from_numpy(ag.options.dW.astype(np.float32))
        cnt = torch.from_numpy(ag.options.cnt.astype(np.int32))
        allreduce_deltas(dW, cnt)
        ag.options.apply(dW.numpy().astype(np.float64), cnt.numpy().astype(np.int64))
    q.put((rank, lo, hi, ag.env.state.copy(), ag.options.W.copy(), ag.action.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_the_unsharded_run():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = oracle.SkillChainAgent(_cfg(B, 0))
    rng = np.random.default_rng(0)
    full.options.W[:] = (rng.standard_normal(full.options.W.shape) * 0.1).astype(np.float32)
    for _ in range(2):
        _run_window(full, T)
        full.options.apply()
    state = np.concatenate([r[3] for r in res])
    action = np.concatenate([r[5] for r in res])
    assert res[0][1:3] == (0, 16) and res[1][1:3] == (16, 32)
    assert np.array_equal(res[0][4], res[1][4])                               # replicas bit-identical
    assert np.abs(res[0][4] - full.options.W).max() < 1e-6 * max(1.0, np.abs(full.options.W).max())
    same = (state == full.env.state).all(axis=1)
    assert same.mean() > 0.9                                                  # only argmax near-ties may differ
    assert (action == full.action)[same].mean() > 0.9


def _fit_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from skill_chaining_with_graphs_b200.sync import fit_union
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, y = _examples(rank)
    o = oracle.OptionSet(2, 1, 1)

    def grad_sum(theta):                     # this rank's gradient SUM, from the oracle's mean gradient
        if len(X) == 0:
            return torch.zeros(6)
        o.theta[1] = theta.numpy()
        return torch.from_numpy(o.clf_grad(1, X, y) * np.float32(len(X)))

    theta = fit_union(grad_sum, len(X), torch.zeros(6), 60, 2.0)
    q.put((rank, theta.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _examples(rank):
    """Rank-dependent example sets of different sizes (rank 2 of a 3-rank world has none)."""
    n = (300, 130, 0)[rank]
    rng = np.random.default_rng(50 + rank)
    X = rng.random((n, 2)).astype(np.float32)
    y = ((X[:, 0] - 0.6) ** 2 + (X[:, 1] - 0.4) ** 2 < 0.1).astype(np.uint8)
    return X, y


@pytest.mark.parametrize("world", [2, 3])
def test_union_fit_over_ranks_equals_the_oracle_fit_on_the_concatenated_examples(world):
    """The multi-rank classifier fit of the controller (sync.fit_union: per-step gradient sums and counts added over
    ranks) gives every rank the same theta, equal to oracle fit_initiation on the union of the ranks' examples - not
    the average of per-rank fits."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_fit_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=120) for _ in ps), key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(1, world):
        assert np.array_equal(res[r][1], res[0][1])                         # identical on every rank
    X = np.concatenate([_examples(r)[0] for r in range(world)])
    y = np.concatenate([_examples(r)[1] for r in range(world)])
    ora = oracle.OptionSet(2, 1, 1)
    ora.fit_initiation(1, X, y, steps=60, lr=2.0)
    from oracle.compare import assert_close
    assert_close(res[0][1], ora.theta[1], what="theta of the union fit")
    avg = np.zeros(6)
    for r in range(world):                                                   # what averaging per-rank fits would give
        Xr, yr = _examples(r)
        if len(Xr):
            o = oracle.OptionSet(2, 1, 1)
            avg += o.fit_initiation(1, Xr, yr, steps=60, lr=2.0) / world
    assert np.abs(avg - ora.theta[1]).max() > 1e-2                           # ... is a different (wrong) answer


def test_single_rank_is_a_no_op():
    import torch
    from skill_chaining_with_graphs_b200.sync import allreduce_deltas, world_size
    dW, cnt = torch.ones(3), torch.ones(2, dtype=torch.int32)
    assert world_size() == 1
    a, b = allreduce_deltas(dW, cnt)
    assert a is dW and b is cnt and float(dW.sum()) == 3.0
