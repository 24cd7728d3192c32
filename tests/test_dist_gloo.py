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


def test_single_rank_is_a_no_op():
    import torch
    from skill_chaining_with_graphs_b200.sync import allreduce_deltas, world_size
    dW, cnt = torch.ones(3), torch.ones(2, dtype=torch.int32)
    assert world_size() == 1
    a, b = allreduce_deltas(dW, cnt)
    assert a is dW and b is cnt and float(dW.sum()) == 3.0
