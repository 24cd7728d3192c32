"""Cross-rank weight-delta exchange (S5): every rank owns an env slice and a full replica of the option
weights; at each sync the window's accumulated dW (K, A, F) and update counts cnt (K,) are summed over
ranks so that every rank applies the identical normalised delta (oracle/option.py OptionSet.apply's
"dW and cnt are summed over ranks before apply()").  Works on CUDA tensors over NCCL (the product
path) and on CPU tensors over gloo (the world_size-2 tests of the host logic)."""


def world_size(group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


def env_slice(global_envs, rank, world):
    """Contiguous env-id slice [lo, hi) owned by `rank` (the first `global_envs % world` ranks get one
    extra env).  env ids are global, so Philox draws do not depend on the sharding."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(int(global_envs), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_deltas(dW, cnt, group=None):
    """In-place sum of dW (float32) and cnt (int32) over the ranks of `group`; no-op for one rank.
    Both collectives are issued before either is waited on."""
    import torch.distributed as dist
    if world_size(group) == 1:
        return dW, cnt
    h1 = dist.all_reduce(dW, op=dist.ReduceOp.SUM, group=group, async_op=True)
    h2 = dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group, async_op=True)
    h1.wait()
    h2.wait()
    return dW, cnt


def allreduce_scalar_sum(t, group=None):
    import torch.distributed as dist
    if world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def fit_union(grad_sum, n_local, theta0, steps, lr, group=None):
    """Batch gradient descent of a logistic initiation classifier on the UNION of every rank's examples, without moving
    the examples: at each step every rank computes its own gradient SUM `grad_sum(theta)` (6 values) over its `n_local`
    examples, the sums and counts are added over ranks (one 7-value all-reduce), and every rank takes the same step
        theta <- theta - lr * (sum over ranks of grad_sum) / (sum over ranks of n_local)
    so all ranks end with the same theta, equal to one fit on the concatenated example sets (oracle/option.py
    fit_initiation on the union).  Works on CUDA tensors over NCCL and CPU tensors over gloo."""
    import torch
    import torch.distributed as dist
    theta = theta0.clone()
    multi = world_size(group) > 1
    for _ in range(int(steps)):
        v = torch.cat([grad_sum(theta).to(torch.float32).reshape(6),
                       torch.tensor([float(n_local)], dtype=torch.float32, device=theta.device)])
        if multi:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        n = float(v[6])
        if n > 0:
            theta = theta - float(lr) * (v[:6] / n)
    return theta
