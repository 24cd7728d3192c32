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
