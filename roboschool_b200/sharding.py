"""Env sharding over GPUs: contiguous env-index blocks, one process per GPU, no step-path collective.

The per-env RNG is keyed by the GLOBAL env id (env_id0 + local index), so results do not depend on the
number of ranks.  The only collective of the path is the reduction of rollout statistics
(8 doubles), done with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""


def shard_range(n_total, rank, world_size):
    """[lo, hi) of the contiguous env block owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def reduce_stats(stats, group=None):
    """All-reduce (sum) of the rollout-statistics vector; identity when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarize(stats):
    """stats[8] as written by rs_world_step_stats -> dict."""
    s = [float(x) for x in stats]
    steps = max(s[6], 1.0)
    return {"mean_reward": s[0] / steps, "episodes": s[1], "mean_episode_len": s[2] / max(s[1], 1.0),
            "nonfinite": s[3], "mean_rows": s[4] / steps, "mean_contacts": s[5] / steps, "env_steps": s[6]}
