"""Multi-GPU sharding: environments are independent, so ranks own contiguous env ranges and never communicate on the
step path.  The only collective is the optional all-reduce of the 16-slot episode-statistics vector."""
import torch
import torch.distributed as dist


def shard(total_envs, rank, world):
    """[start, stop) of the global env ids owned by `rank`; RNG streams are keyed by global id (env_id_base = start)."""
    per, rem = divmod(int(total_envs), int(world))
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)


def all_reduce_stats(stats):
    """SUM the per-rank statistics vector in place (NCCL on GPU tensors, gloo on CPU tensors); no-op for world size 1."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def summarize(stats, names):
    v = stats.detach().cpu().tolist()
    d = {k: v[i] for i, k in enumerate(names)}
    ep = max(d.get("episodes", 0.0), 1.0)
    d["mean_return"] = d.get("return_sum", 0.0) / ep
    d["mean_length"] = d.get("length_sum", 0.0) / ep
    d["success_rate"] = d.get("successes", 0.0) / ep
    return d
