"""Multi-GPU plumbing: matches are independent, so ranks own contiguous ranges of global match ids
and the step path has NO collective.  The only exchange is an end-of-run gather of the episode
statistics (a few hundred bytes per rank) over torch.distributed (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations


def shard_range(total_envs: int, rank: int, world_size: int):
    """Contiguous range [first, first+count) of global match ids owned by `rank`."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside 0..%d" % (rank, world_size - 1))
    base, rem = divmod(int(total_envs), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


STAT_KEYS = ["episodes", "wins0", "wins1", "ties", "total_turns", "score0", "score1",
             "status0", "status1", "status2", "status3", "env_turns", "fought_unit_slots"]


def stats_to_vector(stats: dict):
    return [stats["episodes"], stats["wins"][0], stats["wins"][1], stats["ties"], stats["total_turns"],
            stats["total_score"][0], stats["total_score"][1], *stats["status_count"], stats["env_turns"],
            stats.get("fought_unit_slots", 0)]


def vector_to_stats(v) -> dict:
    v = [int(x) for x in v]
    return {"episodes": v[0], "wins": [v[1], v[2]], "ties": v[3], "total_turns": v[4], "total_score": [v[5], v[6]],
            "status_count": v[7:11], "env_turns": v[11], "fought_unit_slots": v[12]}


def gather_episode_stats(stats: dict, device=None) -> dict:
    """Sum per-rank episode statistics over the default process group (all_gather then add)."""
    import torch
    import torch.distributed as dist

    vec = torch.tensor(stats_to_vector(stats), dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return vector_to_stats(vec.tolist())
    parts = [torch.zeros_like(vec) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, vec)
    return vector_to_stats(torch.stack(parts).sum(dim=0).tolist())
