"""N>1 host logic on CPU: shard ranges and the end-of-run statistics gather over gloo (world_size 2)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import evgsim
from evgsim import dist as evd


def test_shard_range_partitions_exactly():
    for total in (1, 7, 4096, 1048576, 1000003):
        for world in (1, 2, 3, 8):
            got = [evd.shard_range(total, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == total
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = {"episodes": 10 + rank, "wins": [3 + rank, 4], "ties": 3, "total_turns": 1500 * (rank + 1),
             "total_score": [100, 200 * rank], "status_count": [0, 9 + rank, 1, 0], "env_turns": 1000 * (rank + 1),
             "fought_unit_slots": 500 + 7 * rank}
    total = evd.gather_episode_stats(stats)
    if rank == 0:
        torch.save(total, out)
    dist.destroy_process_group()


def test_gather_episode_stats_gloo_world2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "total.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    total = torch.load(out)
    assert total == {"episodes": 21, "wins": [7, 8], "ties": 6, "total_turns": 4500, "total_score": [200, 200],
                     "status_count": [0, 19, 2, 0], "env_turns": 3000, "fought_unit_slots": 1007}


def test_gather_without_process_group_is_identity():
    stats = {"episodes": 1, "wins": [1, 0], "ties": 0, "total_turns": 150, "total_score": [5, 6],
             "status_count": [0, 1, 0, 0], "env_turns": 150, "fought_unit_slots": 96}
    assert evd.gather_episode_stats(stats) == stats
