"""The multi-process side of the sharded path (bench.py --gpus N under torchrun) on CPU: two processes, `gloo`
backend, 127.0.0.1 rendezvous.  No device call is made: what is covered is everything around them -- every rank
derives the same strips from the same scene, the halo lists of neighbouring ranks line up, the link records are
exchanged and validated, timings are max-reduced and counts sum-reduced the way bench.py does it."""
import os
import socket
import sys

import numpy as np
import pytest

import b2cuda_types as T
import b2shard
import scenes

WORLD_SIZE = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, margin, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        scene = scenes.pile(60, 8, seed=0)
        arrays = scene.arrays()
        plan, bounds = b2shard.rank_plan(arrays, rank, world_size, margin)

        # every rank must have derived the same strip boundaries
        all_bounds = [None] * world_size
        dist.all_gather_object(all_bounds, bounds.tolist())
        assert all(b == all_bounds[0] for b in all_bounds)

        # the record b2cuShardGetLink would fill (no device here: handle and pointer stay zero)
        link = np.zeros((), T.SHARD_LINK)
        link["rank"], link["rankCount"] = rank, world_size
        link["ghostCount"], link["exportCount"] = len(plan.ghost_local), len(plan.export_local)
        link["processId"] = os.getpid()
        lower, upper = b2shard.exchange_links(dist, rank, world_size, link)
        assert (lower is None) == (rank == 0) and (upper is None) == (rank == world_size - 1)
        if lower is not None:
            assert int(lower["rank"]) == rank - 1 and int(lower["ghostCount"]) == len(plan.export_local)
            assert int(lower["processId"]) != os.getpid()
        if upper is not None:
            assert int(upper["rank"]) == rank + 1 and int(upper["exportCount"]) == len(plan.ghost_local)

        # ghosts of rank r are the exports of rank r+1: same global bodies, same order
        halo = [None] * world_size
        dist.all_gather_object(halo, (plan.body_ids[plan.ghost_local].tolist(), plan.body_ids[plan.export_local].tolist()))
        for r in range(world_size - 1):
            assert halo[r][0] == halo[r + 1][1] and len(halo[r][0]) > 0

        # every dynamic body is owned exactly once: the owned counts add up to the scene
        owned = len(plan.body_ids) - len(plan.ghost_local) - int((arrays[0]["type"] != T.DYNAMIC_BODY).sum())
        total = b2shard.reduce_scalar(dist, float(owned), "sum")
        assert int(total) == int((arrays[0]["type"] == T.DYNAMIC_BODY).sum())
        # timings: the slowest rank counts
        slowest = b2shard.reduce_scalar(dist, 1.0 + rank, "max")
        assert slowest == float(world_size)

        # a rank whose halo does not match its neighbour's is refused before any device memory is mapped
        bad = link.copy()
        bad["ghostCount"] = int(link["ghostCount"]) + (1 if rank == 0 else 0)
        try:
            b2shard.exchange_links(dist, rank, world_size, bad)
            refused = False
        except RuntimeError:
            refused = True
        assert refused
        dist.barrier()
        with open(os.path.join(out_dir, "ok%d" % rank), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_plumbing_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(WORLD_SIZE, port, 1.5, str(tmp_path)), nprocs=WORLD_SIZE, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(WORLD_SIZE))
