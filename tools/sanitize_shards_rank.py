#!/usr/bin/env python3
"""One rank of a two-process sharded pile for compute-sanitizer (each process under its own sanitizer: inside ONE
process the sanitizer serialises kernel launches, and shards wait for each other inside their kernels).

    python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        compute-sanitizer --tool memcheck python tools/sanitize_shards_rank.py 30
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("box2d-mt_b200/python", "box2d-mt_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import torch.distributed as dist
import b2host, b2shard, scenes, b2cuda_types as T
rank, world_size, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dist.init_process_group("gloo")
scene = scenes.pile(60 * world_size, 12, seed=0)
scene.world_flags &= ~T.WORLD_CONTINUOUS
plan, _ = b2shard.rank_plan(scene.arrays(), rank, world_size, 2.5)
world = b2host.HostWorld(arrays=plan.arrays, gravity=scene.gravity, world_flags=scene.world_flags, device=local_rank,
                         download_bodies=True, events=True)
world.shard_configure(rank, world_size, plan.ghost_local, plan.export_local)
lower, upper = b2shard.exchange_links(dist, rank, world_size, world.shard_link())
world.shard_connect(lower, upper)
dist.barrier()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for s in range(steps):
    world.step()
i = world.step_info()
print("rank %d: %d steps, bodies %d (ghosts %d, exports %d), constraints %d" % (rank, steps, i["bodyCount"], len(plan.ghost_local),
                                                                            len(plan.export_local), i["constraintCount"]), flush=True)
dist.barrier()
dist.destroy_process_group()
