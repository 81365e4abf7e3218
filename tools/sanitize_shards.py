#!/usr/bin/env python3
"""Workload for compute-sanitizer on TWO GPUs: a pile cut into two shards (one device each), 40 steps through the
dataflow solver with the halo rows crossing NVLink peer memory.  compute-sanitizer --tool memcheck python tools/sanitize_shards.py"""
import os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("box2d-mt_b200/python", "box2d-mt_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import b2cuda, b2cuda_types as T, scenes
import test_sharding as ts
scene = scenes.pile(36, 8)
scene.world_flags &= ~T.WORLD_CONTINUOUS
worlds, plans = ts._shard_worlds(b2cuda, scene, 2, margin=2.5)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for s in range(steps):
    infos = ts._step_all(worlds)
print("2 shards, %d steps, constraints %s" % (steps, [int(i["constraintCount"]) for i in infos]))
