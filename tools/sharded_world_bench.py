#!/usr/bin/env python3
"""b2CudaShardedWorld (one process, one host thread + one GPU per strip) on the bench pile: ms per Step, with and
without the per-step body download, and the cost of one Rebalance().

    python tools/sharded_world_bench.py GPUS [BODIES] [SETTLE] [STEPS]

Wall clock around b2CudaShardedWorld::Step (the strips run concurrently; the call returns when all are done)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "box2d-mt_b200", "python"))
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import b2host  # noqa: E402
import scenes  # noqa: E402


def main():
    gpus = int(sys.argv[1])
    bodies = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    settle = int(sys.argv[3]) if len(sys.argv) > 3 else 120
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    scene = scenes.pile(max(16, bodies // 100), 100, seed=0)
    t0 = time.perf_counter()
    host = b2host.HostWorld(scene, events=False)
    t1 = time.perf_counter()
    sharded = host.shard(gpus, margin=2.0)
    t2 = time.perf_counter()
    sharded.set_transport(False, False)
    for _ in range(settle):
        sharded.step()
    out = {"gpus": gpus, "bodies": host.counts()[0], "build_scene_s": t1 - t0, "shard_s": t2 - t1}
    for name, download in (("resident", False), ("with_body_download", True)):
        sharded.set_transport(download, False)
        for _ in range(3):
            sharded.step()
        ta = time.perf_counter()
        for _ in range(steps):
            sharded.step()
        out["ms_per_step_" + name] = 1e3 * (time.perf_counter() - ta) / steps
    sharded.set_transport(False, False)
    ta = time.perf_counter()
    sharded.rebalance()
    out["rebalance_s"] = time.perf_counter() - ta
    out["lost_contacts"] = sharded.lost_contacts()
    for _ in range(3):
        sharded.step()
    ta = time.perf_counter()
    for _ in range(steps):
        sharded.step()
    out["ms_per_step_resident_after_rebalance"] = 1e3 * (time.perf_counter() - ta) / steps
    out["strip_bodies"] = [len(sharded.strip_plan(r)[0]) for r in range(gpus)]
    out["body_steps_per_s_resident"] = out["bodies"] / (out["ms_per_step_resident"] * 1e-3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
