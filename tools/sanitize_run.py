#!/usr/bin/env python3
"""Workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): a few steps of each kind of world
through the C ABI -- a 100k-body pile (every discrete kernel incl. the persistent solver), a joint scene (the joint
instance of the solver kernels), the bullets scene with continuous physics (every time-of-impact kernel), Add Pair
(pair search under churn, serial overflow list).

    compute-sanitizer --tool memcheck python tools/sanitize_run.py [pile|joints|bullets|add_pair ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("box2d-mt_b200/python", "box2d-mt_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import b2cuda
import scenes

CASES = {
    "pile": (lambda: scenes.pile(int(os.environ.get("SAN_COLUMNS", "1000")), 100), 6),
    # a small pile that has settled: full colour set, warm starting, position iterations with early exits, sleeping
    "pile_settled": (lambda: scenes.pile(60, 20, sleep=True), 150),
    "joints": (scenes.machines, 30),
    "bullets": (scenes.bullets, 60),
    "add_pair": (lambda: scenes.add_pair(600), 30),
}
for name in (sys.argv[1:] or list(CASES)):
    make, steps = CASES[name]
    scene = make()
    bodies, shapes, fixtures = scene.arrays()
    import b2host
    w = b2host.HostWorld(scene, download_bodies=True, events=True)
    sub = 0
    for _ in range(steps):
        w.step()
        sub += int(w.step_info()["toiSubSteps"])
    i = w.step_info()
    print("%s: %d steps, bodies %d contacts %d constraints %d, toi sub-steps %d" % (name, steps, i["bodyCount"], i["contactCount"],
                                                                                 i["constraintCount"], sub), flush=True)
