#!/usr/bin/env python3
"""Developer tool (not part of the product path or the bench): per-phase timing of b2cuStep on a pile scene whose
initial state is exported from the oracle.  usage: perf_probe.py COLUMNS ROWS STEPS [REPORT_EVERY]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("box2d-mt_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import b2cuda, parity, ref, scenes

cols, rows, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
every = int(sys.argv[4]) if len(sys.argv) > 4 else max(1, steps // 10)
t = time.time()
scene = scenes.pile(cols, rows)
scene.world_flags &= ~4
r = ref.RefWorld(scene, threads=8)
print("oracle build %.2fs" % (time.time() - t), r.counts(), flush=True)
t = time.time()
g = parity.gpu_world_from_ref(b2cuda, r)
print("gpu upload %.2fs" % (time.time() - t), flush=True)
fields = ["step", "collide", "solveTraversal", "solveInit", "solveVelocity", "solvePosition", "broadphase", "solveTOI"]
acc = np.zeros(len(fields))
wall = 0.0
n = 0
for s in range(steps):
    t0 = time.perf_counter()
    info = g.step()
    wall += time.perf_counter() - t0
    acc += [float(info[f]) for f in fields]
    n += 1
    if (s + 1) % every == 0 or s == steps - 1:
        print("step %4d wall %.3f ms | " % (s + 1, 1e3 * wall / n) + " ".join("%s %.3f" % (f, a / n) for f, a in zip(fields, acc))
              + " | contacts %d touching %d constraints %d colours %d overflow %d moved %d new %d destroyed %d launches %d"
              % (info["contactCount"], info["touchingCount"], info["constraintCount"], info["colourCount"], info["overflowCount"],
                 info["moveCount"], info["newContactCount"], info["destroyedContactCount"], info["kernelLaunches"]), flush=True)
        acc[:] = 0; wall = 0.0; n = 0
b = g.get_bodies()
print("min y %.3f max y %.3f  max |v| %.3f" % (b["py"][1:].min(), b["py"][1:].max(), np.hypot(b["vx"], b["vy"]).max()))
