#!/usr/bin/env python3
"""Developer tool: step two device worlds built from the same scene and report the first step / field where they
differ.  usage: determinism_probe.py COLUMNS ROWS STEPS"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "box2d-mt_b200", "python"))
import numpy as np
import b2host, scenes
import b2cuda_types as T
cols, rows, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
scene = scenes.pile(cols, rows)
a = b2host.HostWorld(scene, download_bodies=False, events=False)
b = b2host.HostWorld(scene, download_bodies=False, events=False)
for s in range(steps):
    a.step(); b.step()
    ia, ib = a.step_info(), b.step_info()
    ba, bb = a.bodies(), b.bodies()
    diff_fields = [f for f in ba.dtype.names if not np.array_equal(ba[f].view(np.uint32), bb[f].view(np.uint32))]
    counters = {k: (int(ia[k]), int(ib[k])) for k in ("contactCount", "constraintCount", "colourCount", "newContactCount", "moveCount", "touchingCount", "overflowCount") if int(ia[k]) != int(ib[k])}
    if diff_fields or counters:
        print("step", s, "differs: body fields", diff_fields, "counters", counters)
        for f in diff_fields[:3]:
            idx = np.nonzero(ba[f].view(np.uint32) != bb[f].view(np.uint32))[0]
            print("  ", f, len(idx), "bodies, first", idx[:8], ba[f][idx[:4]], bb[f][idx[:4]])
        da, db = a.device_world(), b.device_world()
        ca, cb = da.get_contacts(), db.get_contacts()
        print("   contacts", len(ca), len(cb))
        if len(ca) == len(cb):
            for f in ("proxyA", "proxyB", "flags"):
                print("   contact", f, "equal:", np.array_equal(ca[f], cb[f]))
            m = ca["manifold"].tobytes() == cb["manifold"].tobytes()
            print("   manifolds equal:", m)
        ka, kb = da.solver_order(), db.solver_order()
        print("   solver order equal:", np.array_equal(ka[0], kb[0]), np.array_equal(ka[1], kb[1]))
        break
else:
    print("identical for", steps, "steps")
