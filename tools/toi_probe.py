"""Developer probe: scenes with continuous physics ON stepped in lockstep with the oracle; prints the first divergence."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("box2d-mt_b200/python", "box2d-mt_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import b2cuda
import b2cuda_types as T
import parity
import ref
import scenes

CASES = {
    "hello": (scenes.hello_world, 90),
    "pyramid6": (lambda: scenes.pyramid(6), 200),
    "bullets": (scenes.bullets, 200),
    "add_pair": (lambda: scenes.add_pair(300), 90),
    "chains": (lambda: scenes.chain_terrain(40), 200),
    "pyramid20": (lambda: scenes.pyramid(20), 150),
    "sensors": (lambda: scenes.sensors(30), 200),
}
names = sys.argv[1:] or list(CASES)
teacher = bool(os.environ.get("TEACHER"))
for name in names:
    make, steps = CASES[name]
    scene = make()
    assert scene.world_flags & T.WORLD_CONTINUOUS
    r = ref.RefWorld(scene)
    g = parity.gpu_world_from_ref(b2cuda, r)
    sub = ev = new = 0
    try:
        def on_step(s, gpu, rf):
            pass
        infos = []
        for s in range(steps):
            infos += parity.lockstep(g, r, 1, teacher=False, tol=0.0)
            i = infos[-1]
            sub += int(i["toiSubSteps"]); ev += int(i["toiEventCount"]); new += int(i["toiNewContactCount"])
        print("%-10s OK   %d steps, %d sub-steps, %d toi events, %d toi contacts, max candidates %d" %
              (name, steps, sub, ev, new, max(int(i["toiCandidateCount"]) for i in infos)))
    except Exception as e:
        print("%-10s FAIL at step %d after %d sub-steps (%d toi events): %s" % (name, len(infos), sub, ev, str(e)[:600]))
        if os.environ.get("TRACE"):
            traceback.print_exc()
